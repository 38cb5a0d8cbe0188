"""Generate tests/golden/geom_ref.npz with the REFERENCE's own ca_code/utils/geom.py (`vert_normals`, `values_to_uv`),
imported from /root/reference (build container only), on a small synthetic mesh: a perturbed grid surface with its
index / barycentric images rasterised analytically.  Stores inputs, outputs and autograd gradients.

Usage: python tests/golden/make_geom_golden.py"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch as th

sys.path.insert(0, "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geom_ref.npz")


def main():
    for name in ("pytorch3d", "pytorch3d.renderer", "pytorch3d.renderer.mesh", "pytorch3d.renderer.mesh.rasterize_meshes",
                 "pytorch3d.structures", "pytorch3d.io", "drtk", "addict", "omegaconf", "igl", "trimesh"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock() if name != "addict" else types.SimpleNamespace(Dict=dict)
    from ca_code.utils import geom

    g = th.Generator().manual_seed(308)
    n = 9                                   # 9 x 9 grid of vertices, 128 triangles
    ys, xs = th.meshgrid(th.arange(n), th.arange(n), indexing="ij")
    base = th.stack([xs, ys, th.zeros_like(xs)], -1).reshape(-1, 3).double()
    B = 2
    v = (base[None] + 0.3 * th.randn(B, n * n, 3, generator=g, dtype=th.float64)).requires_grad_()
    quads = [(y * n + x, y * n + x + 1, (y + 1) * n + x, (y + 1) * n + x + 1) for y in range(n - 1) for x in range(n - 1)]
    vi = th.tensor([t for a, b, c, d in quads for t in ((a, b, c), (b, d, c))], dtype=th.long)
    vn = geom.vert_normals(v, vi)
    w = th.randn(vn.shape, generator=g, dtype=th.float64)
    (g_v,) = th.autograd.grad((vn * w).sum(), [v])

    # index / bary images of a U x U texture: texel -> the triangle of the grid cell it falls in; a border of uncovered texels
    U = 40
    idx = th.full((U, U, 3), -1, dtype=th.int32)
    bary = th.zeros(U, U, 3, dtype=th.float32)
    for ty in range(3, U - 2):
        for tx in range(2, U - 3):
            fx, fy = (tx - 2 + 0.5) / (U - 5) * (n - 1), (ty - 3 + 0.5) / (U - 5) * (n - 1)
            cx, cy = min(int(fx), n - 2), min(int(fy), n - 2)
            lx, ly = fx - cx, fy - cy
            a, b_, c, d = quads[cy * (n - 1) + cx]
            if lx + ly <= 1.0:
                idx[ty, tx] = th.tensor([a, b_, c]); bary[ty, tx] = th.tensor([1 - lx - ly, lx, ly])
            else:
                idx[ty, tx] = th.tensor([b_, d, c]); bary[ty, tx] = th.tensor([1 - ly, lx + ly - 1, 1 - lx])
    vals = th.randn(B, n * n, 5, generator=g, dtype=th.float64).requires_grad_()
    uv = geom.values_to_uv(vals, idx, bary.double())
    w2 = th.randn(uv.shape, generator=g, dtype=th.float64)
    (g_vals,) = th.autograd.grad((uv * w2).sum(), [vals])
    d = dict(v=v, vi=vi, vn=vn, w_vn=w, g_v=g_v, index=idx, bary=bary, values=vals, uv=uv, w_uv=w2, g_values=g_vals)
    np.savez_compressed(OUT, **{k: t.detach().numpy() for k, t in d.items()})
    print("wrote %s (%d bytes); covered texels %d / %d" % (OUT, os.path.getsize(OUT), int((idx[..., 0] >= 0).sum()), U * U))


if __name__ == "__main__":
    main()
