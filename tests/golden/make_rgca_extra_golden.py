"""Generate tests/golden/rgca_extra_ref.npz from the REFERENCE's own code for the two remaining branches of
`PrimDecoder.forward` (/root/reference/ca_code/models/rgca.py):
  * :548-556  environment-map specular branch (einsum rotation, `dir2uv`, `mipmap_grid_sample`, clamp * spec_vis), with
              the reference's own `ca_code.utils.envmap.dir2uv` and `ca_code.utils.mipmap_sampler.mipmap_grid_sample`;
  * :590-618  training-mode random back light (`cos_weight`, `color_rand`), with the reference's `sh.dir2sh_torch`.
Nothing is copied: the statements are sliced out of the reference file at run time (AST) and executed on small random
inputs; inputs, outputs and autograd gradients go to the fixture.  Needs /root/reference (build container only).

Usage: python tests/golden/make_rgca_extra_golden.py
"""
import ast
import os
import sys
import types

import numpy as np
import torch as th
import torch.nn.functional as F

sys.path.insert(0, "/root/reference")
REF = "/root/reference/ca_code/models/rgca.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rgca_extra_ref.npz")


def forward_lines():
    src = open(REF).read()
    lines = src.split("\n")
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PrimDecoder"][0]
    fwd = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward"][0]
    return lines, fwd


def block_after(lines, fwd, header):
    """the indented body of the `if ...:` statement whose header line starts with `header`, dedented"""
    i0 = [i for i in range(fwd.lineno, fwd.end_lineno) if lines[i].strip().startswith(header)][0]
    ind = len(lines[i0]) - len(lines[i0].lstrip())
    body = []
    for l in lines[i0 + 1:fwd.end_lineno]:
        if l.strip() and (len(l) - len(l.lstrip())) <= ind:
            break
        body.append(l)
    bi = min(len(l) - len(l.lstrip()) for l in body if l.strip())
    return "\n".join(l[bi:] for l in body), (i0 + 2, i0 + 1 + len(body))


def main():
    if not os.path.exists(REF):
        sys.exit("needs /root/reference (build container only)")
    from ca_code.utils import sh
    from ca_code.utils.envmap import dir2uv
    from ca_code.utils.mipmap_sampler import mipmap_grid_sample

    lines, fwd = forward_lines()
    g = th.Generator().manual_seed(548)
    f64 = dict(generator=g, dtype=th.float64)
    d = {}

    # ---------------- environment-map branch (:548-556)
    text, span = block_after(lines, fwd, "if preconv_envmap is not None:")
    B, G = 2, 301
    ref_dirs = F.normalize(th.randn(B, G, 3, **f64), dim=-1)
    ref_dirs[0, 0] = th.tensor([0.0, 1.0, 0.0], dtype=th.float64)      # the poles of dir2uv
    ref_dirs[0, 1] = th.tensor([0.0, -1.0, 0.0], dtype=th.float64)
    ref_dirs[0, 2] = th.tensor([0.0, 0.0, -1.0], dtype=th.float64)     # the u = +-1 seam
    ref_dirs.requires_grad_()
    sigma = (0.1 * th.exp(th.randn(B, G, **f64))).clamp(min=0.01)      # sigma*5 spans all four levels and beyond
    sigma[1, :5] = th.tensor([0.0, 0.2, 0.4, 0.6, 5.0], dtype=th.float64)
    spec_vis = th.sigmoid(th.randn(B, G, 1, **f64)).requires_grad_()
    ax = F.normalize(th.randn(B, 3, **f64), dim=-1)
    ang = th.rand(B, **f64) * 6.28
    K = th.zeros(B, 3, 3, dtype=th.float64)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 2], ax[:, 1], ax[:, 2], -ax[:, 0], -ax[:, 1], ax[:, 0]
    lightrot = th.eye(3, dtype=th.float64)[None] + th.sin(ang)[:, None, None] * K + (1 - th.cos(ang))[:, None, None] * (K @ K)
    preconv_envmap = [2.5 * th.rand(B, 3, 16 >> i, 32 >> i, **f64) for i in range(4)]  # values above 1 hit clamp(max=1)
    ns = dict(th=th, F=F, dir2uv=dir2uv, mipmap_grid_sample=mipmap_grid_sample, ref_dirs=ref_dirs, sigma=sigma,
              spec_vis=spec_vis, lightrot=lightrot, preconv_envmap=preconv_envmap)
    exec(text, ns)
    spec = ns["spec_color"]
    w = th.randn(spec.shape, **f64)
    g_dirs, g_vis = th.autograd.grad((spec * w).sum(), [ref_dirs, spec_vis])
    d.update(env_ref_dirs=ref_dirs, env_sigma=sigma, env_spec_vis=spec_vis, env_lightrot=lightrot, env_w=w,
             env_spec=spec, env_g_ref_dirs=g_dirs, env_g_spec_vis=g_vis)
    d.update({"env_level%d" % i: t for i, t in enumerate(preconv_envmap)})
    print("environment-map branch: executed reference lines %d-%d" % span)

    # ---------------- training-mode branch (:590-618)
    text, span = block_after(lines, fwd, "if self.training:")
    B, G = 2, 77
    n_color, n_mono = 16, 65
    planes = th.randn(B, 3 * n_color + n_mono, G, **f64).requires_grad_()      # f_vnocond[:, :113] as [B,113,G]
    ds = planes.permute(0, 2, 1)
    diff_shs = th.cat([ds[..., :3 * n_color].reshape(B, -1, 3, n_color),
                       ds[..., 3 * n_color:].reshape(B, -1, 1, n_mono).expand(-1, -1, 3, -1)], -1)  # as rgca.py:506-516
    spec_nml = F.normalize(th.randn(B, G, 3, **f64), dim=-1)
    light_intensity = th.rand(B, 5, 3, **f64)
    n_lights = th.full((B,), 5, dtype=th.int32)
    headrel_light_pos = th.randn(B, 5, 3, **f64)
    self = types.SimpleNamespace(diff_sh_degree=8)
    preds = {}
    th.manual_seed(590)
    ns = dict(th=th, F=F, sh=sh, self=self, B=B, diff_shs=diff_shs, spec_nml=spec_nml, light_intensity=light_intensity,
              n_lights=n_lights, headrel_light_pos=headrel_light_pos, preds=preds)
    exec(text, ns)
    w = th.randn(preds["color_rand"].shape, **f64)
    (g_planes,) = th.autograd.grad((preds["color_rand"] * w).sum(), [planes])
    d.update(tr_planes=planes, tr_spec_nml=spec_nml, tr_light_dir=ns["light_dir"], tr_light_sh=ns["light_sh"],
             tr_cos_weight=preds["cos_weight"], tr_color_rand=preds["color_rand"], tr_w=w, tr_g_planes=g_planes)
    print("training branch: executed reference lines %d-%d" % span)
    np.savez_compressed(OUT, **{k: v.detach().numpy() for k, v in d.items()})
    print("wrote %s (%d bytes)" % (OUT, os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
