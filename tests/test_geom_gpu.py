"""goliath_b200.geom (csrc/geom_uv.cu) against the reference's own ca_code/utils/geom.py functions
(tests/golden/geom_ref.npz, made by tests/golden/make_geom_golden.py): vertex normals and the vertex -> UV gather,
forward and backward, plus the GeometryModule-shaped holder inside the RGCA decoder."""
import os

import numpy as np
import pytest
import torch

from util import assert_close, t2n

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geom_ref.npz")


def test_vert_normals_and_values_to_uv_match_reference(cuda):
    from goliath_b200.geom import values_to_uv, vert_normals

    z = np.load(GOLD)
    f = lambda k: torch.from_numpy(z[k]).float().to(cuda)
    v = f("v").requires_grad_()
    vn = vert_normals(v, torch.from_numpy(z["vi"]).to(cuda))
    assert_close(t2n(vn), z["vn"], rtol=1e-5, atol=1e-6, what="vert_normals")
    (vn * f("w_vn")).sum().backward()
    assert_close(t2n(v.grad), z["g_v"], rtol=1e-4, atol=1e-5 * float(np.abs(z["g_v"]).max()), what="grad verts")
    vals = f("values").requires_grad_()
    uv = values_to_uv(vals, torch.from_numpy(z["index"]).to(cuda), f("bary"))
    assert uv.shape == z["uv"].shape
    assert_close(t2n(uv), z["uv"], rtol=1e-5, atol=1e-6, what="values_to_uv")
    assert (t2n(uv)[:, :, z["index"][..., 0] < 0] == 0).all(), "uncovered texels are zero"
    (uv * f("w_uv")).sum().backward()
    assert_close(t2n(vals.grad), z["g_values"], rtol=1e-4, atol=1e-5 * float(np.abs(z["g_values"]).max()), what="grad values")


def test_geometry_module_feeds_the_decoder(cuda):
    """PrimDecoder.forward with a real geo_fn: postex = to_uv(geom), tn = normalize(to_uv(vn(geom))) (rgca.py:478-491)."""
    from goliath_b200.geom import GeometryModule
    from goliath_b200.rgca import PrimDecoder

    z = np.load(GOLD)
    S = 128
    # stretch the fixture's 40x40 index / bary images to the decoder's slab size by nearest-neighbour repetition
    rep = lambda a: np.repeat(np.repeat(a, 4, 0), 4, 1)[:S, :S]
    pad = lambda a, fill: np.pad(rep(a), ((0, S - min(S, 160)), (0, S - min(S, 160)), (0, 0)), constant_values=fill)
    geo = GeometryModule(torch.from_numpy(z["vi"]), torch.from_numpy(pad(z["index"], -1)), torch.from_numpy(pad(z["bary"], 0))).to(cuda)
    gen = torch.Generator().manual_seed(2)
    dec = PrimDecoder(16, geo, 255 * torch.rand(3, S, S, generator=gen), slabsize=S).to(cuda).eval()
    B = 2
    geom = torch.from_numpy(z["v"]).float().to(cuda).requires_grad_()
    embs = torch.randn(B, 16, generator=gen).to(cuda)
    campos = torch.tensor([[0.0, 0.0, 50.0], [10.0, 0.0, 40.0]], device=cuda)
    lint, lpos = torch.rand(B, 2, 3, generator=gen).to(cuda), (torch.randn(B, 2, 3, generator=gen) * 100).to(cuda)
    preds = dec(embs, geom, campos, lint, lpos, torch.randn(B, 3, 81, generator=gen).to(cuda),
                torch.full((B,), 2, dtype=torch.int32, device=cuda))
    assert preds["primpos"].shape == (B, S * S, 3) and torch.isfinite(preds["color"]).all()
    preds["primpos"].sum().backward()
    assert geom.grad is not None and torch.isfinite(geom.grad).all() and float(geom.grad.abs().max()) > 0
