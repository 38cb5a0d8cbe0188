"""Multi-condition render of one view (gsplat/olat.py, BASELINE config 3): C colour sets against one projection and one
tile binning must give, per condition, exactly the pixels of the fused single-condition render, and gradients equal to
the sum over the conditions of the single-condition gradients."""
import numpy as np
import pytest
import torch

from util import assert_close, small_scene, t2n

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["multi", "single"])
def olat_mode(request):
    """four conditions per blend pass (default) or one condition per pass"""
    from goliath_b200.gsplat import olat

    before = olat.MODE
    olat.MODE = request.param
    yield request.param
    olat.MODE = before


@pytest.mark.parametrize("capacity,C", [(None, 4), (1 << 16, 6), (1 << 16, 9)])
def test_render_shared_equals_per_condition_renders(cuda, capacity, C, olat_mode):
    from goliath_b200.gsplat.fused import check_overflow, render_fused
    from goliath_b200.gsplat.olat import render_shared

    s = small_scene(G=3000, img_h=96, img_w=80)
    mult = 12.0
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    H, W = s["img_h"], s["img_w"]
    gen = torch.Generator().manual_seed(3)
    cols = torch.rand(C, 3000, 3, generator=gen).to(cuda)
    w_rgb = torch.randn(C, H, W, 3, generator=gen).to(cuda)
    w_dep = torch.randn(H, W, generator=gen).to(cuda) * 1e-3
    w_alpha = torch.randn(H, W, generator=gen).to(cuda)
    bg = torch.tensor([0.1, 0.2, 0.3], device=cuda)
    cam = (d(s["viewmat"]), s["fx"], s["fy"], s["cx"], s["cy"], H, W)

    def leaves():
        return [d(s["means3d"]).requires_grad_(), d(s["scales"] * np.float32(mult)).requires_grad_(),
                d(s["quats"]).requires_grad_(), d(s["opacity"]).requires_grad_()]

    m, sc, q, op = leaves()
    cl = cols.clone().requires_grad_()
    rgb, depth_raw, alpha, radii = render_shared(m, sc, 1.0, q, *cam, op, cl, bg, 0.1, capacity)
    ((rgb * w_rgb).sum() + (depth_raw * w_dep).sum() + (alpha * w_alpha).sum()).backward()
    torch.cuda.synchronize()
    assert not check_overflow(cuda)
    got = [t2n(x.grad) for x in (m, sc, q, op)] + [t2n(cl.grad)]

    sums = None
    for c in range(C):
        m1, sc1, q1, op1 = leaves()
        c1 = cols[c].clone().requires_grad_()
        out4, a1, r1 = render_fused(m1, sc1, 1.0, q1, *cam, op1, c1, bg, 0.1, capacity)
        assert torch.equal(out4[..., :3], rgb[c]), "condition %d pixels" % c
        assert torch.equal(a1, alpha) and torch.equal(r1, radii)
        loss = (out4[..., :3] * w_rgb[c]).sum()
        if c == 0:  # depth and alpha are produced once per view, by condition 0
            assert torch.equal(out4[..., 3], depth_raw)
            loss = loss + (out4[..., 3] * w_dep).sum() + (a1 * w_alpha).sum()
        loss.backward()
        g = [t2n(x.grad) for x in (m1, sc1, q1, op1)]
        sums = g if sums is None else [a + b for a, b in zip(sums, g)]
        assert_close(got[4][c], t2n(c1.grad), rtol=1e-4, atol=1e-5 * float(c1.grad.abs().max()), frac=0.999,
                     what="grad colors[%d]" % c)
    for name, a, b in zip(("means3d", "scales", "quats", "opacity"), got[:4], sums):
        assert_close(a, b, rtol=1e-4, atol=2e-5 * float(np.abs(b).max()), frac=0.999, what="grad " + name)


def test_render_views_shared_shapes_and_depth(cuda, olat_mode):
    from goliath_b200.gsplat.olat import render_views_shared
    from goliath_b200.render import render_views

    s = small_scene(G=2000, img_h=64, img_w=48)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    V, C, G, H, W = 2, 3, 2000, 64, 48
    gen = torch.Generator().manual_seed(5)
    geom = dict(primpos=d(s["means3d"])[None].repeat(V, 1, 1), primqvec=d(s["quats"])[None].repeat(V, 1, 1),
                primscale=(d(s["scales"]) * 10.0)[None].repeat(V, 1, 1), opacity=d(s["opacity"])[None].repeat(V, 1, 1))
    cols = torch.rand(V, C, G, 3, generator=gen).to(cuda)
    Rt = d(s["viewmat"])[None].repeat(V, 1, 1)
    intr = [(s["fx"], s["fy"], s["cx"], s["cy"])] * V
    rgb, alpha, depth = render_views_shared(W, H, Rt, geom, cols, intr, capacity=1 << 16)
    assert rgb.shape == (V, C, 3, H, W) and alpha.shape == (V, 1, H, W) and depth.shape == (V, 1, H, W)
    assert not alpha.requires_grad
    for c in range(C):  # against the per-view renderer of the single-condition path
        preds = dict(geom, color=cols[:, c])
        r1, a1, d1 = render_views(W, H, None, Rt, preds, intrinsics_host=intr, capacity=1 << 16)
        assert torch.equal(r1, rgb[:, c]) and torch.equal(a1, alpha)
        assert_close(t2n(d1), t2n(depth), rtol=1e-6, atol=1e-6, what="depth")
