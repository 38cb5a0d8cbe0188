"""Fused post-render chain + photometric losses (csrc/photo_loss.cu) against vectors produced by the
reference's own CalV5, LearnableBlur, rgb_l1 and rgb_ssim (tests/golden/photo_loss_ref.npz, made by
tests/golden/make_photo_loss_golden.py)."""
import os

import numpy as np
import pytest
import torch

from util import assert_close, t2n

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "photo_loss_ref.npz")


def _t(z, k, dev, grad=False, dtype=torch.float32):
    t = torch.from_numpy(z[k]).to(dtype).to(dev).contiguous()
    return t.requires_grad_() if grad else t


def test_post_render_and_losses_match_reference_modules(cuda):
    from goliath_b200.photo_loss import photometric_loss, post_render

    z = np.load(GOLD)
    rgb, cw, cb, bw = (_t(z, k, cuda, True) for k in ("rgb", "cal_w", "cal_b", "blur_w"))
    grey = _t(z, "grey", cuda, dtype=torch.int32)
    assert int(grey.sum()) == 1
    pred = post_render(rgb, _t(z, "alpha", cuda), _t(z, "bg", cuda), cw, cb, grey, bw)
    assert_close(t2n(pred), z["pred"], rtol=1e-5, atol=1e-3, what="pred (cal + bg + blur)")     # values ~ 0..500
    loss, parts = photometric_loss(pred, _t(z, "image", cuda), _t(z, "mask", cuda), 10.0, 0.2)
    assert abs(float(parts["rgb_l1"]) - float(z["l1"])) < 1e-5 * float(z["l1"])
    assert abs(float(parts["rgb_ssim"]) - float(z["ssim_loss"])) < 1e-5
    assert abs(float(loss) - float(z["loss"])) < 1e-5 * float(z["loss"])
    loss.backward()
    # the identity camera's frame (w = 1, b = 0 rows): the reference never touches those parameters, the fused rows get the
    # plain d/dw, d/db of an affine map — compare the calibrated frames only
    cal_frames = ~((z["cal_w"] == 1.0).all(1) & (z["cal_b"] == 0.0).all(1))
    assert cal_frames.sum() == 3
    for name, got, want in (("rgb", rgb.grad, z["g_rgb"]), ("cal_w", cw.grad[cal_frames], z["g_cal_w"][cal_frames]),
                            ("cal_b", cb.grad[cal_frames], z["g_cal_b"][cal_frames]), ("blur_w", bw.grad, z["g_blur_w"])):
        assert_close(t2n(got), want, rtol=1e-4, atol=1e-5 * float(np.abs(want).max()), frac=0.9995, what="grad " + name)


def test_loss_gradient_alone_and_optional_stages(cuda):
    from goliath_b200.photo_loss import photometric_loss, post_render

    z = np.load(GOLD)
    pred = _t(z, "pred", cuda, True)
    loss, _ = photometric_loss(pred, _t(z, "image", cuda), _t(z, "mask", cuda), 10.0, 0.2)
    (3.0 * loss).backward()      # upstream gradient is read on the device
    want = 3.0 * z["g_pred"]
    assert_close(t2n(pred.grad), want, rtol=1e-4, atol=1e-5 * float(np.abs(want).max()), frac=0.9995, what="dL/dpred")
    # every stage optional: no cal, no background, no blur is the identity with an identity gradient
    rgb = _t(z, "rgb", cuda, True)
    out = post_render(rgb)
    assert torch.equal(out, rgb)
    out.sum().backward()
    assert torch.equal(rgb.grad, torch.ones_like(rgb))
    # cal only == x1 of the reference; background only
    x1 = post_render(_t(z, "rgb", cuda), cal_w=_t(z, "cal_w", cuda), cal_b=_t(z, "cal_b", cuda),
                     grey=_t(z, "grey", cuda, dtype=torch.int32))
    assert_close(t2n(x1), z["x1"], rtol=1e-5, atol=1e-3, what="cal")
    with pytest.raises(RuntimeError):
        post_render(rgb, background=_t(z, "bg", cuda))           # needs alpha
    with pytest.raises(RuntimeError):
        post_render(rgb.detach().cpu())


def test_train_step_render_to_loss_chain(cuda):
    """render_views -> post_render -> photometric_loss -> backward: the gradient reaches the Gaussians (the blend backward
    consumes post_render's dL/d(rendered rgb) as v_out), at a non-tile-multiple image size."""
    from goliath_b200.photo_loss import photometric_loss, post_render
    from goliath_b200.render import render_views
    from util import small_scene

    s = small_scene(G=3000, img_h=70, img_w=93, seed=11, cam=3)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    H, W = s["img_h"], s["img_w"]
    leaves = dict(primpos=d(s["means3d"])[None].requires_grad_(), primqvec=d(s["quats"])[None],
                  primscale=(d(s["scales"]) * 15.0)[None].requires_grad_(), opacity=d(s["opacity"])[None].requires_grad_(),
                  color=(d(s["colors"]) * 255.0)[None].requires_grad_())
    rgb, alpha, depth = render_views(W, H, None, d(s["viewmat"])[None], leaves, intrinsics_host=[(s["fx"], s["fy"], s["cx"], s["cy"])],
                                     capacity=1 << 17)
    gen = torch.Generator().manual_seed(1)
    target = (255 * torch.rand(1, 3, H, W, generator=gen)).to(cuda)
    mask = (torch.rand(1, 1, H, W, generator=gen) > 0.2).float().to(cuda)
    bw = torch.softmax(torch.randn(1, 3, generator=gen), -1).to(cuda).requires_grad_()
    pred = post_render(rgb, alpha, (255 * torch.rand(1, 3, H, W, generator=gen)).to(cuda),
                       torch.tensor([[1.1, 0.9, 1.0]], device=cuda), torch.tensor([[2.0, -1.0, 0.5]], device=cuda), None, bw)
    loss, parts = photometric_loss(pred, target, mask)
    loss.backward()
    for k in ("primpos", "primscale", "opacity", "color"):
        g = leaves[k].grad
        assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0, k
    assert torch.isfinite(bw.grad).all() and float(bw.grad.abs().max()) > 0
