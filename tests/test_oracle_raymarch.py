"""CPU: pin the C restatement of the MVP raymarch path (oracle/raymarch_oracle.c) against
  * the reference's own PyTorch oracle (inline raymarcher of mvpraymarch.gradcheck, fixtures in
    tests/golden/raymarch_pytorch_ref.npz made by tests/golden/make_raymarch_golden.py), forward and gradients;
  * the reference's PyTorch restatement of compute_raydirs (extensions/utils/utils.py:127-143), re-derived here.
"""
import os

import numpy as np
import pytest

from util import assert_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raymarch_pytorch_ref.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("tag,algo", [("n", 0), ("w", 1)])
def test_raymarch_forward_matches_reference_pytorch_oracle(orc, gold, tag, algo):
    g = lambda k: gold[tag + "_" + k]
    warp = g("warp") if algo == 1 else None
    rgba, raysat, _ = orc.raymarch_fwd(g("raypos"), g("raydir"), float(g("stepsize")), g("tminmax"), g("primpos"),
                                       g("primrot"), g("primscale"), g("template"), warp, algo=algo,
                                       fadescale=float(gold["fadescale"]), fadeexp=float(gold["fadeexp"]))
    assert_close(rgba, g("rayrgba"), rtol=1e-4, atol=1e-6, what="rayrgba")
    assert np.all(raysat == -1)  # no ray saturates in the fixture


@pytest.mark.parametrize("tag,algo", [("n", 0)])
def test_raymarch_backward_matches_reference_pytorch_autograd(orc, gold, tag, algo):
    g = lambda k: gold[tag + "_" + k]
    warp = g("warp") if algo == 1 else None
    args = (g("raypos"), g("raydir"), float(g("stepsize")), g("tminmax"), g("primpos"), g("primrot"), g("primscale"),
            g("template"), warp)
    rgba, raysat, _ = orc.raymarch_fwd(*args, algo=algo, fadescale=float(gold["fadescale"]), fadeexp=float(gold["fadeexp"]))
    gp, gr, gs, gt, gw = orc.raymarch_bwd(*args, raysat, np.ones_like(rgba), algo=algo,
                                          fadescale=float(gold["fadescale"]), fadeexp=float(gold["fadeexp"]))
    assert_close(gt, g("g_template"), rtol=2e-4, what="grad_template")
    assert_close(gp, g("g_primpos"), rtol=1e-3, what="grad_primpos")
    assert_close(gr, g("g_primrot"), rtol=1e-3, what="grad_primrot")
    assert_close(gs, g("g_primscale"), rtol=1e-3, what="grad_primscale")


def test_raymarch_warp_backward_template_grad(orc, gold):
    """algo=1: the template gradient is the true one; the fade / warp-grid gradients are evaluated at the WARPED
    coordinate by the reference kernels and are therefore only checked loosely."""
    g = lambda k: gold["w_" + k]
    args = (g("raypos"), g("raydir"), float(g("stepsize")), g("tminmax"), g("primpos"), g("primrot"), g("primscale"),
            g("template"), g("warp"))
    kw = dict(algo=1, fadescale=float(gold["fadescale"]), fadeexp=float(gold["fadeexp"]))
    rgba, raysat, _ = orc.raymarch_fwd(*args, **kw)
    gp, gr, gs, gt, gw = orc.raymarch_bwd(*args, raysat, np.ones_like(rgba), **kw)
    assert_close(gt, g("g_template"), rtol=2e-4, what="grad_template (warp)")
    assert gw is not None and np.isfinite(gw).all()
    c = float((gw * g("g_warp")).sum() / np.sqrt((gw ** 2).sum() * (g("g_warp") ** 2).sum()))
    assert c > 0.9, c


def test_hit_list_order_and_saturation(orc):
    """K = 3 (not a power of two): DFS order over the implicit heap is k = 1, 2, 0; opaque primitives saturate the ray
    and raysat records the colour of the saturating sample."""
    N, H, W, K, M = 1, 4, 8, 3, 4
    raypos = np.zeros((N, H, W, 3), np.float32); raypos[..., 2] = -2
    raydir = np.zeros((N, H, W, 3), np.float32); raydir[..., 2] = 1
    tminmax = np.zeros((N, H, W, 2), np.float32); tminmax[..., 1] = 4
    primpos = np.zeros((N, K, 3), np.float32)          # three coincident boxes at the origin
    primrot = np.tile(np.eye(3, dtype=np.float32), (N, K, 1, 1))
    primscale = np.full((N, K, 3), 2.0, np.float32)
    tpl = np.zeros((N, K, M, M, M, 4), np.float32)
    for k in range(K):
        tpl[:, k, ..., k] = 1.0       # prim k is pure colour channel k
        tpl[:, k, ..., 3] = 100.0     # very dense
    rgba, raysat, _ = orc.raymarch_fwd(raypos, raydir, 0.05, tminmax, primpos, primrot, primscale, tpl, fadescale=0.0,
                                       fadeexp=1.0)
    # first sample: prim 1 (green) is evaluated first and saturates immediately (alpha = 100*0.05 = 5 >= 1)
    assert np.allclose(rgba[0, 0, 0], [0, 1, 0, 1])
    assert np.allclose(raysat[0, 0, 0], [0, 1, 0])


def test_raydirs_against_reference_formula(orc):
    """extensions/utils/utils.py:127-143 (the reference's PyTorch restatement inside its gradcheck), re-derived."""
    rng = np.random.default_rng(1113)
    N, H, W = 2, 9, 13
    viewpos = rng.standard_normal((N, 3)).astype(np.float32) * 0.2 + np.array([0, 0, -2.5], np.float32)
    q = np.linalg.qr(rng.standard_normal((N, 3, 3)))[0].astype(np.float32)
    focal = np.full((N, 2), 30.0, np.float32)
    princpt = np.array([[W / 2, H / 2]] * N, np.float32)
    px, py = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    pix = np.tile(np.stack([px, py], -1)[None], (N, 1, 1, 1)) + 0.25
    volradius = 1.7
    raypos, raydir, tminmax = orc.raydirs_fwd(viewpos, q, focal, princpt, pix, volradius)
    d = (pix - princpt[:, None, None]) / focal[:, None, None]
    d = np.concatenate([d, np.ones_like(d[..., :1])], -1)
    d = np.einsum("nhwi,nij->nhwj", d, q)          # sum_i d_i * viewrot[i, :]
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    rp = np.broadcast_to((viewpos / volradius)[:, None, None], d.shape)
    t1, t2 = (-1 - rp) / d, (1 - rp) / d
    tmin = np.maximum(np.minimum(t1, t2).max(-1), 0)
    tmax = np.maximum(t1, t2).min(-1)
    assert_close(raydir, d, rtol=1e-5, what="raydir")
    assert_close(raypos, rp, rtol=1e-6, what="raypos")
    assert_close(tminmax, np.stack([tmin, tmax], -1), rtol=1e-4, what="tminmax")
    # pixelcoords=None means integer pixel grid
    a = orc.raydirs_fwd(viewpos, q, focal, princpt, None, volradius, H=H, W=W)
    b = orc.raydirs_fwd(viewpos, q, focal, princpt, pix - 0.25, volradius)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
