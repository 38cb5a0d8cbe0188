"""Shared helpers for the parity tests."""
import numpy as np
import torch

from goliath_b200 import synthetic


def small_scene(G=3000, img_h=96, img_w=80, seed=7, cam=1, spread=1.0, depth_quant=False):
    """A small RGCA-like scene that projects to a few hundred pixels; returns numpy inputs."""
    sc = synthetic.head_gaussians(G, seed=seed)
    c = synthetic.ring_camera(cam, img_h=img_h, img_w=img_w)
    # shrink the focal length so the 100 mm head fills the small image
    f = 0.35 * min(img_h, img_w) / 120.0 * 1000.0 * spread
    c.update(fx=f, fy=f)
    d = {k: v.numpy() for k, v in sc.items()}
    if depth_quant:  # force exact depth ties between different Gaussians (sort stability case)
        d["means3d"] = np.round(d["means3d"] / 8.0).astype(np.float32) * 8.0
    d.update(viewmat=c["viewmat"].numpy(), fx=c["fx"], fy=c["fy"], cx=c["cx"], cy=c["cy"], img_h=img_h, img_w=img_w)
    return d


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def assert_close(a, b, rtol=1e-4, atol=None, frac=1.0, what=""):
    """Element-wise |a-b| <= atol + rtol*|b| for at least `frac` of the elements, and norm-wise rel err <= rtol.
    atol defaults to rtol * 1e-2 * max|b| (fp32 cancellation noise scales with the summands, not the result)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if atol is None:
        atol = rtol * 1e-2 * (np.abs(b).max() if b.size else 0.0)
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    got = ok.mean() if ok.size else 1.0
    assert got >= frac, "%s: only %.6f of elements within tol (need %.6f); max abs diff %.3e" % (
        what, got, frac, np.abs(a - b).max())
    r = rel_err(a, b)
    assert r <= rtol, "%s: norm-wise relative error %.3e > %.1e" % (what, r, rtol)


def t2n(t):
    return t.detach().cpu().numpy()
