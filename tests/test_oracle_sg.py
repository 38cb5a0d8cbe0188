"""CPU: the C restatement of the SG shade (oracle/sg_oracle.c <- extensions/sgutils/sg.cu) against an fp64 torch
restatement; its hand-written backward against autograd where the reference's backward IS the true derivative
(cosine strictly inside (-1,1); the -20 clamp-edge rule is checked separately)."""
import numpy as np
import pytest
import torch

from util import assert_close


def torch_sg(dirs, sig, lv, lp, pp, nl, w_type):
    ld = lp[:, None, :, :] - pp[:, :, None, :]
    ld = ld / ld.norm(dim=-1, keepdim=True)
    cos = (ld * dirs[:, :, None, :]).sum(-1).clamp(-1, 1)
    s = sig[:, :, None]
    if w_type in (0, 1):
        ang = torch.acos(cos)
        w = torch.exp(-0.5 * (ang / s) ** 2)
        if w_type == 0:
            w = w / (s * 3.03352966508)
    else:
        w = torch.exp((cos - 1) / s)
        if w_type == 2:
            w = w / (s * 6.28318530718)
    mask = (torch.arange(lv.shape[1])[None] < nl[:, None]).to(w.dtype)
    return ((w * mask[:, None, :])[..., None] * lv[:, None]).sum(2)


def make(N=2, D=257, L=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    dirs = torch.nn.functional.normalize(torch.randn(N, D, 3, generator=g), dim=-1)
    sig = (0.1 * torch.exp(0.5 * torch.randn(N, D, generator=g))).clamp(min=0.01)
    lv = torch.rand(N, L, 3, generator=g) + 0.5
    lp = 1100.0 * torch.nn.functional.normalize(torch.randn(N, L, 3, generator=g), dim=-1)
    pp = 100.0 * torch.randn(N, D, 3, generator=g)
    nl = torch.tensor([L, max(L - 2, 1)][:N], dtype=torch.int32)
    # aim a few lobes straight at a light so the weights are not all ~0
    dirs[:, ::3] = torch.nn.functional.normalize(lp[:, :1] - pp[:, ::3] + 30 * torch.randn(N, len(range(0, D, 3)), 3, generator=g), dim=-1)
    return dirs, sig, lv, lp, pp, nl


@pytest.mark.parametrize("w_type", [0, 1, 2, 3])
def test_sg_forward_backward(orc, w_type):
    dirs, sig, lv, lp, pp, nl = make()
    if w_type >= 2:
        sig = sig * 0 + 0.3  # exp((cos-1)/s) lobes are wide; keep weights in range
    out = orc.sg_fwd(dirs.numpy(), sig.numpy(), lv.numpy(), lp.numpy(), pp.numpy(), nl.numpy(), w_type)
    d = lambda t: t.double()
    dd, ss, ll = d(dirs).requires_grad_(), d(sig).requires_grad_(), d(lv).requires_grad_()
    ref = torch_sg(dd, ss, ll, d(lp), d(pp), nl, w_type)
    assert float(ref.abs().max()) > 1e-3
    # w_type 0/1 are ill-conditioned for narrow lobes (angle = acos(cos) near 1), so element tolerance is loose
    assert_close(out, ref.detach().numpy(), rtol=2e-3, atol=1e-6, what="integral")
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (ref * g.double()).sum().backward()
    gd, gs, gl = orc.sg_bwd(dirs.numpy(), sig.numpy(), lv.numpy(), lp.numpy(), pp.numpy(), nl.numpy(), g.numpy(),
                            w_type)
    assert_close(gl, ll.grad.numpy(), rtol=2e-3, atol=1e-5, what="grad_light_values")
    assert_close(gs, ss.grad.numpy(), rtol=5e-3, atol=1e-4 * float(ss.grad.abs().max()), what="grad_sigmas")
    assert_close(gd, dd.grad.numpy(), rtol=5e-3, atol=1e-4 * float(dd.grad.abs().max()), what="grad_dirs")


def test_sg_padding_lights_are_skipped(orc):
    dirs, sig, lv, lp, pp, nl = make(N=1, L=6)
    nl[:] = 3
    a = orc.sg_fwd(dirs.numpy(), sig.numpy(), lv.numpy(), lp.numpy(), pp.numpy(), nl.numpy(), 0)
    b = orc.sg_fwd(dirs.numpy(), sig.numpy(), lv[:, :3].numpy(), lp[:, :3].numpy(), pp.numpy(), nl.numpy(), 0)
    assert np.array_equal(a, b)


def test_sg_clamp_edge_derivative_is_minus_20(orc):
    """sg.cu:129: when |cos| >= 1 the reference substitutes d acos/d cos = -20."""
    lp = np.array([[[0.0, 0.0, 10.0]]], np.float32)
    pp = np.zeros((1, 1, 3), np.float32)
    dirs = np.array([[[0.0, 0.0, 1.0]]], np.float32)   # cos == 1 exactly
    sig = np.array([[0.5]], np.float32)
    lv = np.ones((1, 1, 3), np.float32)
    nl = np.array([1], np.int32)
    gd, gs, gl = orc.sg_bwd(dirs, sig, lv, lp, pp, nl, np.ones((1, 1, 3), np.float32), 0)
    # angle = 0 -> dL_angle = 0 -> grad_dirs = 0 * -20 = 0, grad finite (no NaN from 1/sqrt(0))
    assert np.all(np.isfinite(gd)) and np.all(gd == 0)
    dirs2 = np.array([[[0.0, 0.0, 1.0000001]]], np.float32) * 1.5  # cos > 1: clamped angle 0
    gd2, _, _ = orc.sg_bwd(dirs2, sig, lv, lp, pp, nl, np.ones((1, 1, 3), np.float32), 0)
    assert np.all(np.isfinite(gd2))


def test_regrouped_pair_arithmetic_of_the_fused_shade_is_within_the_bar():
    """csrc/sg_shade.cu evaluates a (Gaussian, light) pair in the fused shade_compose pass as
    cos = (l.dir) rsqrt(l.l);  angle = sqrt(1-|c|) P7(|c|) (Abramowitz & Stegun 4.4.46), reflected for c < 0;
    weight = ex2(-0.5 log2e (angle / sigma)^2) / (sigma K) with the normalisation applied to the light sum.
    The same arithmetic in numpy fp32 against the fp64 restatement of the reference formula: forward values and the
    closed-form gradients w.r.t. sigma and the (unit) lobe direction, over sharp (sigma 0.01) to broad lobes and cosines
    up to the clamp edges."""
    f = np.float32
    dirs, sig, lv, lp, pp, nl = make(N=2, D=513, L=7, seed=5)
    ref = torch_sg(dirs.double(), sig.double(), lv.double(), lp.double(), pp.double(), nl, 0).numpy()
    d, s_, v, P, Q = (t.numpy().astype(f) for t in (dirs, sig, lv, lp, pp))
    co = [f(c) for c in (-0.0012624911, 0.0066700901, -0.0170881256, 0.0308918810, -0.0501743046, 0.0889789874,
                         -0.2145988016, 1.5707963050)]
    out = np.zeros_like(ref, dtype=f)
    for n in range(2):
        acc = np.zeros((d.shape[1], 3), f)
        for l in range(int(nl[n])):
            lvec = (P[n, l][None] - Q[n]).astype(f)
            rinv = (f(1) / np.sqrt((lvec * lvec).sum(-1, dtype=f))).astype(f)
            c = np.clip(((lvec * d[n]).sum(-1, dtype=f) * rinv).astype(f), f(-1), f(1))
            a = np.abs(c)
            p = np.full_like(a, co[0])
            for k in co[1:]:
                p = (p * a + k).astype(f)
            r = (np.sqrt((f(1) - a).astype(f)) * p).astype(f)
            ang = np.where(c < 0, f(3.14159265358979) - r, r).astype(f)
            t = (ang / s_[n]).astype(f)
            w = np.exp2((f(-0.72134752044448) * (t * t)).astype(f)).astype(f)
            acc += w[:, None] * v[n, l][None]
        out[n] = acc * (f(0.32964899322) / s_[n])[:, None]
    # the same pairs in the REFERENCE's operation order, also fp32 (sg.cu:27-76): for sharp lobes the fp32 cosine itself
    # limits both (d angle = d cos / sin angle), so the bar for the regrouped arithmetic is the reference order's own
    # distance to fp64, not a fixed 1e-4
    ref32 = np.zeros_like(out)
    for n in range(2):
        acc = np.zeros((d.shape[1], 3), f)
        for l in range(int(nl[n])):
            lvec = (P[n, l][None] - Q[n]).astype(f)
            ln = np.sqrt((lvec * lvec).sum(-1, dtype=f)).astype(f)
            lh = (lvec / ln[:, None]).astype(f)
            c = np.clip((lh * d[n]).sum(-1, dtype=f), f(-1), f(1))
            ang = np.arccos(c).astype(f)
            w = (np.exp((f(-0.5) * ((ang / s_[n]) ** 2).astype(f)).astype(f)) / (s_[n] * f(3.03352966508))).astype(f)
            acc += w[:, None] * v[n, l][None]
        ref32[n] = acc
    scale = float(np.abs(ref).max())
    err_new, err_ref = np.abs(out - ref) / scale, np.abs(ref32 - ref) / scale
    assert float(err_new.max()) <= 2.0 * float(err_ref.max()) + 1e-6, (float(err_new.max()), float(err_ref.max()))
    assert float(np.sqrt((err_new ** 2).mean())) <= 2.0 * float(np.sqrt((err_ref ** 2).mean())) + 1e-7
    assert_close(out, ref, rtol=1e-4, atol=1e-5 * scale, frac=0.995, what="regrouped SG forward")
