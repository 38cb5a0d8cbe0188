"""GPU parity of the fused RGCA heads kernel (csrc/rgca_heads.cu, row R2) against vectors produced by the reference's
own source lines (tests/golden/rgca_heads_ref.npz) and, at a larger size, against the pinned torch oracle."""
import os

import numpy as np
import pytest
import torch

from util import assert_close, t2n

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rgca_heads_ref.npz")
LEAVES = ("f_vnocond", "f_vcond", "postex", "tn", "albedo")


def test_heads_vs_reference_lines(cuda):
    from goliath_b200.rgca_heads import gaussian_heads

    g = np.load(GOLD)
    ins = {k[3:]: torch.from_numpy(g[k]).float().to(cuda).requires_grad_(k[3:] in LEAVES) for k in g.files if k.startswith("in_")}
    out = gaussian_heads(ins["f_vnocond"], ins["f_vcond"], ins["postex"], ins["tn"], ins["albedo"], ins["light_sh"], ins["campos"])
    keys = [k[4:] for k in g.files if k.startswith("out_")]
    for k in keys:
        assert_close(t2n(out[k]), g["out_" + k], rtol=1e-4, atol=1e-5 * np.abs(g["out_" + k]).max(), what=k)
    loss = sum((torch.from_numpy(g["w_" + k]).float().to(cuda) * out[k]).sum() for k in keys)
    loss.backward()
    for n in LEAVES:
        ref = g["g_" + n]
        assert_close(t2n(ins[n].grad), ref, rtol=1e-4, atol=2e-5 * np.abs(ref).max(), frac=0.999, what="grad " + n)


def test_heads_vs_oracle_large(cuda):
    from goliath_b200.rgca_heads import gaussian_heads
    from oracle import heads_oracle as ho

    gen = torch.Generator().manual_seed(77)
    B, H, W = 2, 96, 64
    G = H * W
    mk = lambda *s: torch.randn(*s, generator=gen)
    f_vnocond, f_vcond = mk(B, 125, H, W), mk(B, 4, H, W)
    postex = 100 * mk(B, 3, H, W)
    tn = torch.nn.functional.normalize(mk(B, 3, H, W), dim=1)
    albedo = torch.rand(1, G, 3, generator=gen)
    light_sh, campos = mk(B, 3, 81), 1000 * torch.nn.functional.normalize(mk(B, 3), dim=1)
    cpu = [t.double().requires_grad_(i < 5) for i, t in enumerate((f_vnocond, f_vcond, postex, tn, albedo, light_sh, campos))]
    gpu = [t.to(cuda).requires_grad_(i < 5) for i, t in enumerate((f_vnocond, f_vcond, postex, tn, albedo, light_sh, campos))]
    ref = ho.gaussian_heads(*cpu)
    out = gaussian_heads(*gpu)
    ws = {k: torch.randn(ref[k].shape, generator=gen) for k in ref}
    for k in ref:
        r = ref[k].detach().numpy()
        assert_close(t2n(out[k]), r, rtol=1e-4, atol=1e-5 * np.abs(r).max(), frac=0.9999, what=k)
    sum((ws[k].double() * ref[k]).sum() for k in ref).backward()
    sum((ws[k].to(cuda) * out[k]).sum() for k in ref).backward()
    for n, c, gq in zip(LEAVES, cpu, gpu):
        r = c.grad.numpy()
        assert_close(t2n(gq.grad), r, rtol=2e-4, atol=2e-5 * np.abs(r).max(), frac=0.999, what="grad " + n)


def test_shade_and_compose_matches_reference_formula(orc, cuda):
    """rgca.py:557-575: color = clamp(clamp(diff,0) + SG(ref_dirs, sigma, ...) * spec_vis, 0)."""
    from goliath_b200 import synthetic
    from goliath_b200.rgca_heads import gaussian_heads, shade_and_compose

    gen = torch.Generator().manual_seed(5)
    B, H, W = 1, 32, 32
    mk = lambda *s: torch.randn(*s, generator=gen)
    ins = [mk(B, 125, H, W), mk(B, 4, H, W), 100 * mk(B, 3, H, W), torch.nn.functional.normalize(mk(B, 3, H, W), dim=1),
           torch.rand(1, H * W, 3, generator=gen), mk(B, 3, 81), torch.tensor([[0.0, 0.0, 1000.0]])]
    li = synthetic.lights(8)
    heads = gaussian_heads(*[t.to(cuda) for t in ins])
    out = shade_and_compose(heads, li["light_intensity"].to(cuda), li["light_pos"].to(cuda), li["n_lights"].to(cuda))
    n = t2n
    rd = n(torch.nn.functional.normalize(heads["ref_dirs"], dim=-1))
    spec = orc.sg_fwd(rd, n(heads["sigma"]), li["light_intensity"].numpy(), li["light_pos"].numpy(), n(heads["primpos"]),
                      li["n_lights"].numpy(), 0) * n(heads["spec_vis"])
    color = np.maximum(np.maximum(n(heads["diff_color"]), 0) + spec, 0)
    assert_close(n(out["color"]), color, rtol=1e-3, atol=1e-4 * np.abs(color).max(), frac=0.999, what="color")
