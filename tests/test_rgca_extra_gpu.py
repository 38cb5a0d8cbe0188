"""The two remaining branches of rgca.PrimDecoder.forward against vectors produced by the reference's own code
(tests/golden/rgca_extra_ref.npz, made by tests/golden/make_rgca_extra_golden.py): the environment-map specular
branch (rgca.py:548-556 -> csrc/envmap_spec.cu) and the training-mode random back light (rgca.py:590-618 -> second
light-SH table of csrc/rgca_heads.cu), plus key parity of the assembled decoder in train mode."""
import os

import numpy as np
import pytest
import torch

from util import assert_close, t2n

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rgca_extra_ref.npz")
REF_KEYS = {"color", "opacity", "primpos", "primqvec", "primscale", "primscale_preclip", "sigma", "spec_vis", "spec_nml",
            "spec_dnml", "diff_color", "spec_color", "primnmlbase"}  # rgca.py:574-588


def _t(z, k, dev, grad=False):
    t = torch.from_numpy(z[k]).float().to(dev).contiguous()
    return t.requires_grad_() if grad else t


def test_envmap_specular_matches_reference_branch(cuda):
    from goliath_b200.envmap_spec import envmap_specular

    z = np.load(GOLD)
    dirs, vis = _t(z, "env_ref_dirs", cuda, True), _t(z, "env_spec_vis", cuda, True)
    levels = [_t(z, "env_level%d" % i, cuda) for i in range(4)]
    spec = envmap_specular(levels, dirs, _t(z, "env_sigma", cuda), vis, _t(z, "env_lightrot", cuda))
    want = z["env_spec"]
    assert (want >= 0.999 * z["env_spec_vis"]).mean() > 0.02, "clamp(max=1) must be active somewhere"
    assert_close(t2n(spec), want, rtol=1e-4, atol=1e-5 * float(np.abs(want).max()), what="spec_color")
    (spec * _t(z, "env_w", cuda)).sum().backward()
    assert_close(t2n(vis.grad), z["env_g_spec_vis"], rtol=1e-4, atol=1e-5 * float(np.abs(z["env_g_spec_vis"]).max()),
                 what="grad spec_vis")
    gd, want_gd = t2n(dirs.grad), z["env_g_ref_dirs"]
    ok = np.isfinite(want_gd).all(-1) & (np.abs(z["env_ref_dirs"][..., 1]) < 0.999)  # d acos / dy is singular at the poles
    assert ok.mean() > 0.95
    # a texel-boundary crossing between fp32 and fp64 coordinates changes the bilinear slope of a few samples
    assert_close(gd[ok], want_gd[ok], rtol=2e-3, atol=1e-3 * float(np.abs(want_gd[ok]).max()), frac=0.995, what="grad ref_dirs")
    assert np.isfinite(gd).all()


def test_envmap_single_level_and_errors(cuda):
    from goliath_b200.envmap_spec import envmap_specular

    z = np.load(GOLD)
    dirs, vis, sig, rot = (_t(z, "env_ref_dirs", cuda), _t(z, "env_spec_vis", cuda), _t(z, "env_sigma", cuda),
                           _t(z, "env_lightrot", cuda))
    lv0 = _t(z, "env_level0", cuda)
    one = envmap_specular([lv0], dirs, sig, vis, rot)
    # with a single level the reference returns grid_sample of that level; same as a 2-level pyramid at level 0
    two = envmap_specular([lv0, _t(z, "env_level1", cuda)], dirs, torch.zeros_like(sig), vis, rot)
    assert torch.allclose(one, two, atol=1e-6)
    with pytest.raises(RuntimeError):
        envmap_specular([lv0.cpu()], dirs, sig, vis, rot)
    with pytest.raises(RuntimeError):
        envmap_specular([lv0[:, :2]], dirs, sig, vis, rot)


def test_random_backlight_matches_reference_branch(cuda):
    from goliath_b200.rgca_heads import gaussian_heads
    from goliath_b200.sh import dir2sh

    z = np.load(GOLD)
    B, _, G = z["tr_planes"].shape
    H, W = 7, 11
    assert H * W == G
    light_dir = torch.from_numpy(z["tr_light_dir"]).float()
    light_sh = (dir2sh(8, light_dir)[:, :, None] * torch.ones(B, 1, 3, 1)).sum(1)          # rgca.py:610-613
    assert_close(light_sh.numpy(), z["tr_light_sh"], rtol=1e-5, what="random light SH")
    gen = torch.Generator().manual_seed(5)
    planes = _t(z, "tr_planes", cuda, True)
    rest = torch.randn(B, 12, G, generator=gen).to(cuda)
    f_vnocond = torch.cat([planes, rest], 1).view(B, 125, H, W)
    f_vcond = torch.randn(B, 4, H, W, generator=gen).to(cuda)
    postex = torch.randn(B, 3, H, W, generator=gen).to(cuda) * 50
    tn = torch.nn.functional.normalize(torch.randn(B, 3, H, W, generator=gen), dim=1).to(cuda)
    albedo = torch.rand(1, G, 3, generator=gen).to(cuda)
    sh_main = torch.randn(B, 3, 81, generator=gen).to(cuda)
    campos = torch.tensor([[0.0, 0.0, 1000.0]] * B, device=cuda)
    outs = gaussian_heads(f_vnocond, f_vcond, postex, tn, albedo, sh_main, campos, rand_light_sh=light_sh.to(cuda))
    ref_no = gaussian_heads(f_vnocond.detach(), f_vcond, postex, tn, albedo, sh_main, campos)
    assert "diff_color_rand" in outs and "diff_color_rand" not in ref_no
    for k in ref_no:  # the second table must not disturb the regular outputs
        assert torch.equal(outs[k], ref_no[k]), k
    color_rand = outs["diff_color_rand"].clamp(min=0.0)
    assert_close(t2n(color_rand), z["tr_color_rand"], rtol=1e-4, atol=1e-5 * float(np.abs(z["tr_color_rand"]).max()),
                 what="color_rand")
    (color_rand * _t(z, "tr_w", cuda)).sum().backward()
    assert_close(t2n(planes.grad), z["tr_g_planes"], rtol=1e-4, atol=1e-5 * float(np.abs(z["tr_g_planes"]).max()),
                 what="grad diffuse planes via color_rand")
    cos_w = (light_dir.to(cuda) * _t(z, "tr_spec_nml", cuda)).sum(-1, keepdim=True)
    assert_close(t2n(cos_w), z["tr_cos_weight"], rtol=1e-5, atol=1e-6, what="cos_weight")


def test_prim_decoder_train_and_envmap_modes(cuda):
    """Key / shape parity of the assembled decoder (small slab): train mode emits cos_weight and color_rand and
    backpropagates through both; the environment-map branch runs through the same forward."""
    from goliath_b200.rgca import PrimDecoder

    S = 128

    class Geo:
        def __init__(self, nml): self.nml = nml
        def to_uv(self, x): return x
        def vn(self, g): return self.nml

    gen = torch.Generator().manual_seed(9)
    B = 2
    pos = (torch.randn(B, 3, S, S, generator=gen) * 60).to(cuda)
    nml = torch.nn.functional.normalize(torch.randn(B, 3, S, S, generator=gen), dim=1).to(cuda)
    dec = PrimDecoder(16, Geo(nml), 255 * torch.rand(3, S, S, generator=gen), slabsize=S).to(cuda)
    # glorot (rgca.py:458-460): tied sub-pixel phases, zero biases
    last = dec._last_deconv(dec.vnocond_mod)
    assert torch.equal(last.weight_v[:, :, 0::2, 0::2], last.weight_v[:, :, 1::2, 1::2]) and float(last.bias.abs().max()) == 0.0
    embs = torch.randn(B, 16, generator=gen).to(cuda)
    campos = torch.tensor([[0.0, 0.0, 1000.0], [300.0, 0.0, 900.0]], device=cuda)
    L = 4
    lint = torch.rand(B, L, 3, generator=gen).to(cuda)
    lpos = (torch.randn(B, L, 3, generator=gen) * 1000).to(cuda)
    lsh = torch.randn(B, 3, 81, generator=gen).to(cuda)
    nl = torch.full((B,), L, dtype=torch.int32, device=cuda)
    dec.train()
    preds = dec(embs, pos, campos, lint, lpos, lsh, nl)
    assert REF_KEYS | {"cos_weight", "color_rand"} <= set(preds)
    G = S * S
    assert preds["cos_weight"].shape == (B, G, 1) and preds["color_rand"].shape == (B, G, 3)
    assert not preds["cos_weight"].requires_grad and preds["color_rand"].requires_grad
    (preds["color_rand"].sum() + preds["color"].sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for n, p in dec.named_parameters() if "vnocond" in n)
    dec.eval()
    with torch.no_grad():
        preds_e = dec(embs, pos, campos, lint, lpos, lsh, nl)
        assert "cos_weight" not in preds_e and REF_KEYS <= set(preds_e)
        env = [torch.rand(B, 3, 32 >> i, 64 >> i, generator=gen).to(cuda) for i in range(4)]
        rot = torch.eye(3, device=cuda)[None].repeat(B, 1, 1)
        preds_v = dec(embs, pos, campos, lint, lpos, lsh, nl, preconv_envmap=env, lightrot=rot)
        assert REF_KEYS <= set(preds_v) and torch.isfinite(preds_v["color"]).all()
        assert torch.equal(preds_v["diff_color"], preds_e["diff_color"])
