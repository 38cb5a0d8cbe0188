"""oracle/heads_oracle.py — torch restatement (any dtype, differentiable by autograd) of row R2: Gaussian heads +
SH diffuse + reflection direction + colour compose of the RGCA PrimDecoder
(/root/reference/ca_code/models/rgca.py:506-546, :572-575; channel map SURVEY.md Appendix C).

TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/rgca_heads_ref.npz, which tests/golden/make_heads_golden.py produces
by executing the reference's own source lines (tests/test_oracle_heads.py)."""
import torch
import torch.nn.functional as F

N_COLOR_SH, N_MONO_SH = 16, 65
N_DIFF = 3 * N_COLOR_SH + N_MONO_SH  # 113


def gaussian_heads(f_vnocond, f_vcond, postex, tn, albedo, light_sh, campos, scale_range=(0.1, 20.0)):
    """f_vnocond [B,125,H,W], f_vcond [B,4,H,W], postex / tn [B,3,H,W] (tn already unit), albedo [1,G,3],
    light_sh [B,3,81], campos [B,3]  ->  dict of [B,G,*] tensors with the reference's key names."""
    B = f_vnocond.shape[0]
    pl = lambda x, c: x.permute(0, 2, 3, 1).reshape(B, -1, c)
    fv = pl(f_vcond, 4)
    primposbase, primnmlbase = pl(postex, 3), pl(tn, 3)
    sh = pl(f_vnocond[:, :N_DIFF], N_DIFF)
    sh_color = sh[..., : 3 * N_COLOR_SH].reshape(B, -1, 3, N_COLOR_SH)
    sh_mono = sh[..., 3 * N_COLOR_SH:].reshape(B, -1, 1, N_MONO_SH)
    diff_shs = torch.cat([sh_color, sh_mono.expand(-1, -1, 3, -1)], -1)
    fg = pl(f_vnocond[:, N_DIFF:N_DIFF + 11], 11)
    primpos = fg[..., 0:3] + primposbase
    primqvec = F.normalize(fg[..., 3:7], dim=-1)
    primscale_preclip = F.softplus(fg[..., 7:10])
    opacity = torch.sigmoid(fg[..., 10:11])
    sigma = (torch.exp(pl(f_vnocond[:, N_DIFF + 11:], 1)[..., 0]) * 0.1).clamp(min=0.01)
    spec_vis = torch.sigmoid(fv[..., :1])
    spec_dnml = fv[..., 1:]
    spec_nml = F.normalize(spec_dnml + primnmlbase, dim=-1)
    diff_color = albedo.expand(B, -1, -1) * (diff_shs * light_sh[:, None]).sum(dim=-1)
    view_local = F.normalize(primpos - campos[:, None], dim=-1, p=2.0)
    ref_dirs = view_local - 2.0 * (view_local * spec_nml).sum(-1, keepdim=True) * spec_nml
    return dict(primpos=primpos, primqvec=primqvec, primscale=primscale_preclip.clamp(*scale_range),
                primscale_preclip=primscale_preclip, opacity=opacity, sigma=sigma, spec_vis=spec_vis, spec_dnml=spec_dnml,
                spec_nml=spec_nml, diff_color=diff_color, ref_dirs=ref_dirs, primnmlbase=primnmlbase)


def compose_color(diff_color, spec_color):
    """rgca.py:572-575: color = clamp(clamp(diff, 0) + spec, 0)."""
    return (diff_color.clamp(min=0.0) + spec_color).clamp(min=0.0)
