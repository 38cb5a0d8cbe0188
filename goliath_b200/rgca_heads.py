"""Row R2: the Gaussian heads + SH diffuse + reflection direction block of the RGCA PrimDecoder
(ca_code/models/rgca.py:506-546) and the colour compose (:557-575) as ONE fused sm_100a kernel per direction
(csrc/rgca_heads.cu) instead of ~60 eager element-wise kernels and a [B,G,3,81] temporary.

`gaussian_heads(...)` returns a dict with the reference's `preds` key names and shapes (rgca.py:574-588);
`shade_and_compose(...)` adds the SG specular term and the final colour exactly as rgca.py:557-575 does."""
from typing import Dict

import torch
from torch.autograd import Function

from . import _lib
from .sgutils import evaluate_gaussian

PRIMSCALE_RANGE = (0.1, 20.0)  # ca_code/models/rgca.py:47
_OUT = ("primpos", "primqvec", "primscale", "primscale_preclip", "opacity", "sigma", "spec_vis", "spec_dnml", "spec_nml",
        "diff_color", "ref_dirs", "primnmlbase")


class _GaussianHeads(Function):
    @staticmethod
    def forward(ctx, f_vnocond, f_vcond, postex, tn, albedo, light_sh, campos, scale_lo, scale_hi):
        ins = [t.contiguous() for t in (f_vnocond, f_vcond, postex, tn, albedo, light_sh, campos)]
        names = ("f_vnocond", "f_vcond", "postex", "tn", "albedo", "light_sh", "campos")
        for t, n in zip(ins, names):
            _lib.check_input(t, n)
        f_vnocond, f_vcond, postex, tn, albedo, light_sh, campos = ins
        B, C, H, W = f_vnocond.shape
        G = H * W
        if C != 125 or f_vcond.shape[1] != 4 or light_sh.shape[1:] != (3, 81) or albedo.numel() != G * 3:
            raise RuntimeError("gaussian_heads expects f_vnocond [B,125,H,W], f_vcond [B,4,H,W], light_sh [B,3,81], "
                               "albedo [1,H*W,3]")
        dev = f_vnocond.device
        e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        outs = dict(primpos=e(B, G, 3), primqvec=e(B, G, 4), primscale=e(B, G, 3), primscale_preclip=e(B, G, 3),
                    opacity=e(B, G, 1), sigma=e(B, G), spec_vis=e(B, G, 1), spec_dnml=e(B, G, 3), spec_nml=e(B, G, 3),
                    diff_color=e(B, G, 3), ref_dirs=e(B, G, 3), primnmlbase=e(B, G, 3))
        shsum = e(B, G, 3)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gb_rgca_heads_fwd(
                B, G, *[_lib.ptr(t) for t in ins], float(scale_lo), float(scale_hi),
                *[_lib.ptr(outs[k]) for k in _OUT], _lib.ptr(shsum), _lib.stream_ptr(dev)), "rgca_heads_fwd")
        ctx.save_for_backward(*ins, shsum)
        ctx.scale = (float(scale_lo), float(scale_hi))
        ctx.shape = (B, C, H, W)
        ctx.set_materialize_grads(False)
        return tuple(outs[k] for k in _OUT)

    @staticmethod
    def backward(ctx, *gouts):
        f_vnocond, f_vcond, postex, tn, albedo, light_sh, campos, shsum = ctx.saved_tensors
        B, C, H, W = ctx.shape
        G = H * W
        dev = f_vnocond.device
        gouts = [None if g is None else g.contiguous() for g in gouts]
        g_fn = torch.empty_like(f_vnocond)
        g_fv = torch.empty_like(f_vcond)
        g_pt = torch.empty_like(postex)
        g_tn = torch.empty_like(tn)
        g_al = torch.empty(B, G, 3, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gb_rgca_heads_bwd(
                B, G, _lib.ptr(f_vnocond), _lib.ptr(f_vcond), _lib.ptr(postex), _lib.ptr(tn), _lib.ptr(albedo),
                _lib.ptr(light_sh), _lib.ptr(campos), ctx.scale[0], ctx.scale[1], _lib.ptr(shsum),
                *[_lib.ptr(g) for g in gouts], _lib.ptr(g_fn), _lib.ptr(g_fv), _lib.ptr(g_pt), _lib.ptr(g_tn),
                _lib.ptr(g_al), _lib.stream_ptr(dev)), "rgca_heads_bwd")
        g_albedo = g_al.sum(0, keepdim=True).view_as(albedo) if ctx.needs_input_grad[4] else None
        return g_fn, g_fv, g_pt, g_tn, g_albedo, None, None, None, None


def gaussian_heads(f_vnocond, f_vcond, postex, tn, albedo, headrel_light_sh, headrel_campos,
                   primscale_range=PRIMSCALE_RANGE) -> Dict[str, torch.Tensor]:
    """f_vnocond [B,125,H,W] and f_vcond [B,4,H,W] are the decoder towers' raw outputs (rgca.py:495,503), postex the UV
    position map, tn the UNIT normal map (rgca.py:483-491), albedo the [1,H*W,3] parameter, headrel_light_sh [B,3,81],
    headrel_campos [B,3].  Returns the reference's per-Gaussian tensors plus `ref_dirs`."""
    outs = _GaussianHeads.apply(f_vnocond, f_vcond, postex, tn, albedo, headrel_light_sh, headrel_campos,
                                primscale_range[0], primscale_range[1])
    return dict(zip(_OUT, outs))


def shade_and_compose(heads: Dict[str, torch.Tensor], light_intensity, headrel_light_pos, n_lights) -> Dict[str, torch.Tensor]:
    """rgca.py:557-575 (point-light branch): SG specular via evaluate_gaussian, times spec_vis, plus clamped diffuse."""
    spec_color = evaluate_gaussian(heads["ref_dirs"].contiguous(), heads["sigma"].contiguous(), light_intensity.contiguous(),
                                   headrel_light_pos.contiguous(), heads["primpos"].contiguous(), n_lights.int(),
                                   w_type=0) * heads["spec_vis"]
    color = heads["diff_color"].clamp(min=0.0) + spec_color
    out = dict(heads)
    out.update(spec_color=spec_color, color=color.clamp(min=0.0))
    return out
