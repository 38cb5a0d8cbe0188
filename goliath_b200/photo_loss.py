"""Post-render chain and photometric losses of the RGCA train step as four fused kernels (csrc/photo_loss.cu,
SURVEY.md section 8f-2):

  post_render(rgb, ...)          CalV5 colour calibration (ca_code/nn/color_cal.py:211-241) -> background composite
                                 (ca_code/models/rgca.py:226-230) -> LearnableBlur (ca_code/nn/dof_cal.py:44-56)
  photometric_loss(pred, ...)    loss = l1_weight * rgb_l1 + ssim_weight * rgb_ssim with the reference's definitions
                                 (ca_code/loss/__init__.py:391-410 and :479-494 over ca_code/utils/ssim.py)

Both are autograd Functions; the gradient that leaves `post_render` is dL/d(rendered rgb), i.e. the `v_out` of the blend
backward.  No host synchronisation anywhere (the sums stay on the device), so the whole train step stays capturable."""
from typing import Dict, Optional, Tuple

import torch
from torch.autograd import Function

from . import _lib


class _PostRender(Function):
    @staticmethod
    def forward(ctx, rgb, alpha, background, cal_w, cal_b, grey, blur_w):
        rgb = rgb.contiguous()
        _lib.check_input(rgb, "rgb")
        B, C, H, W = rgb.shape
        if C != 3:
            raise RuntimeError("post_render: rgb must be [B,3,H,W]")
        opt = {}
        for name, t, shape, dt in (("alpha", alpha, (B, 1, H, W), torch.float32), ("background", background, (B, 3, H, W), torch.float32),
                                   ("cal_w", cal_w, (B, 3), torch.float32), ("cal_b", cal_b, (B, 3), torch.float32),
                                   ("grey", grey, (B,), torch.int32), ("blur_w", blur_w, (B, 3), torch.float32)):
            if t is not None:
                t = t.contiguous()
                _lib.check_input(t, name, dt)
                if tuple(t.shape) != shape:
                    raise RuntimeError("post_render: %s must have shape %s" % (name, (shape,)))
            opt[name] = t
        if (opt["cal_w"] is None) != (opt["cal_b"] is None):
            raise RuntimeError("post_render: cal_w and cal_b go together")
        if opt["background"] is not None and opt["alpha"] is None:
            raise RuntimeError("post_render: the background composite needs alpha")
        pred = torch.empty_like(rgb)
        with torch.cuda.device(rgb.device):
            _lib.check(_lib.lib().gb_post_render_fwd(
                B, H, W, _lib.ptr(rgb), _lib.ptr(opt["alpha"]), _lib.ptr(opt["background"]), _lib.ptr(opt["cal_w"]),
                _lib.ptr(opt["cal_b"]), _lib.ptr(opt["grey"]), _lib.ptr(opt["blur_w"]), _lib.ptr(pred),
                _lib.stream_ptr(rgb.device)), "post_render_fwd")
        ctx.save_for_backward(rgb, *[opt[k] for k in ("alpha", "background", "cal_w", "cal_b", "grey", "blur_w")])
        return pred

    @staticmethod
    def backward(ctx, g_pred):
        rgb, alpha, background, cal_w, cal_b, grey, blur_w = ctx.saved_tensors
        B, _, H, W = rgb.shape
        g_pred = g_pred.contiguous()
        g_rgb = torch.empty_like(rgb)
        z = lambda t: None if t is None else torch.zeros_like(t)
        g_cw, g_cb, g_bw = z(cal_w), z(cal_b), z(blur_w)
        with torch.cuda.device(rgb.device):
            _lib.check(_lib.lib().gb_post_render_bwd(
                B, H, W, _lib.ptr(rgb), _lib.ptr(alpha), _lib.ptr(background), _lib.ptr(cal_w), _lib.ptr(cal_b), _lib.ptr(grey),
                _lib.ptr(blur_w), _lib.ptr(g_pred), _lib.ptr(g_rgb), _lib.ptr(g_cw), _lib.ptr(g_cb), _lib.ptr(g_bw),
                _lib.stream_ptr(rgb.device)), "post_render_bwd")
        return g_rgb, None, None, g_cw, g_cb, None, g_bw


def post_render(rgb: torch.Tensor, alpha: Optional[torch.Tensor] = None, background: Optional[torch.Tensor] = None,
                cal_w: Optional[torch.Tensor] = None, cal_b: Optional[torch.Tensor] = None,
                grey: Optional[torch.Tensor] = None, blur_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rgb [B,3,H,W] -> blur(cal(rgb) + (1 - alpha) * background).  cal_w / cal_b [B,3]: per-frame calibration rows
    (identity camera: w = 1, b = 0); grey [B] int32 marks grey cameras (out = sum_c img_c w_c + sum_c b_c on all three
    channels); background [B,3,H,W] with alpha [B,1,H,W] (detached upstream, rgca.py:137); blur_weights [B,3] = the
    softmax-ed LearnableBlur weights (identity, 3x3, 7x7).  Any stage whose tensors are None is skipped.
    Gradients: rgb, cal_w, cal_b, blur_weights."""
    return _PostRender.apply(rgb, alpha, background, cal_w, cal_b, grey, blur_weights)


class _SsimL1(Function):
    @staticmethod
    def forward(ctx, pred, target, mask, l1_weight, ssim_weight):
        pred, target, mask = pred.contiguous(), target.contiguous(), mask.contiguous()
        for t, n in ((pred, "pred"), (target, "target"), (mask, "mask")):
            _lib.check_input(t, n)
        B, C, H, W = pred.shape
        if C != 3 or target.shape != pred.shape or mask.shape != (B, 1, H, W):
            raise RuntimeError("photometric_loss: pred / target [B,3,H,W], mask [B,1,H,W]")
        dev = pred.device
        d_mu, d_pp, d_tp = torch.empty_like(pred), torch.empty_like(pred), torch.empty_like(pred)
        sums = torch.zeros(3, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().gb_ssim_l1_fwd(B, H, W, _lib.ptr(pred), _lib.ptr(target), _lib.ptr(mask), _lib.ptr(d_mu),
                                                 _lib.ptr(d_pp), _lib.ptr(d_tp), _lib.ptr(sums), _lib.stream_ptr(dev)),
                       "ssim_l1_fwd")
        l1 = (sums[0] / float(B * 3 * H * W)).float()
        ssim = (sums[1] / sums[2].clamp(min=1.0)).float()
        loss = l1_weight * l1 + ssim_weight * (1.0 - ssim)
        ctx.save_for_backward(pred, target, mask, d_mu, d_pp, d_tp, sums)
        ctx.w = (float(l1_weight), float(ssim_weight))
        ctx.mark_non_differentiable(l1, ssim)
        return loss, l1, ssim

    @staticmethod
    def backward(ctx, g_loss, _g_l1, _g_ssim):
        pred, target, mask, d_mu, d_pp, d_tp, sums = ctx.saved_tensors
        B, _, H, W = pred.shape
        g_pred = torch.empty_like(pred)
        g_loss = g_loss.contiguous().float()
        with torch.cuda.device(pred.device):
            _lib.check(_lib.lib().gb_ssim_l1_bwd(B, H, W, _lib.ptr(pred), _lib.ptr(target), _lib.ptr(mask), _lib.ptr(d_mu),
                                                 _lib.ptr(d_pp), _lib.ptr(d_tp), _lib.ptr(sums), _lib.ptr(g_loss), ctx.w[0],
                                                 ctx.w[1], _lib.ptr(g_pred), _lib.stream_ptr(pred.device)), "ssim_l1_bwd")
        return g_pred, None, None, None, None


def photometric_loss(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor, l1_weight: float = 10.0,
                     ssim_weight: float = 0.2) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """loss = l1_weight * rgb_l1 + ssim_weight * rgb_ssim (config/rgca_example.yml weights by default) with
    rgb_l1 = mean(|(pred - target) * mask|) and rgb_ssim = 1 - sum(ssim_map(target, pred) * mask) / clamp(sum(mask), 1)
    (mask [B,1,H,W] broadcast over the channels, 11x11 window).  Returns (loss, {"rgb_l1", "rgb_ssim"})."""
    loss, l1, ssim = _SsimL1.apply(pred, target, mask, l1_weight, ssim_weight)
    return loss, {"rgb_l1": l1, "rgb_ssim": 1.0 - ssim}
