"""Python surface of the ray generator with the names of the reference wrapper (extensions/utils/utils.py:
`ComputeRaydirs` :26-50, `compute_raydirs` :52-54), on top of csrc/raydirs.cu (through `utilslib`).

pixelcoords is either a [N,H,W,2] tensor of pixel positions or a (W, H) tuple meaning the integer pixel grid.
Like the reference, the op provides no gradients (its CUDA backward is an empty stub, utils_kernel.cu:53-94)."""
import torch
from torch.autograd import Function

from . import utilslib


def _grid_size(pixelcoords):
    if isinstance(pixelcoords, tuple):
        width, height = pixelcoords
        return int(height), int(width), None
    assert pixelcoords.is_contiguous(), "pixelcoords must be contiguous"
    return pixelcoords.size(1), pixelcoords.size(2), pixelcoords


class ComputeRaydirs(Function):
    @staticmethod
    def forward(ctx, viewpos, viewrot, focal, princpt, pixelcoords, volradius):
        for name, t in (("viewpos", viewpos), ("viewrot", viewrot), ("focal", focal), ("princpt", princpt)):
            assert t.is_contiguous(), name + " must be contiguous"
        H, W, coords = _grid_size(pixelcoords)
        n_views, dev = viewpos.size(0), viewpos.device
        raypos, raydirs = (torch.empty((n_views, H, W, 3), device=dev) for _ in range(2))
        tminmax = torch.empty((n_views, H, W, 2), device=dev)
        utilslib.compute_raydirs_forward(viewpos, viewrot, focal, princpt, coords, W, H, volradius, raypos, raydirs, tminmax)
        return raypos, raydirs, tminmax

    @staticmethod
    def backward(ctx, *unused_grads):
        return (None,) * 6


def compute_raydirs(viewpos, viewrot, focal, princpt, pixelcoords, volradius):
    """-> (raypos [N,H,W,3] = viewpos / volradius, raydir [N,H,W,3] unit, tminmax [N,H,W,2] vs the cube [-1,1]^3)."""
    return ComputeRaydirs.apply(viewpos, viewrot, focal, princpt, pixelcoords, volradius)
