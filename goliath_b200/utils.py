"""Host-side mirror of the reference wrapper extensions/utils/utils.py:26-54 (`ComputeRaydirs`, `compute_raydirs`)."""
import torch
from torch.autograd import Function

from . import utilslib


class ComputeRaydirs(Function):
    @staticmethod
    def forward(self, viewpos, viewrot, focal, princpt, pixelcoords, volradius):
        for tensor in [viewpos, viewrot, focal, princpt, pixelcoords]:
            assert isinstance(tensor, tuple) or tensor.is_contiguous()
        N = viewpos.size(0)
        if isinstance(pixelcoords, tuple):
            W, H = pixelcoords
            pixelcoords = None
        else:
            H = pixelcoords.size(1)
            W = pixelcoords.size(2)
        raypos = torch.empty((N, H, W, 3), device=viewpos.device)
        raydirs = torch.empty((N, H, W, 3), device=viewpos.device)
        tminmax = torch.empty((N, H, W, 2), device=viewpos.device)
        utilslib.compute_raydirs_forward(viewpos, viewrot, focal, princpt, pixelcoords, W, H, volradius, raypos,
                                         raydirs, tminmax)
        return raypos, raydirs, tminmax

    @staticmethod
    def backward(self, grad_raypos, grad_raydirs, grad_tminmax):
        return None, None, None, None, None, None  # extensions/utils/utils.py:48-50


def compute_raydirs(viewpos, viewrot, focal, princpt, pixelcoords, volradius):
    raypos, raydirs, tminmax = ComputeRaydirs.apply(viewpos, viewrot, focal, princpt, pixelcoords, volradius)
    return raypos, raydirs, tminmax
