"""RGCA `PrimDecoder` (ca_code/models/rgca.py:372-620) assembled from the sm_100a pieces: the two 7-layer
weight-normalised untied-bias deconv towers (rows R1, `goliath_b200.nn`), the fused Gaussian heads + SH diffuse
(row R2, `goliath_b200.rgca_heads`) and the SG specular shade (row R3, `goliath_b200.sgutils`).

Same constructor / forward arguments, the same `preds` keys (rgca.py:574-588) and the same parameter names and shapes
as the reference module, so its checkpoints load: `viewmod.0.*`, `encmod.0.*`, `vnocond_mod.{0,2,..,12}.*`,
`vcond_mod.{0,2,..,12}.*`, `albedo`.  `slabsize` (1024 upstream, hard-coded at rgca.py:385) is a parameter here so the
module can be exercised at small sizes.  The environment-map branch (`preconv_envmap`, rgca.py:548-556) and the
training-only random back-light regulariser (rgca.py:590-618) stay in PyTorch exactly as upstream."""
from typing import Optional

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from . import nn as gnn
from .rgca_heads import PRIMSCALE_RANGE, gaussian_heads, shade_and_compose


def _tower(c_in, c_out, slab):
    plan = [(c_in, 256), (256, 128), (128, 128), (128, 64), (64, 32), (32, 16), (16, c_out)]
    layers, size = [], slab >> 6
    for i, (a, b) in enumerate(plan):
        act = nn.LeakyReLU(0.2, inplace=True) if i < len(plan) - 1 else None
        layers += gnn.make_conv_trans(a, b, 4, 2, 1, "wn", act, ub=(size, size))
        size *= 2
    return nn.Sequential(*layers)


class PrimDecoder(nn.Module):
    """A decoder for relightable Gaussians (same role and interface as rgca.PrimDecoder)."""

    def __init__(self, n_embs, geo_fn, color_mean: th.Tensor, n_diff_sh: int = 8, n_color_sh: int = 3, slabsize: int = 1024):
        super().__init__()
        assert (n_diff_sh, n_color_sh) == (8, 3), "the fused heads kernel implements the reference's SH split (8, 3)"
        assert slabsize >= 128 and slabsize & (slabsize - 1) == 0
        self.slabsize, self.n_splats, self.n_embs, self.geo_fn = slabsize, slabsize ** 2, n_embs, geo_fn
        self.base = slabsize >> 7  # 8 upstream
        self.viewmod = nn.Sequential(*gnn.make_linear(3, 8, "wn", nn.LeakyReLU(0.2, inplace=True)))
        self.encmod = nn.Sequential(*gnn.make_linear(n_embs, 256 * self.base * self.base, "wn", nn.LeakyReLU(0.2, inplace=True)))
        self.n_diff_coeffs = 113
        self.vnocond_mod = _tower(256, self.n_diff_coeffs + 12, slabsize)
        self.vcond_mod = _tower(256 + 8, 4, slabsize)
        rgb = color_mean / 255.0
        self.albedo = nn.Parameter((2.0 * rgb / 2.2974).permute(1, 2, 0).reshape(1, -1, 3))
        # inference (no_grad) runs the towers on the tensor cores (tcgen05 + TMA, csrc/deconv_tc.cu); training uses
        # the SIMT kernels that also provide the backward (csrc/deconv_wnub.cu)
        self.use_tensor_cores = True

    def _run_tower(self, tower, x):
        if self.use_tensor_cores and not th.is_grad_enabled():
            return gnn.tower_forward_tc(tower, x)
        return tower(x)

    def forward(self, embs, geom, headrel_campos, light_intensity, headrel_light_pos, headrel_light_sh, n_lights,
                preconv_envmap: Optional[th.Tensor] = None, lightrot: Optional[th.Tensor] = None):
        if preconv_envmap is not None:
            raise NotImplementedError("environment-map specular branch (rgca.py:548-556): use the reference's PyTorch code")
        postex = self.geo_fn.to_uv(geom)
        tn = F.normalize(self.geo_fn.to_uv(self.geo_fn.vn(geom)), dim=1)
        x = self.encmod(embs).view(-1, 256, self.base, self.base)
        f_vnocond = self._run_tower(self.vnocond_mod, x)
        view = self.viewmod(F.normalize(headrel_campos, dim=1))[:, :, None, None].expand(-1, -1, self.base, self.base)
        f_vcond = self._run_tower(self.vcond_mod, th.cat([x, view], dim=1))
        heads = gaussian_heads(f_vnocond, f_vcond, postex, tn, self.albedo, headrel_light_sh, headrel_campos, PRIMSCALE_RANGE)
        return shade_and_compose(heads, light_intensity, headrel_light_pos, n_lights)
