"""CPU restatement (plain PyTorch) of the reference's body decoder, `mesh_vae.ConvDecoder` — TEST / BENCH
INFRASTRUCTURE ONLY (BASELINE.json configs[0]: "mesh_vae_example.yml: 1-frame body decoder forward
on CPU"; no kernel is written for this row, it is a CPU-timed baseline).

Follows /root/reference/ca_code/models/mesh_vae.py:440-630 (ConvDecoder.__init__ / forward) with the blocks of
ca_code/nn/blocks.py:232-280 (ConvBlock), :382-434 (UpConvBlockDeep), :731-743 (tile2d) and the weight-normalised
layers of ca_code/nn/layers.py:157-265,276-327,468-472 (LinearWN, Conv2dWN, Conv2dWNUB: w = v * g / ||v||_F, g per
output channel, untied bias [C,H,W] for the "UB" layers).  Parameter names and shapes are the reference's
(`weight_v`, `weight_g`, `bias`), so a reference state_dict loads; tests/golden/make_mesh_vae_golden.py runs the
reference class itself (sliced out of the reference file at run time) against this module on the same weights and
freezes the result in tests/golden/mesh_vae_ref.npz, which tests/test_oracle_mesh_vae.py checks on any box.

What is NOT restated: the seam sampler (ca_code/utils/seams.py, asset-dependent index maps) and `geo_fn.from_uv`
(mesh topology) — both are gathers over the 1024^2 feature maps; callers pass functions for them (the golden script and
the bench use the same bilinear identity resampling on both sides, which has the cost profile of the real gathers).
Config of mesh_vae_example.yml:25-34: uv 1024, init 64, pose 98 dims -> 16 ch, embs 1024 -> 32 ch, face 256,
trunk 2 x (64 -> 32 -> 16 -> 8 -> 4) channels at 128..1024 (groups = 2)."""
import numpy as np
import torch as th
import torch.nn as nn
import torch.nn.functional as F


def _wn(v, g):
    """weight norm of layers.py:157-265 with g_dim = 0, v_dim = None: one norm over the WHOLE direction tensor"""
    return v * (g / th.linalg.vector_norm(v))


class LinearWN(nn.Module):  # layers.py:468
    def __init__(self, n_in, n_out):
        super().__init__()
        self.weight_v = nn.Parameter(th.empty(n_out, n_in))
        self.weight_g = nn.Parameter(th.empty(n_out, 1))
        self.bias = nn.Parameter(th.zeros(n_out))

    def forward(self, x):
        return F.linear(x, _wn(self.weight_v, self.weight_g), self.bias)


class Conv2dWN(nn.Module):  # layers.py:470 (tied bias)
    def __init__(self, cin, cout, kernel_size, padding=0, groups=1):
        super().__init__()
        self.padding, self.groups = padding, groups
        self.weight_v = nn.Parameter(th.empty(cout, cin // groups, kernel_size, kernel_size))
        self.weight_g = nn.Parameter(th.empty(cout, 1, 1, 1))
        self.bias = nn.Parameter(th.zeros(cout))

    def forward(self, x):
        return F.conv2d(x, _wn(self.weight_v, self.weight_g), self.bias, 1, self.padding, 1, self.groups)


class Conv2dWNUB(nn.Module):  # layers.py:276-327,472 (untied bias [C,H,W])
    def __init__(self, cin, cout, height, width, kernel_size, padding=0, groups=1):
        super().__init__()
        self.padding, self.groups = padding, groups
        self.weight_v = nn.Parameter(th.empty(cout, cin // groups, kernel_size, kernel_size))
        self.weight_g = nn.Parameter(th.empty(cout, 1, 1, 1))
        self.bias = nn.Parameter(th.zeros(cout, height, width))

    def forward(self, x):
        return F.conv2d(x, _wn(self.weight_v, self.weight_g), None, 1, self.padding, 1, self.groups) + self.bias[None]


class ConvBlock(nn.Module):  # blocks.py:232-280
    def __init__(self, cin, cout, size, kernel_size=3, padding=1):
        super().__init__()
        self.conv_resize = Conv2dWN(cin, cout, 1)
        self.conv1 = Conv2dWNUB(cin, cin, size, size, kernel_size, padding)
        self.conv2 = Conv2dWNUB(cin, cout, size, size, kernel_size, padding)

    def forward(self, x):
        skip = self.conv_resize(x)
        x = F.leaky_relu(self.conv1(x), 0.2)
        x = F.leaky_relu(self.conv2(x), 0.2)
        return x + skip


class UpConvBlockDeep(nn.Module):  # blocks.py:382-434
    def __init__(self, cin, cout, size, groups=1):
        super().__init__()
        self.size = size
        self.conv_resize = Conv2dWN(cin, cout, 1, groups=groups)
        self.conv1 = Conv2dWNUB(cin, cin, size, size, 3, 1, groups)
        self.conv2 = Conv2dWNUB(cin, cout, size, size, 3, 1, groups)

    def forward(self, x):
        up = F.interpolate(x, size=(self.size, self.size), mode="bilinear", align_corners=True)  # UpsamplingBilinear2d
        skip = self.conv_resize(up)
        x = F.leaky_relu(self.conv1(up), 0.2)
        x = F.leaky_relu(self.conv2(x), 0.2)
        return x + skip


class _Seq(nn.Sequential):
    """nn.Sequential whose LeakyReLU placeholder keeps the reference's child indices ('0', then the activation)"""


class ConvDecoder(nn.Module):
    """mesh_vae.py:439-630.  `masks`: dict of the four asset masks (pose_cond_mask [P,S,S], head/face/body_cond_mask
    [S,S], S = init_uv_size); `resample`: callable standing for seam_sampler.impaint / .resample; `from_uv`: callable
    standing for geo_fn.from_uv."""

    def __init__(self, masks, resample, from_uv, uv_size=1024, init_uv_size=64, n_pose_dims=98, n_pose_enc_channels=16,
                 n_embs=1024, n_embs_enc_channels=32, n_face_embs=256, n_init_channels=64, n_min_channels=4,
                 tex_scale=0.001, verts_scale=0.01):
        super().__init__()
        self.resample, self.from_uv = resample, from_uv
        self.uv_size, self.init_uv_size = uv_size, init_uv_size
        self.tex_scale, self.verts_scale = tex_scale, verts_scale
        self.n_blocks = int(np.log2(uv_size // init_uv_size))
        self.sizes = [init_uv_size * 2 ** s for s in range(self.n_blocks + 1)]
        self.n_channels = [max(n_init_channels // 2 ** b, n_min_channels) for b in range(self.n_blocks + 1)]
        self.local_pose_conv_block = ConvBlock(n_pose_dims, n_pose_enc_channels, init_uv_size, kernel_size=1, padding=0)
        self.embs_fc = _Seq(LinearWN(n_embs, 4 * 4 * 128), nn.LeakyReLU(0.2))
        self.embs_conv_block = nn.Sequential(UpConvBlockDeep(128, 128, 8), UpConvBlockDeep(128, 128, 16),
                                             UpConvBlockDeep(128, 64, 32), UpConvBlockDeep(64, n_embs_enc_channels, 64))
        self.face_embs_fc = _Seq(LinearWN(n_face_embs, 4 * 4 * 32), nn.LeakyReLU(0.2))
        self.face_embs_conv_block = nn.Sequential(UpConvBlockDeep(32, 64, 8), UpConvBlockDeep(64, 64, 16),
                                                  UpConvBlockDeep(64, n_embs_enc_channels, 32))
        g = 2
        self.joint_conv_block = ConvBlock(n_pose_enc_channels + n_embs_enc_channels, n_init_channels, init_uv_size)
        self.conv_blocks = nn.ModuleList([
            UpConvBlockDeep(self.n_channels[b] * g, self.n_channels[b + 1] * g, self.sizes[b + 1], groups=g)
            for b in range(self.n_blocks)])
        self.verts_conv = Conv2dWNUB(self.n_channels[-1], 3, uv_size, uv_size, 3, 1)
        self.tex_conv = Conv2dWNUB(self.n_channels[-1], 3, uv_size, uv_size, 3, 1)
        f32 = lambda a: th.as_tensor(np.asarray(a), dtype=th.float32)
        # mesh_vae.py:560-575
        self.register_buffer("pose_cond_mask", (f32(masks["pose_cond_mask"])[None]
                                                * (1 - f32(masks["head_cond_mask"])[None, None])).to(th.int32))
        self.register_buffer("face_cond_mask", f32(masks["face_cond_mask"])[None, None])
        self.register_buffer("body_cond_mask", f32(masks["body_cond_mask"])[None, None])

    def forward(self, pose, embs, face_embs):  # mesh_vae.py:577-630
        B = pose.shape[0]
        local_pose = pose[:, 6:]
        non_head_mask = (self.body_cond_mask * (1.0 - self.face_cond_mask)).clip(0.0, 1.0)
        S = self.init_uv_size
        pose_masked = local_pose[:, :, None, None].expand(-1, -1, S, S) * self.pose_cond_mask
        pose_conv = self.local_pose_conv_block(pose_masked) * non_head_mask
        embs_conv = self.embs_conv_block(self.embs_fc(embs).reshape(B, 128, 4, 4))
        face_conv = self.face_embs_conv_block(self.face_embs_fc(face_embs).reshape(B, 32, 4, 4))
        embs_conv[:, :, 32:, :32] = (face_conv * self.face_cond_mask[:, :, 32:, :32]
                                     + embs_conv[:, :, 32:, :32] * non_head_mask[:, :, 32:, :32])
        joint = self.joint_conv_block(th.cat([pose_conv, embs_conv], 1))
        x = th.cat([joint, joint], 1)
        for blk in self.conv_blocks:
            x = blk(x)
        x = self.resample(self.resample(self.resample(x)))  # impaint + 2 x resample (mesh_vae.py:607-610)
        verts_features, tex_features = th.split(x, self.n_channels[-1], 1)
        verts_uv_delta_rec = self.verts_conv(verts_features) * self.verts_scale
        return {"geom_delta_rec": self.from_uv(verts_uv_delta_rec), "geom_uv_delta_rec": verts_uv_delta_rec,
                "tex_mean_rec": self.tex_conv(tex_features) * self.tex_scale, "embs_conv": embs_conv,
                "pose_conv": pose_conv}


# ---------------------------------------------------------------- shared by the golden script, the test and the bench
def synthetic_masks(n_pose_dims=98, size=64, seed=11):
    """asset masks of the right shapes (the real ones are per-identity assets of the dataset)"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    face = ((yy >= size // 2) & (xx < size // 2)).astype(np.float32)       # the region ConvDecoder.forward merges into
    head = (face * (rng.random((size, size)) < 0.9)).astype(np.float32)
    body = (rng.random((size, size)) < 0.85).astype(np.float32)
    pose = (rng.random((n_pose_dims, size, size)) < 0.3).astype(np.float32)
    return {"pose_cond_mask": pose, "head_cond_mask": head, "face_cond_mask": face, "body_cond_mask": body}


def identity_resample(x):
    """stand-in for the seam sampler's gathers: one bilinear grid_sample over the map with a half-texel shifted grid"""
    n, _, h, w = x.shape
    ys = th.linspace(-1, 1, h).view(1, h, 1).expand(n, h, w) + 1.0 / h
    xs = th.linspace(-1, 1, w).view(1, 1, w).expand(n, h, w)
    return F.grid_sample(x, th.stack([xs, ys], -1).to(x.dtype), mode="bilinear", padding_mode="border",
                         align_corners=True)


def uv_vertex_gather(n_verts=7306, seed=5):
    """stand-in for geo_fn.from_uv (geom.py: grid_sample of the UV map at the vertices' UV coordinates)"""
    uv = th.from_numpy(np.random.default_rng(seed).random((1, 1, n_verts, 2)).astype(np.float32)) * 2 - 1

    def from_uv(t):
        return F.grid_sample(t, uv.expand(t.shape[0], -1, -1, -1).to(t.dtype), mode="bilinear",
                             align_corners=False)[:, :, 0].permute(0, 2, 1)

    return from_uv


def seeded_fill(module, seed=20240613):
    """deterministic parameters of plausible scale for ANY module with the reference's names (used on both sides)"""
    g = th.Generator().manual_seed(seed)
    with th.no_grad():
        for name, p in sorted(module.named_parameters()):
            if name.endswith("weight_v"):
                p.copy_(th.randn(p.shape, generator=g) * (2.0 / max(1, p[0].numel())) ** 0.5)
            elif name.endswith("weight_g"):
                p.copy_(0.5 + th.rand(p.shape, generator=g))
            else:
                p.copy_(0.05 * th.randn(p.shape, generator=g))
    return module


def seeded_inputs(seed=7, batch=1):
    g = th.Generator().manual_seed(seed)
    return (th.randn(batch, 104, generator=g) * 0.3, th.randn(batch, 1024, generator=g), th.randn(batch, 256, generator=g))
