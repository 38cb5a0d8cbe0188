"""Seeded synthetic inputs for the RGCA splat / shade path (recipe: SURVEY.md §8d).

Generated on the CPU with a fixed torch.Generator so every rank / device / test sees identical bytes.
"""
import math
from typing import Dict

import torch

SEED = 20240613


def ring_camera(k: int, n_ring: int = 16, radius: float = 1000.0, img_h: int = 1024, img_w: int = 667) -> Dict:
    """Camera k on a ring around the origin, looking at it.  Returns viewmat [3,4] (world->camera, row-major)
    and pinhole intrinsics (fx = fy = 3300 * W / 1334)."""
    yaw = 2.0 * math.pi * k / n_ring
    cam = torch.tensor([radius * math.sin(yaw), 0.0, radius * math.cos(yaw)], dtype=torch.float64)
    z = -cam / cam.norm()                       # camera looks at the origin
    up = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)   # image y points down
    x = torch.linalg.cross(up, z)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    R = torch.stack([x, y, z])                  # rows = camera axes in world space
    t = -R @ cam
    viewmat = torch.cat([R, t[:, None]], 1).to(torch.float32).contiguous()
    f = 3300.0 * (img_w / 1334.0)
    return dict(viewmat=viewmat, fx=f, fy=f, cx=img_w / 2.0, cy=img_h / 2.0, img_h=img_h, img_w=img_w,
                campos=cam.to(torch.float32))


def head_gaussians(G: int = 300_000, seed: int = SEED) -> Dict[str, torch.Tensor]:
    """Ellipsoid-shell Gaussians: radii (90,120,100) mm * (1 + 0.05 N), unit quats (w,x,y,z),
    scales = clamp(softplus(N(0.5,0.7)), 0.1, 20) mm, opacity = sigmoid(N(1,2)), colors ~ U(0,1)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(G, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    rad = torch.tensor([90.0, 120.0, 100.0]) * (1.0 + 0.05 * torch.randn(G, 1, generator=g))
    means = (d * rad).contiguous()
    q = torch.randn(G, 4, generator=g)
    quats = (q / q.norm(dim=-1, keepdim=True)).contiguous()
    scales = torch.nn.functional.softplus(0.5 + 0.7 * torch.randn(G, 3, generator=g)).clamp(0.1, 20.0).contiguous()
    opacity = torch.sigmoid(1.0 + 2.0 * torch.randn(G, 1, generator=g)).contiguous()
    colors = torch.rand(G, 3, generator=g).contiguous()
    return dict(means3d=means, quats=quats, scales=scales, opacity=opacity, colors=colors)


def lights(L: int = 32, radius: float = 1100.0, seed: int = SEED + 1) -> Dict[str, torch.Tensor]:
    """L point lights on a Fibonacci lattice, grey intensity ~ U(0.5, 1.5) expanded to 3 channels."""
    g = torch.Generator().manual_seed(seed)
    i = torch.arange(L, dtype=torch.float64) + 0.5
    phi = torch.acos(1.0 - 2.0 * i / L)
    theta = math.pi * (1.0 + 5.0 ** 0.5) * i
    pos = radius * torch.stack([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)], -1)
    inten = (0.5 + torch.rand(L, 1, generator=g)).expand(L, 3)
    return dict(light_pos=pos.to(torch.float32)[None].contiguous(), light_intensity=inten[None].contiguous(),
                n_lights=torch.tensor([L], dtype=torch.int32))


def shade_inputs(G: int, seed: int = SEED + 2) -> Dict[str, torch.Tensor]:
    """SG lobes: un-normalised lobe_dirs ~ N(0,I), sigma = clamp(0.1 exp(N(0,0.5)), 0.01)."""
    g = torch.Generator().manual_seed(seed)
    lobe_dirs = torch.randn(1, G, 3, generator=g).contiguous()
    sigma = (0.1 * torch.exp(0.5 * torch.randn(1, G, generator=g))).clamp(min=0.01).contiguous()
    return dict(lobe_dirs=lobe_dirs, lobe_sigmas=sigma)
