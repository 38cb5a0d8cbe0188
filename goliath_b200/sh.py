"""Real spherical-harmonic basis in the reference's convention (ca_code/utils/sh.py:82-127, used by
rgca.PrimDecoder.forward:610 and rgca.AutoEncoder.forward:187-191): index k = n^2 + n + m for degree n and order
m in [-n, n]; Y_n^m = sqrt(2) K_n^|m| cos(m phi) P_n^m(cos theta) for m > 0, the sine for m < 0, K_n^0 P_n^0 for m = 0,
with the Condon-Shortley phase inside P, theta = acos(clamp(z, -1, 1)) and phi = atan2(y, x).

Own formulation: all associated Legendre values of one order are produced by the upward recurrence in the degree, and
cos(m phi) / sin(m phi) by the angle-addition recurrence instead of one trigonometric call per coefficient.  Pinned to
the reference's `dir2sh_torch` by tests/golden/sh_ref.npz (tests/test_sh.py)."""
import math
from typing import List

import torch as th


def _norm(m: int, n: int) -> float:
    """K_n^m = sqrt((2n+1)/(4 pi) * (n-m)!/(n+m)!)"""
    r = 1.0
    for i in range(n - m + 1, n + m + 1):
        r *= i
    return math.sqrt((2 * n + 1) / (4.0 * math.pi) / r)


def dir2sh(deg: int, dirs: th.Tensor) -> th.Tensor:
    """dirs [..., 3] (unit vectors) -> [..., (deg+1)^2] real SH values, reference ordering."""
    x, y, z = dirs[..., 0], dirs[..., 1], dirs[..., 2].clamp(-1.0, 1.0)
    ct = th.cos(th.acos(z))  # the reference evaluates P at cos(acos(z)); keep the same rounding
    st = th.sqrt(((1.0 + ct) * (1.0 - ct)).clamp(min=1e-8))
    phi = th.atan2(y, x)
    c1, s1 = th.cos(phi), th.sin(phi)
    out: List[th.Tensor] = [None] * ((deg + 1) ** 2)  # type: ignore[list-item]
    pmm = th.ones_like(ct)  # P_m^m
    cm, sm = th.ones_like(ct), th.zeros_like(ct)  # cos(m phi), sin(m phi)
    for m in range(0, deg + 1):
        if m > 0:
            pmm = -pmm * (2 * m - 1) * st
            cm, sm = cm * c1 - sm * s1, sm * c1 + cm * s1
        p_prev, p_cur = None, pmm
        for n in range(m, deg + 1):
            if n == m + 1:
                p_prev, p_cur = p_cur, ct * (2 * m + 1) * p_cur
            elif n > m + 1:
                p_prev, p_cur = p_cur, (ct * (2 * n - 1) * p_cur - (n + m - 1) * p_prev) / (n - m)
            k = _norm(m, n)
            base = n * n + n
            if m == 0:
                out[base] = k * p_cur
            else:
                out[base + m] = math.sqrt(2.0) * k * cm * p_cur
                out[base - m] = math.sqrt(2.0) * k * sm * p_cur
    return th.stack(out, dim=-1)
