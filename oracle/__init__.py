"""CPU oracle for the goliath render hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; nothing under ``goliath_b200/`` does (tests/test_layout.py enforces it).

numpy-in / numpy-out wrappers over ``liboracle.so`` (plain C, built by ``oracle/Makefile`` with
``-ffp-contract=off``).  Each C function cites the reference file:line it restates.

Parity status per path (also in DESIGN.md):
  * SG shade (sg_oracle.c)        — pinned on the GPU box against the reference kernels rebuilt into oracle/_ref
  * splat path (splat_oracle.c)   — PARITY UNPINNED: gsplat==0.1.11 is a third-party dependency absent from
                                    /root/reference; restated from its published algorithm
  * raydirs / raymarch            — pinned against oracle/_ref (reference kernels) on the GPU box and against the
                                    reference's inline PyTorch raymarcher fixtures (tests/golden)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, a second or two)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("splat_oracle.c", "sg_oracle.c", "raymarch_oracle.c")]
    srcs = [s for s in srcs if os.path.exists(s)]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wall", "-o", so] + srcs + ["-lm"]
        )
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_cumsum.restype = ctypes.c_int64
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def num_threads():
    return int(lib().orc_num_threads())


# ----------------------------------------------------------------------------------------- splat path


def tile_bounds(H, W, bw):
    return (W + bw - 1) // bw, (H + bw - 1) // bw


def project_fwd(means3d, scales, glob_scale, quats, viewmat, fx, fy, cx, cy, H, W, block_width, clip_thresh=0.01):
    means3d, scales, quats = _f32(means3d), _f32(scales), _f32(quats)
    viewmat = _f32(np.asarray(viewmat).reshape(-1)[:12])
    G = means3d.shape[0]
    out = dict(
        cov3d=np.zeros((G, 6), np.float32), xys=np.zeros((G, 2), np.float32), depths=np.zeros(G, np.float32),
        radii=np.zeros(G, np.int32), conics=np.zeros((G, 3), np.float32), compensation=np.zeros(G, np.float32),
        num_tiles_hit=np.zeros(G, np.int32),
    )
    c = ctypes
    lib().orc_project_fwd(
        c.c_int(G), _p(means3d), _p(scales), c.c_float(glob_scale), _p(quats), _p(viewmat), c.c_float(fx),
        c.c_float(fy), c.c_float(cx), c.c_float(cy), c.c_int(H), c.c_int(W), c.c_int(block_width),
        c.c_float(clip_thresh), _p(out["cov3d"]), _p(out["xys"]), _p(out["depths"]), _p(out["radii"]),
        _p(out["conics"]), _p(out["compensation"]), _p(out["num_tiles_hit"]),
    )
    return out


def project_bwd(means3d, scales, glob_scale, quats, viewmat, fx, fy, cov3d, radii, conics, compensation,
                v_xy, v_depth, v_conic, v_compensation):
    means3d, scales, quats = _f32(means3d), _f32(scales), _f32(quats)
    viewmat = _f32(np.asarray(viewmat).reshape(-1)[:12])
    G = means3d.shape[0]
    out = dict(
        v_cov2d=np.zeros((G, 3), np.float32), v_cov3d=np.zeros((G, 6), np.float32),
        v_mean3d=np.zeros((G, 3), np.float32), v_scale=np.zeros((G, 3), np.float32),
        v_quat=np.zeros((G, 4), np.float32),
    )
    c = ctypes
    args = [_f32(cov3d), _i32(radii), _f32(conics), _f32(compensation), _f32(v_xy), _f32(v_depth), _f32(v_conic),
            _f32(v_compensation)]
    lib().orc_project_bwd(
        c.c_int(G), _p(means3d), _p(scales), c.c_float(glob_scale), _p(quats), _p(viewmat), c.c_float(fx),
        c.c_float(fy), *[_p(a) for a in args], _p(out["v_cov2d"]), _p(out["v_cov3d"]), _p(out["v_mean3d"]),
        _p(out["v_scale"]), _p(out["v_quat"]),
    )
    return out


def bin_and_sort(xys, depths, radii, num_tiles_hit, H, W, block_width):
    """cumsum -> map_gaussian_to_intersects -> stable sort -> tile bin edges (SURVEY.md App. A 'Binning')."""
    xys, depths, radii, nth = _f32(xys), _f32(depths), _i32(radii), _i32(num_tiles_hit)
    G = xys.shape[0]
    cum = np.zeros(G, np.int32)
    n = int(lib().orc_cumsum(ctypes.c_int(G), _p(nth), _p(cum)))
    isect = np.zeros(n, np.int64)
    gids = np.zeros(n, np.int32)
    lib().orc_map_to_intersects(ctypes.c_int(G), _p(xys), _p(depths), _p(radii), _p(cum), ctypes.c_int(H),
                                ctypes.c_int(W), ctypes.c_int(block_width), _p(isect), _p(gids))
    isect_s = np.zeros(n, np.int64)
    gids_s = np.zeros(n, np.int32)
    lib().orc_sort_intersects(ctypes.c_int64(n), _p(isect), _p(gids), _p(isect_s), _p(gids_s))
    tbx, tby = tile_bounds(H, W, block_width)
    bins = np.zeros((tbx * tby, 2), np.int32)
    lib().orc_tile_bin_edges(ctypes.c_int64(n), _p(isect_s), _p(bins))
    return dict(num_intersects=n, cum_tiles_hit=cum, isect_ids=isect, gaussian_ids=gids, isect_ids_sorted=isect_s,
                gaussian_ids_sorted=gids_s, tile_bins=bins)


def rasterize_fwd(H, W, block_width, gids_sorted, tile_bins, xys, conics, colors, opacities, background):
    colors = _f32(colors)
    C = colors.shape[1]
    assert C <= 8
    out_img = np.zeros((H, W, C), np.float32)
    final_Ts = np.zeros((H, W), np.float32)
    final_idx = np.zeros((H, W), np.int32)
    a = [_i32(gids_sorted), _i32(tile_bins), _f32(xys), _f32(conics), colors, _f32(np.asarray(opacities).reshape(-1)),
         _f32(background)]
    lib().orc_rasterize_fwd(ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(block_width), ctypes.c_int(C),
                            *[_p(x) for x in a], _p(out_img), _p(final_Ts), _p(final_idx))
    return out_img, final_Ts, final_idx


def rasterize_bwd(H, W, block_width, gids_sorted, tile_bins, xys, conics, colors, opacities, background, final_Ts,
                  final_idx, v_output, v_output_alpha):
    colors = _f32(colors)
    G, C = colors.shape
    v_xy = np.zeros((G, 2), np.float32)
    v_conic = np.zeros((G, 3), np.float32)
    v_colors = np.zeros((G, C), np.float32)
    v_opacity = np.zeros((G, 1), np.float32)
    a = [_i32(gids_sorted), _i32(tile_bins), _f32(xys), _f32(conics), colors, _f32(np.asarray(opacities).reshape(-1)),
         _f32(background), _f32(final_Ts), _i32(final_idx), _f32(v_output), _f32(v_output_alpha)]
    lib().orc_rasterize_bwd(ctypes.c_int(G), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(block_width),
                            ctypes.c_int(C), *[_p(x) for x in a], _p(v_xy), _p(v_conic), _p(v_colors), _p(v_opacity))
    return v_xy, v_conic, v_colors, v_opacity


# ----------------------------------------------------------------------------------------- SG shade


def sg_fwd(lobe_dirs, lobe_sigmas, light_values, light_pts, prim_pts, n_lights, w_type=0):
    lobe_dirs = _f32(lobe_dirs)
    N, D = lobe_dirs.shape[:2]
    light_values = _f32(light_values)
    L = light_values.shape[1]
    out = np.zeros((N, D, 3), np.float32)
    a = [lobe_dirs, _f32(np.asarray(lobe_sigmas).reshape(N, D)), light_values, _f32(light_pts), _f32(prim_pts),
         _i32(n_lights)]
    lib().orc_sg_fwd(ctypes.c_int(N), ctypes.c_int(D), ctypes.c_int(L), *[_p(x) for x in a], _p(out),
                     ctypes.c_int(w_type))
    return out


def sg_bwd(lobe_dirs, lobe_sigmas, light_values, light_pts, prim_pts, n_lights, grad_integral, w_type=0,
           want_light_grad=True):
    lobe_dirs = _f32(lobe_dirs)
    N, D = lobe_dirs.shape[:2]
    light_values = _f32(light_values)
    L = light_values.shape[1]
    gd = np.zeros((N, D, 3), np.float32)
    gs = np.zeros((N, D), np.float32)
    gl = np.zeros((N, L, 3), np.float32) if want_light_grad else None
    a = [lobe_dirs, _f32(np.asarray(lobe_sigmas).reshape(N, D)), light_values, _f32(light_pts), _f32(prim_pts),
         _i32(n_lights), _f32(grad_integral)]
    lib().orc_sg_bwd(ctypes.c_int(N), ctypes.c_int(D), ctypes.c_int(L), *[_p(x) for x in a], _p(gd), _p(gs), _p(gl),
                     ctypes.c_int(w_type))
    return gd, gs, gl
