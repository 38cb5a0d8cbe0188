"""CPU oracle of the WHOLE benchmarked step (bench.py's `gpu_step`): SG shade + colour compose -> EWA projection ->
bin/sort -> one 4-channel blend (rgb + depth) -> rgca.AutoEncoder.render post-processing, forward and backward with
v_out = 1, from the C oracle's pieces (oracle/*.c).  Test infrastructure only.

Reference call sequence restated: ca_code/models/rgca.py:557-575 (shade + compose), ca_code/utils/render_gsplat.py:41-106
(project, rasterise rgb, rasterise depth-as-colour), rgca.py:112-151 (alpha from the DETACHED final_T, depth divided by
alpha.clamp(0.05, 1))."""
import numpy as np


def oracle_step(orc, u, cam, li, H, W, bw=16):
    """u: dict of numpy arrays with bench.FIELDS names; cam: dict(Rt [3,4], intr (fx,fy,cx,cy)); li: numpy lights.
    Returns dict(rgb [3,H,W], alpha [H,W], depth [H,W], n_isect, grads {field: array})."""
    f32 = np.float32
    fx, fy, cx, cy = cam["intr"]
    V = np.asarray(cam["Rt"], f32)
    raw = u["lobe_dirs"].astype(np.float64)
    nrm_len = np.linalg.norm(raw, axis=-1, keepdims=True)
    nrm = (raw / nrm_len).astype(f32)
    spec = orc.sg_fwd(nrm[None], u["sigma"][None], li["light_intensity"], li["light_pos"], u["primpos"][None],
                      li["n_lights"], 0)[0]
    vis = u["spec_vis"].reshape(-1, 1)
    pre = np.maximum(u["diff_color"], 0) + spec * vis
    color = np.maximum(pre, 0).astype(f32)
    p = orc.project_fwd(u["primpos"], u["primscale"], 1.0, u["primqvec"], V, fx, fy, cx, cy, H, W, bw, 0.1)
    b = orc.bin_and_sort(p["xys"], p["depths"], p["radii"], p["num_tiles_hit"], H, W, bw)
    opac = (u["opacity"].reshape(-1) * p["compensation"]).astype(f32)
    col4 = np.concatenate([color, p["depths"][:, None]], 1).astype(f32)
    z4 = np.zeros(4, f32)
    out4, Ts, fi = orc.rasterize_fwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"], col4,
                                     opac, z4)
    alpha = (1.0 - Ts).astype(f32)
    inv_a = (1.0 / np.clip(alpha, 0.05, 1.0)).astype(f32)
    rgb = np.transpose(out4[..., :3], (2, 0, 1))
    depth = out4[..., 3] * inv_a
    # backward, loss = sum(rgb) + sum(depth); alpha is detached (rgca.py:137)
    v_out4 = np.ones((H, W, 4), f32)
    v_out4[..., 3] = inv_a
    g = orc.rasterize_bwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"], col4, opac, z4,
                          Ts, fi, v_out4, np.zeros((H, W), f32))
    v_xy, v_conic, v_col4, v_opeff = g
    v_opeff = v_opeff.reshape(-1)
    pb = orc.project_bwd(u["primpos"], u["primscale"], 1.0, u["primqvec"], V, fx, fy, p["cov3d"], p["radii"], p["conics"],
                         p["compensation"], v_xy, v_col4[:, 3].copy(), v_conic, v_opeff * u["opacity"].reshape(-1))
    g_color = v_col4[:, :3] * (pre >= 0)
    g_diff = g_color * (u["diff_color"] >= 0)
    g_spec = g_color * vis
    g_vis = (g_color * spec).sum(-1)
    gd, gs, _ = orc.sg_bwd(nrm[None], u["sigma"][None], li["light_intensity"], li["light_pos"], u["primpos"][None],
                           li["n_lights"], g_spec[None].astype(f32), 0, want_light_grad=False)
    gd = gd[0].astype(np.float64)
    n64 = nrm.astype(np.float64)
    g_raw = (gd - n64 * (n64 * gd).sum(-1, keepdims=True)) / nrm_len
    grads = dict(primpos=pb["v_mean3d"], primqvec=pb["v_quat"], primscale=pb["v_scale"],
                 opacity=(v_opeff * p["compensation"]).reshape(u["opacity"].shape), diff_color=g_diff.astype(f32),
                 lobe_dirs=g_raw.astype(f32), sigma=gs[0], spec_vis=g_vis.reshape(u["spec_vis"].shape).astype(f32))
    return dict(rgb=rgb, alpha=alpha, depth=depth, n_isect=int(b["num_intersects"]), grads=grads,
                bins=b["tile_bins"], gids=b["gaussian_ids_sorted"])
