"""CPU: pin the C oracle of the splat path against an independent float64 torch restatement + autograd.

The splat arithmetic lives in gsplat==0.1.11 (absent, parity UNPINNED — see oracle/splat_oracle.c); these tests
make sure the restatement is at least self-consistent: forward == fp64 restatement, hand-written backward ==
autograd of the forward, binning invariants, and the documented quirks behave as documented."""
import numpy as np
import pytest
import torch

from util import assert_close, rel_err, small_scene


def torch_project(means, scales, quats, V, fx, fy, cx, cy):
    """float64, differentiable; quats treated as already-unit polynomial (gsplat vjp convention)."""
    R, t = V[:, :3], V[:, 3]
    pv = means @ R.T + t
    w, x, y, z = quats.unbind(-1)
    Rq = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    M = Rq * scales[:, None, :]
    S3 = M @ M.transpose(1, 2)
    tx, ty, tz = pv.unbind(-1)
    rz = 1 / tz
    zero = torch.zeros_like(rz)
    J = torch.stack([fx * rz, zero, -fx * tx * rz * rz, zero, fy * rz, -fy * ty * rz * rz], -1).reshape(-1, 2, 3)
    T = J @ R
    cov = T @ S3 @ T.transpose(1, 2)
    c00, c01, c11 = cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 1]
    det0 = c00 * c11 - c01 * c01
    A, B, C = c00 + 0.3, c01, c11 + 0.3
    det1 = A * C - B * B
    comp = torch.sqrt(torch.clamp(det0 / det1, min=0))
    conic = torch.stack([C / det1, -B / det1, A / det1], -1)
    rw = 1 / (tz + 1e-6)
    xys = torch.stack([tx * rw * fx + cx, ty * rw * fy + cy], -1)
    return xys, tz, conic, comp


@pytest.fixture(scope="module")
def scene():
    return small_scene(G=1500)


def test_project_forward_matches_fp64(orc, scene):
    s = scene
    out = orc.project_fwd(s["means3d"], s["scales"], 1.0, s["quats"], s["viewmat"], s["fx"], s["fy"], s["cx"],
                          s["cy"], s["img_h"], s["img_w"], 16, 0.1)
    vis = out["radii"] > 0
    assert vis.sum() > 500
    d = lambda a: torch.tensor(a, dtype=torch.float64)
    xys, depth, conic, comp = torch_project(d(s["means3d"]), d(s["scales"]), d(s["quats"]), d(s["viewmat"]),
                                            s["fx"], s["fy"], s["cx"], s["cy"])
    assert_close(out["xys"][vis], xys.numpy()[vis], rtol=1e-5, what="xys")
    assert_close(out["depths"][vis], depth.numpy()[vis], rtol=1e-6, what="depths")
    assert_close(out["conics"][vis], conic.numpy()[vis], rtol=2e-4, atol=1e-5, what="conics")
    assert_close(out["compensation"][vis], comp.numpy()[vis], rtol=1e-4, what="compensation")
    # radius / tile-count definitions
    cov = np.linalg.inv(np.stack([[out["conics"][:, 0], out["conics"][:, 1]],
                                  [out["conics"][:, 1], out["conics"][:, 2]]]).transpose(2, 0, 1)[vis].astype(np.float64))
    lam = np.linalg.eigvalsh(cov)[:, 1]
    assert np.all(np.abs(out["radii"][vis] - np.ceil(3 * np.sqrt(lam))) <= 1)
    # culled Gaussians keep zeros
    assert np.all(out["num_tiles_hit"][~vis] == 0) and np.all(out["xys"][~vis] == 0)


def test_project_near_clip_and_offscreen(orc):
    means = np.array([[0, 0, 5.0], [0, 0, 0.05], [0, 0, -3.0], [1e4, 0, 5.0]], np.float32)
    scales = np.full((4, 3), 0.01, np.float32)
    quats = np.tile(np.array([[1, 0, 0, 0]], np.float32), (4, 1))
    V = np.eye(4, dtype=np.float32)[:3]
    out = orc.project_fwd(means, scales, 1.0, quats, V, 100.0, 100.0, 32.0, 32.0, 64, 64, 16, 0.1)
    assert out["radii"][0] > 0 and out["num_tiles_hit"][0] >= 1
    assert out["radii"][1] == 0 and out["radii"][2] == 0          # z <= clip_thresh
    assert out["num_tiles_hit"][3] == 0 and out["radii"][3] == 0  # bbox outside the tile grid
    assert np.any(out["conics"][3] != 0)                          # conic is written before the bbox test


def test_project_backward_matches_autograd(orc, scene):
    s = scene
    out = orc.project_fwd(s["means3d"], s["scales"], 1.0, s["quats"], s["viewmat"], s["fx"], s["fy"], s["cx"],
                          s["cy"], s["img_h"], s["img_w"], 16, 0.1)
    vis = out["radii"] > 0
    G = len(vis)
    rng = np.random.default_rng(0)
    v_xy = rng.standard_normal((G, 2)).astype(np.float32)
    v_depth = rng.standard_normal(G).astype(np.float32)
    v_conic = rng.standard_normal((G, 3)).astype(np.float32)
    v_comp = rng.standard_normal(G).astype(np.float32)
    g = orc.project_bwd(s["means3d"], s["scales"], 1.0, s["quats"], s["viewmat"], s["fx"], s["fy"], out["cov3d"],
                        out["radii"], out["conics"], out["compensation"], v_xy, v_depth, v_conic, v_comp)
    d = lambda a: torch.tensor(a, dtype=torch.float64)
    m, sc, q = d(s["means3d"]).requires_grad_(), d(s["scales"]).requires_grad_(), d(s["quats"]).requires_grad_()
    xys, depth, conic, comp = torch_project(m, sc, q, d(s["viewmat"]), s["fx"], s["fy"], s["cx"], s["cy"])
    mask = torch.tensor(vis)
    loss = ((xys * d(v_xy)).sum(-1) + depth * d(v_depth) + (conic * d(v_conic)).sum(-1) + comp * d(v_comp))[mask].sum()
    loss.backward()
    assert_close(g["v_mean3d"], m.grad.numpy(), rtol=2e-4, what="v_mean3d")
    assert_close(g["v_scale"], sc.grad.numpy(), rtol=2e-4, what="v_scale")
    assert_close(g["v_quat"], q.grad.numpy(), rtol=2e-4, what="v_quat")
    assert np.all(g["v_mean3d"][~vis] == 0) and np.all(g["v_quat"][~vis] == 0)


def test_binning_invariants(orc):
    s = small_scene(G=4000, depth_quant=True)
    p = orc.project_fwd(s["means3d"], s["scales"], 1.0, s["quats"], s["viewmat"], s["fx"], s["fy"], s["cx"], s["cy"],
                        s["img_h"], s["img_w"], 16, 0.1)
    b = orc.bin_and_sort(p["xys"], p["depths"], p["radii"], p["num_tiles_hit"], s["img_h"], s["img_w"], 16)
    n = b["num_intersects"]
    assert n == int(p["num_tiles_hit"].sum()) and n > 1000
    ks, gs = b["isect_ids_sorted"], b["gaussian_ids_sorted"]
    assert np.all(np.diff(ks) >= 0)
    # exact depth ties exist and are broken by ascending Gaussian id (stable sort of id-ordered emission)
    ties = np.diff(ks) == 0
    assert ties.sum() > 0
    assert np.all(np.diff(gs)[ties] > 0)
    # key layout
    assert np.array_equal((ks & 0xFFFFFFFF).astype(np.uint32), p["depths"][gs].view(np.uint32))
    tiles = (ks >> 32).astype(np.int64)
    T = b["tile_bins"].shape[0]
    for t in np.unique(tiles):
        lo, hi = b["tile_bins"][t]
        assert np.all(tiles[lo:hi] == t) and (lo == 0 or tiles[lo - 1] != t) and (hi == n or tiles[hi] != t)
    untouched = np.setdiff1d(np.arange(T), np.unique(tiles))
    assert np.all(b["tile_bins"][untouched] == 0)
    # multiset of (tile, gaussian) pairs is preserved by the sort
    a = np.stack([b["isect_ids"] >> 32, b["gaussian_ids"]], 1)
    c = np.stack([tiles, gs], 1)
    assert np.array_equal(a[np.lexsort(a.T[::-1])], c[np.lexsort(c.T[::-1])])


@pytest.mark.parametrize("quant", [False, True])
def test_bucket_binning_is_equivalent_to_the_key_sort(orc, quant):
    """The premise of csrc/splat_bin_tiles.cu, checked on the CPU against the oracle's key sort: rank the Gaussians once
    by (depth bits, id), bucket the (Gaussian, tile) pairs per tile in ANY order, then order every bucket by setting
    bits in a G-bit bitmap and reading them back — this reproduces gaussian_ids_sorted and tile_bins exactly, depth
    ties included (a tile holds a Gaussian at most once, so its ranks are unique)."""
    s = small_scene(G=4000, depth_quant=quant, seed=13)
    bw = 16
    p = orc.project_fwd(s["means3d"], s["scales"] * np.float32(8.0), 1.0, s["quats"], s["viewmat"], s["fx"], s["fy"],
                        s["cx"], s["cy"], s["img_h"], s["img_w"], bw, 0.1)
    b = orc.bin_and_sort(p["xys"], p["depths"], p["radii"], p["num_tiles_hit"], s["img_h"], s["img_w"], bw)
    G = len(p["radii"])
    tbx, tby = (s["img_w"] + bw - 1) // bw, (s["img_h"] + bw - 1) // bw
    # 1. depth ranks: stable sort of the 32 depth bits, emission order = ascending id
    order = np.argsort(p["depths"].view(np.uint32), kind="stable")
    rank_of = np.empty(G, np.int64)
    rank_of[order] = np.arange(G)
    # 2. buckets, filled in a scrambled order (the kernel claims slots with atomics)
    tiles_of = b["isect_ids"] >> 32
    perm = np.random.default_rng(5).permutation(len(tiles_of))
    buckets = {}
    for t, g in zip(tiles_of[perm], b["gaussian_ids"][perm]):
        buckets.setdefault(int(t), []).append(int(rank_of[g]))
    # 3. bitmap sort per tile, bins from the counts
    gids, bins, start = [], np.zeros((tbx * tby, 2), np.int32), 0
    for t in range(tbx * tby):
        ranks = buckets.get(t, [])
        if not ranks:
            continue
        bitmap = np.zeros(G, bool)
        bitmap[ranks] = True
        assert bitmap.sum() == len(ranks), "a tile holds every Gaussian at most once"
        gids.append(order[np.nonzero(bitmap)[0]])
        bins[t] = (start, start + len(ranks))
        start += len(ranks)
    assert start == b["num_intersects"] > 1000
    assert np.array_equal(bins, b["tile_bins"])
    assert np.array_equal(np.concatenate(gids).astype(np.int32), b["gaussian_ids_sorted"])
    if quant:
        assert (np.diff(b["isect_ids_sorted"]) == 0).sum() > 0, "the scene must contain exact depth ties"


def torch_blend(H, W, bw, gids_sorted, tile_bins, xys, conics, colors, opac, bg, alpha_max=0.999):
    """float64 differentiable restatement of the per-pixel front-to-back blend (python loop over the list)."""
    tbx = (W + bw - 1) // bw
    C = colors.shape[1]
    out = torch.zeros(H, W, C, dtype=torch.float64)
    Tout = torch.ones(H, W, dtype=torch.float64)
    fidx = torch.zeros(H, W, dtype=torch.int64)
    for ty in range((H + bw - 1) // bw):
        for tx in range(tbx):
            lo, hi = (int(v) for v in tile_bins[ty * tbx + tx])
            ys = torch.arange(ty * bw, min((ty + 1) * bw, H))
            xs = torch.arange(tx * bw, min((tx + 1) * bw, W))
            py, px = torch.meshgrid(ys.double() + 0.5, xs.double() + 0.5, indexing="ij")
            T = torch.ones_like(px)
            done = torch.zeros_like(px, dtype=torch.bool)
            acc = torch.zeros(*px.shape, C, dtype=torch.float64)
            fi = torch.zeros_like(px, dtype=torch.int64)
            for k in range(lo, hi):
                g = int(gids_sorted[k])
                dx, dy = xys[g, 0] - px, xys[g, 1] - py
                sigma = 0.5 * (conics[g, 0] * dx * dx + conics[g, 2] * dy * dy) + conics[g, 1] * dx * dy
                alpha = torch.clamp(opac[g] * torch.exp(-sigma), max=alpha_max)
                skip = (sigma < 0) | (alpha < 1.0 / 255.0)
                nT = T * (1 - alpha)
                stop = (~skip) & (~done) & (nT <= 1e-4)
                done = done | stop
                use = (~skip) & (~done)
                acc = acc + torch.where(use[..., None], colors[g] * (alpha * T)[..., None], torch.zeros(1, dtype=torch.float64))
                T = torch.where(use, nT, T)
                fi = torch.where(use, torch.full_like(fi, k), fi)
            sl = (slice(ty * bw, ty * bw + len(ys)), slice(tx * bw, tx * bw + len(xs)))
            out[sl] = acc + T[..., None] * bg
            Tout[sl] = T
            fidx[sl] = fi
    return out, Tout, fidx


@pytest.fixture(scope="module")
def blend_case(orc):
    s = small_scene(G=400, img_h=40, img_w=36, seed=3)
    s["opacity"] = np.clip(s["opacity"], 0.05, 0.95).astype(np.float32)  # keep alpha < 0.99: no clamp quirk
    p = orc.project_fwd(s["means3d"], s["scales"] * 25.0, 1.0, s["quats"], s["viewmat"], s["fx"], s["fy"], s["cx"],
                        s["cy"], s["img_h"], s["img_w"], 8, 0.1)
    b = orc.bin_and_sort(p["xys"], p["depths"], p["radii"], p["num_tiles_hit"], s["img_h"], s["img_w"], 8)
    return s, p, b


@pytest.mark.parametrize("C", [3, 4])
def test_blend_forward_and_backward_match_autograd(orc, blend_case, C):
    s, p, b = blend_case
    H, W, bw = s["img_h"], s["img_w"], 8
    rng = np.random.default_rng(1)
    colors = rng.random((len(s["colors"]), C)).astype(np.float32)
    bg = rng.random(C).astype(np.float32)
    opac = (s["opacity"][:, 0] * p["compensation"]).astype(np.float32)
    out, Ts, fidx = orc.rasterize_fwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"],
                                      colors, opac, bg)
    d = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    xy_t, con_t, col_t, op_t = (d(p["xys"]).requires_grad_(), d(p["conics"]).requires_grad_(),
                                d(colors).requires_grad_(), d(opac).requires_grad_())
    o2, T2, f2 = torch_blend(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], xy_t, con_t, col_t, op_t, d(bg))
    assert (Ts < 0.5).mean() > 0.1, "scene should actually cover pixels"
    assert_close(out, o2.detach().numpy(), rtol=1e-5, atol=1e-6, what="out_img")
    assert_close(Ts, T2.detach().numpy(), rtol=1e-5, atol=1e-7, what="final_Ts")
    assert (fidx == f2.numpy()).mean() > 0.999
    v_out = rng.standard_normal((H, W, C)).astype(np.float32)
    v_alpha = rng.standard_normal((H, W)).astype(np.float32)
    ((o2 * d(v_out)).sum() + ((1 - T2) * d(v_alpha)).sum()).backward()
    v_xy, v_conic, v_col, v_op = orc.rasterize_bwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"],
                                                   p["conics"], colors, opac, bg, Ts, fidx, v_out, v_alpha)
    assert_close(v_col, col_t.grad.numpy(), rtol=1e-4, what="v_colors")
    assert_close(v_op[:, 0], op_t.grad.numpy(), rtol=1e-4, what="v_opacity")
    assert_close(v_xy, xy_t.grad.numpy(), rtol=1e-4, what="v_xy")
    assert_close(v_conic, con_t.grad.numpy(), rtol=1e-4, what="v_conic")


def test_blend_backward_alpha_clamp_quirk(orc):
    """gsplat 0.1.x backward clamps alpha at 0.99 (forward 0.999): one opaque Gaussian over one pixel."""
    H = W = 4
    xys = np.array([[2.0, 2.0]], np.float32)
    conics = np.array([[0.01, 0.0, 0.01]], np.float32)
    colors = np.array([[1.0, 0.5, 0.25]], np.float32)
    opac = np.array([1.0], np.float32)
    bins = np.array([[0, 1]], np.int32)
    gids = np.array([0], np.int32)
    bg = np.zeros(3, np.float32)
    out, Ts, fidx = orc.rasterize_fwd(H, W, 4, gids, bins, xys, conics, colors, opac, bg)
    a_f = min(0.999, float(np.exp(-0.5 * 0.01 * 0.5)))  # pixel (1.5,1.5): dx=dy=0.5
    assert abs(Ts[1, 1] - (1 - a_f)) < 1e-6
    v_out = np.ones((H, W, 3), np.float32)
    v_xy, v_conic, v_col, v_op = orc.rasterize_bwd(H, W, 4, gids, bins, xys, conics, colors, opac, bg, Ts, fidx,
                                                   v_out, np.zeros((H, W), np.float32))
    # v_colors = sum_pix alpha_b * T_b with alpha_b = min(.99, .), T_b = T_final / (1 - alpha_b)
    exp = 0.0
    for i in range(H):
        for j in range(W):
            dx, dy = 2.0 - (j + 0.5), 2.0 - (i + 0.5)
            vis = np.exp(-(0.5 * (0.01 * dx * dx + 0.01 * dy * dy)))
            ab = min(0.99, vis)
            exp += ab * float(Ts[i, j]) / (1 - ab)
    assert abs(v_col[0, 0] - exp) / exp < 1e-4
