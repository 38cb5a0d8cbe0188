"""CPU model of WHERE the blend's idle SM time comes from (analysis tool, not a test; uses the oracle).  Round 2.

ncu: both blend kernels keep the SMs active only ~0.72 of the launch.  All 581 occupied tiles of the bench view are
resident at once (4 x 148 CTA slots), so there is no dynamic scheduling: CTA i of the longest-first launch order lands
on SM i mod 148 and an SM's time is the sum of the TRUE costs of its ~4 tiles.  The true cost of a tile is what its
warps actually walk — front-to-back termination skips two thirds of the lists — not its list length.

Prints: processed fraction of the lists; the balance (mean SM load / max SM load) of round-robin placement by list
length (what ships), by true cost, boustrophedon by true cost, greedy LPT by true cost; the share of (tile, Gaussian)
intersections whose ellipse reaches alpha >= 1/255 on no pixel of the tile (what a tighter tile test in the binning
would drop).

Usage: python tests/model_blend_placement.py > profiles/r02_blend_placement_model.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import oracle  # noqa: E402
from goliath_b200 import synthetic  # noqa: E402

H, W, BW, NSM = bench.H, bench.W, bench.BW, 148


def main():
    u = {k: v.numpy() for k, v in bench.unpack(bench.packed_scene(300_000)).items()}
    c = synthetic.ring_camera(0, img_h=H, img_w=W)
    p = oracle.project_fwd(u["primpos"], u["primscale"], 1.0, u["primqvec"], c["viewmat"].numpy(), c["fx"], c["fy"],
                           c["cx"], c["cy"], H, W, BW, 0.1)
    b = oracle.bin_and_sort(p["xys"], p["depths"], p["radii"], p["num_tiles_hit"], H, W, BW)
    bins, gids, n = b["tile_bins"], b["gaussian_ids_sorted"], b["num_intersects"]
    colors = np.concatenate([u["diff_color"], p["depths"][:, None]], 1).astype(np.float32)
    opac = (u["opacity"] * p["compensation"][:, None]).astype(np.float32)
    _, Ts, fi = oracle.rasterize_fwd(H, W, BW, gids, bins, p["xys"], p["conics"], colors, opac, np.zeros(4, np.float32))
    tbx = (W + 15) // 16
    L = bins[:, 1] - bins[:, 0]
    occ = np.nonzero(L > 0)[0]
    # per (tile, 8x4 footprint): records walked before every pixel of the footprint has stopped
    wc = np.zeros((len(occ), 8), np.int64)
    for i, t in enumerate(occ):
        ty, tx = divmod(int(t), tbx)
        s, e = bins[t]
        for w in range(8):
            x0, y0 = tx * 16 + (w & 1) * 8, ty * 16 + (w >> 1) * 4
            f, T = fi[y0:y0 + 4, x0:x0 + 8], Ts[y0:y0 + 4, x0:x0 + 8]
            if f.size == 0:
                continue
            sat = (T < 5e-3) & (f < e - 1)
            wc[i, w] = int(f.max() + 1 - s) + 16 if sat.all() else int(e - s)
    cost = wc.sum(1)
    print("bench view: %d intersections, %d of %d tiles occupied, longest list %d" % (n, len(occ), len(L), L.max()))
    print("records walked by the warps / (8 x list lengths): %.3f" % (wc.sum() / (8.0 * L.sum())))

    def eff(seq):
        sm = np.zeros(NSM)
        for i, k in enumerate(seq):
            sm[i % NSM] += cost[k]
        return sm.mean() / sm.max()

    def snake(seq):
        out = []
        for r in range(0, len(seq), NSM):
            blk = list(seq[r:r + NSM])
            out += blk[::-1] if (r // NSM) % 2 else blk
        return out

    by_len = list(np.argsort(-L[occ], kind="stable"))
    by_cost = list(np.argsort(-cost, kind="stable"))
    loads = np.zeros(NSM)
    for k in by_cost:
        loads[int(np.argmin(loads))] += cost[k]
    print("balance = mean SM load / max SM load under static placement (CTA i -> SM i mod %d):" % NSM)
    print("  longest list first (shipped)      %.3f   <- ncu: SMs active 0.72 of the launch" % eff(by_len))
    print("  longest list first, boustrophedon %.3f" % eff(snake(by_len)))
    print("  true cost first                   %.3f" % eff(by_cost))
    print("  true cost first, boustrophedon    %.3f" % eff(snake(by_cost)))
    print("  greedy LPT by true cost           %.3f" % (loads.mean() / loads.max()))
    print("  (the true cost is known only after the forward; correlation with list length %.2f)"
          % np.corrcoef(L[occ], cost)[0, 1])

    # exact ellipse-vs-tile test of every intersection (the blend's cull test on the 16x16 rectangle)
    tile_of = np.repeat(np.arange(len(L)), L)
    ty, tx = np.divmod(tile_of, tbx)
    xy, con = p["xys"][gids].astype(np.float64), p["conics"][gids].astype(np.float64)
    op = opac[:, 0][gids].astype(np.float64)
    fx0, fy0 = tx * 16 + 0.5, ty * 16 + 0.5
    fx1, fy1 = np.minimum(tx * 16 + 15, W - 1) + 0.5, np.minimum(ty * 16 + 15, H - 1) + 0.5
    dxlo, dxhi, dylo, dyhi = xy[:, 0] - fx1, xy[:, 0] - fx0, xy[:, 1] - fy1, xy[:, 1] - fy0
    ex0, ey0 = np.clip(0, dxlo, dxhi), np.clip(0, dylo, dyhi)
    A, B, C = con[:, 0], con[:, 1], con[:, 2]
    dy1, dx2 = np.clip(-B * ex0 / C, dylo, dyhi), np.clip(-B * ey0 / A, dxlo, dxhi)
    f = np.minimum(0.5 * (A * ex0 ** 2 + C * dy1 ** 2) + B * ex0 * dy1, 0.5 * (A * dx2 ** 2 + C * ey0 ** 2) + B * dx2 * ey0)
    keep = f <= np.log(255 * np.maximum(op, 1e-30)) + 2e-3
    print("intersections whose ellipse reaches alpha >= 1/255 on some pixel of the tile: %.3f (gsplat bins by the 3-sigma"
          " bounding square); longest list %d -> %d" % (keep.mean(), L.max(), np.bincount(tile_of[keep], minlength=len(L)).max()))


if __name__ == "__main__":
    main()
