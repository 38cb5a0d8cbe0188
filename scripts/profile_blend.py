"""GPU profiling driver: build the 300k-Gaussian / 1024x667 bench scene once, then launch the blend kernels a few
times through the C ABI (what ncu attaches to).  Usage: python scripts/profile_blend.py [channels] [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from goliath_b200 import _lib, synthetic
from goliath_b200.gsplat import project_gaussians, utils as gu

C = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "packed"
dev = torch.device("cuda:0")
H, W, BW = bench.H, bench.W, bench.BW
packed = bench.packed_scene(300_000).to(dev)
u = bench.unpack(packed)
c = synthetic.ring_camera(0, img_h=H, img_w=W)
xys, depths, radii, conics, comp, nth, cov3d = project_gaussians(
    u["primpos"].contiguous(), u["primscale"].contiguous(), 1.0, u["primqvec"].contiguous(), c["viewmat"].to(dev),
    c["fx"], c["fy"], c["cx"], c["cy"], H, W, BW, 0.1)
n, cum = gu.compute_cumulative_intersects(nth)
tb = gu._tile_bounds(H, W, BW)
_, _, _, gids, bins = gu.bin_and_sort_gaussians(xys.shape[0], n, xys, depths, radii, cum, tb, BW)
G = xys.shape[0]
colors = u["diff_color"].contiguous() if C == 3 else torch.cat([u["diff_color"], depths[:, None]], 1).contiguous()
opac = (u["opacity"] * comp[:, None]).contiguous()
bg = torch.zeros(C, device=dev)
out = torch.empty(H, W, C, device=dev); Ts = torch.empty(H, W, device=dev); fi = torch.empty(H, W, device=dev, dtype=torch.int32)
v_out = torch.ones(H, W, C, device=dev); v_a = torch.zeros(H, W, device=dev)
gx, gc, gcol, go = torch.zeros(G, 2, device=dev), torch.zeros(G, 3, device=dev), torch.zeros(G, C, device=dev), torch.zeros(G, 1, device=dev)
L = _lib.lib(); st = _lib.stream_ptr(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sched = L.gb_get_blend_mode() == 2
rec = torch.empty(n, 12, device=dev); order = torch.empty(L.gb_tile_schedule_ints(bins.shape[0]), dtype=torch.int32, device=dev)
_lib.check(L.gb_pack_records(n, C, gids.data_ptr(), xys.data_ptr(), conics.data_ptr(), colors.data_ptr(), opac.data_ptr(), rec.data_ptr(), st), "pack")
_lib.check((L.gb_tile_schedule if sched else L.gb_tile_order)(bins.shape[0], bins.data_ptr(), order.data_ptr(), st), "order")
ras_fwd = L.gb_rasterize_sched_fwd if sched else L.gb_rasterize_packed_fwd
ras_bwd = L.gb_rasterize_sched_bwd if sched else L.gb_rasterize_packed_bwd
import time
def ev(fn, k=10):
    ts = []
    for _ in range(k):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts) * 1e3
def pf():
    _lib.check(ras_fwd(H, W, C, bins.data_ptr(), order.data_ptr(), rec.data_ptr(), bg.data_ptr(), out.data_ptr(), Ts.data_ptr(), fi.data_ptr(), st), "pf")
def pb():
    _lib.check(ras_bwd(H, W, C, gids.data_ptr(), bins.data_ptr(), order.data_ptr(), rec.data_ptr(), bg.data_ptr(), Ts.data_ptr(), fi.data_ptr(),
                                         v_out.data_ptr(), v_a.data_ptr(), gx.data_ptr(), gc.data_ptr(), gcol.data_ptr(), go.data_ptr(), st), "pb")
def pk():
    _lib.check(L.gb_pack_records(n, C, gids.data_ptr(), xys.data_ptr(), conics.data_ptr(), colors.data_ptr(), opac.data_ptr(), rec.data_ptr(), st), "pack")
if mode == "packed":
    for _ in range(reps):
        flush.fill_(1); pk(); flush.fill_(1); pf(); flush.fill_(1); pb()
    torch.cuda.synchronize()
    if reps > 1:
        print("us: pack %.1f  packed_fwd %.1f  packed_bwd %.1f" % (ev(pk), ev(pf), ev(pb)))
for _ in range(reps if mode != "packed" else 0):
    flush.fill_(1)
    _lib.check(L.gb_rasterize_fwd(H, W, BW, C, gids.data_ptr(), bins.data_ptr(), xys.data_ptr(), conics.data_ptr(), colors.data_ptr(),
                                  opac.data_ptr(), bg.data_ptr(), out.data_ptr(), Ts.data_ptr(), fi.data_ptr(), st), "fwd")
    flush.fill_(1)
    _lib.check(L.gb_rasterize_bwd(H, W, BW, C, gids.data_ptr(), bins.data_ptr(), xys.data_ptr(), conics.data_ptr(), colors.data_ptr(),
                                  opac.data_ptr(), bg.data_ptr(), Ts.data_ptr(), fi.data_ptr(), v_out.data_ptr(), v_a.data_ptr(),
                                  gx.data_ptr(), gc.data_ptr(), gcol.data_ptr(), go.data_ptr(), st), "bwd")
torch.cuda.synchronize()
# work statistics for the roofline discussion
bins_c = bins.cpu(); fi_c = fi.cpu()
lens = (bins_c[:, 1] - bins_c[:, 0])
print("intersections", n, "tiles", bins_c.shape[0], "nonempty", int((lens > 0).sum()), "max list", int(lens.max()),
      "mean nonempty", float(lens[lens > 0].float().mean()))
# how deep into its list each tile actually walks (max final_idx per tile - bin start)
import numpy as np
fi_n = fi_c.numpy(); Tn = Ts.cpu().numpy()
depth = []
tbx = tb[0]
for ty in range(tb[1]):
    for tx in range(tbx):
        t = ty * tbx + tx
        lo, hi = int(bins_c[t, 0]), int(bins_c[t, 1])
        if hi > lo:
            blk = fi_n[ty * BW:(ty + 1) * BW, tx * BW:(tx + 1) * BW]
            sat = Tn[ty * BW:(ty + 1) * BW, tx * BW:(tx + 1) * BW]
            depth.append((hi - lo, int(blk.max()) - lo + 1, float((sat < 2e-4).mean())))
d = np.array(depth)
print("sum list", d[:, 0].sum(), "sum walked(last contributor)", d[:, 1].sum(), "mean saturated frac", d[:, 2].mean())
