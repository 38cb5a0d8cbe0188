"""GPU profiling driver for the bucket binning (csrc/splat_bin_tiles.cu): build the 300k-Gaussian / 1024x667 bench scene,
project it, then call gb_bin_tiles_pack a few times through the C ABI (what ncu attaches to), flushing L2 before each
call as the bench does.  Usage: python scripts/profile_bin_tiles.py [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from goliath_b200 import _lib, synthetic
from goliath_b200.gsplat import project_gaussians

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
H, W, BW = bench.H, bench.W, bench.BW
u = bench.unpack(bench.packed_scene(300_000).to(dev))
c = synthetic.ring_camera(0, img_h=H, img_w=W)
xys, depths, radii, conics, comp, nth, cov3d = project_gaussians(
    u["primpos"].contiguous(), u["primscale"].contiguous(), 1.0, u["primqvec"].contiguous(), c["viewmat"].to(dev),
    c["fx"], c["fy"], c["cx"], c["cy"], H, W, BW, 0.1)
G = xys.shape[0]
T = ((W + BW - 1) // BW) * ((H + BW - 1) // BW)
cap = 8 * G
L = _lib.lib()
st = _lib.stream_ptr(dev)
col3, op1 = u["diff_color"].contiguous(), u["opacity"].contiguous()
ws = torch.empty(L.gb_bin_tiles_workspace_bytes(G, T, cap), dtype=torch.uint8, device=dev)
bins = torch.empty(T, 2, dtype=torch.int32, device=dev)
order = torch.empty(T, dtype=torch.int32, device=dev)
gids = torch.empty(cap, dtype=torch.int32, device=dev)
rec = torch.empty(cap, 12, device=dev)
ovf = torch.zeros(1, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def call():
    _lib.check(L.gb_bin_tiles_pack(G, xys.data_ptr(), depths.data_ptr(), radii.data_ptr(), conics.data_ptr(),
                                   col3.data_ptr(), op1.data_ptr(), comp.data_ptr(), H, W, BW, cap, bins.data_ptr(),
                                   order.data_ptr(), 0, gids.data_ptr(), rec.data_ptr(), None, ovf.data_ptr(), ws.data_ptr(),
                                   st), "bin_tiles_pack")


for _ in range(reps):
    flush.fill_(1)
    call()
torch.cuda.synchronize()
# event timing of the whole call from a CUDA graph (no host launch gaps), L2 flushed before each replay
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    st = _lib.stream_ptr(dev)
    call()
torch.cuda.current_stream().wait_stream(side)
with torch.cuda.graph(g):
    st = _lib.stream_ptr(dev)
    call()
ts = []
for _ in range(10):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    b.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print("bin_tiles_pack from a CUDA graph: %.1f us (min %.1f)" % (sum(ts) / len(ts), min(ts)))
