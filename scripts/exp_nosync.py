"""Experiment: step time with and without the num_intersects host sync (same scene every step, so the count is known)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from goliath_b200 import synthetic
from goliath_b200.gsplat import utils as gu
dev = torch.device("cuda:0")
flat = bench.packed_scene(300_000).to(dev)
li = {k: v.to(dev) for k, v in synthetic.lights(32).items()}
c = synthetic.ring_camera(0, img_h=bench.H, img_w=bench.W)
cam = dict(Rt=c["viewmat"][None].to(dev), intr=(c["fx"], c["fy"], c["cx"], c["cy"]))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    leaves = {k: v.detach().requires_grad_() for k, v in bench.unpack(flat).items()}
    bench.gpu_step(leaves, cam, li)
    return torch.cat([leaves[k].grad.reshape(-1) for k, _ in bench.FIELDS])
def run(tag, K=30):
    for _ in range(5): step()
    torch.cuda.synchronize(); ts = []
    t0 = time.perf_counter()
    for _ in range(K):
        flush.fill_(1); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); ts.append((a, b))
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / K * 1e3
    print(tag, "event ms/step %.3f" % (sum(a.elapsed_time(b) for a, b in ts) / K), "wall ms/step (incl. flush) %.3f" % wall)
run("with sync   ")
# profiling experiment only: skip the host sync by monkeypatching the count (the scene is the same every step)
_orig = gu.compute_cumulative_intersects
def _no_sync(nth):
    cum = torch.empty_like(nth)
    ws = gu._workspace(nth.device, gu._lib.lib().gb_cumsum_workspace_bytes(nth.numel()))
    gu._lib.check(gu._lib.lib().gb_cumsum_i32(nth.numel(), nth.data_ptr(), cum.data_ptr(), ws.data_ptr(), gu._lib.stream_ptr(nth.device)), "cumsum")
    return 1078203, cum
gu.compute_cumulative_intersects = _no_sync
import goliath_b200.gsplat.rasterize as _r, goliath_b200.gsplat.fused as _f
_r.compute_cumulative_intersects = _f.compute_cumulative_intersects = _no_sync
run("without sync")
# CPU-only cost of issuing one step (GPU idle -> measures launch overhead): issue and do not wait
torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("CPU time to enqueue one step: %.3f ms" % ((t1 - t0) * 1e3))
import cProfile, pstats, io
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(20):
    step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45); print(s.getvalue()[:6000])
