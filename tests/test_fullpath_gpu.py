"""Parity of the BENCHMARKED path at the BENCHMARKED sizes: bench.py's own `gpu_step`
(shade_compose -> render_views(capacity=...) -> default bucket binning -> default blend mode, forward + every gradient)
against the CPU oracle of the whole step (tests/fullstep.py over oracle/*.c) at BASELINE config 2
(300 000 Gaussians, 1024x667, L = 32) for two ring cameras, and at the native 1 048 576 Gaussians."""
import numpy as np
import pytest
import torch

from fullstep import oracle_step
from util import assert_close, t2n

pytestmark = pytest.mark.gpu


def _run_gpu(cuda, G, cam_k, lights=32):
    import bench
    from goliath_b200 import synthetic
    from goliath_b200.gsplat.fused import check_overflow

    packed = bench.packed_scene(G)
    li_h = synthetic.lights(lights)
    li = {k: v.to(cuda) for k, v in li_h.items()}
    c = synthetic.ring_camera(cam_k, img_h=bench.H, img_w=bench.W)
    cam = dict(Rt=c["viewmat"][None].to(cuda), intr=(c["fx"], c["fy"], c["cx"], c["cy"]))
    leaves = {k: v.detach().requires_grad_() for k, v in bench.unpack(packed.to(cuda)).items()}
    cap = max(8 * G, 1 << 20)  # bench.py's sync-free capacity
    rgb, alpha, depth = bench.gpu_step(leaves, cam, li, capacity=cap)
    torch.cuda.synchronize()
    assert not check_overflow(cuda)
    host = dict(Rt=c["viewmat"].numpy(), intr=cam["intr"])
    return packed, li_h, host, rgb, alpha, depth, {k: t2n(v.grad) for k, v in leaves.items()}


@pytest.mark.parametrize("G,cam_k", [(300_000, 0), (300_000, 5), (1_048_576, 2)])
def test_benchmarked_step_matches_oracle(orc, cuda, G, cam_k):
    import bench

    packed, li_h, cam, rgb, alpha, depth, grads = _run_gpu(cuda, G, cam_k)
    u = {k: np.ascontiguousarray(v) for k, v in bench.unpack(packed.numpy()).items()}
    li = {k: v.numpy() for k, v in li_h.items()}
    ref = oracle_step(orc, u, cam, li, bench.H, bench.W)
    assert ref["n_isect"] > 2 * G, "the scene must be the dense bench scene"
    # pixels: 1e-4 relative (north_star); a borderline alpha < 1/255 or T <= 1e-4 decision may flip on a handful of pixels
    assert_close(t2n(rgb[0]), ref["rgb"], rtol=1e-4, atol=2e-5, frac=0.9995, what="rgb")
    assert_close(t2n(alpha[0, 0]), ref["alpha"], rtol=1e-4, atol=2e-6, frac=0.9995, what="alpha")
    assert_close(t2n(depth[0, 0]), ref["depth"], rtol=1e-4, atol=2e-2, frac=0.9995, what="depth")  # depth ~ 1000 mm
    for k, want in ref["grads"].items():
        got = grads[k].reshape(want.shape)
        assert np.isfinite(got).all(), k
        assert_close(got, want, rtol=1e-4, atol=1e-4 * float(np.abs(want).max()) * 0.1, frac=0.999, what="grad " + k)


def test_benchmarked_binning_is_bit_exact_at_bench_size(orc, cuda):
    """tile_bins and sorted ids of the product's bucket binning at 300k / 1024x667 against the ORACLE's key sort
    (bench.py only checks it against this library's own key sort)."""
    import bench
    from goliath_b200 import _lib, synthetic
    from goliath_b200.gsplat import project_gaussians

    G = 300_000
    u_t = bench.unpack(bench.packed_scene(G).to(cuda))
    c = synthetic.ring_camera(3, img_h=bench.H, img_w=bench.W)
    H, W, BW = bench.H, bench.W, 16
    xys, depths, radii, conics, comp, nth, cov3d = project_gaussians(
        u_t["primpos"].contiguous(), u_t["primscale"].contiguous(), 1.0, u_t["primqvec"].contiguous(),
        c["viewmat"].to(cuda), c["fx"], c["fy"], c["cx"], c["cy"], H, W, BW, 0.1)
    u = {k: np.ascontiguousarray(t2n(v)) for k, v in u_t.items()}
    p = orc.project_fwd(u["primpos"], u["primscale"], 1.0, u["primqvec"], c["viewmat"].numpy(), c["fx"], c["fy"], c["cx"],
                        c["cy"], H, W, BW, 0.1)
    for name, got in (("xys", xys), ("depths", depths), ("radii", radii), ("num_tiles_hit", nth)):
        a, r = t2n(got), p[name]
        assert np.array_equal(a.view(np.uint32) if a.dtype == np.float32 else a,
                              r.view(np.uint32) if r.dtype == np.float32 else r), name
    b = orc.bin_and_sort(p["xys"], p["depths"], p["radii"], p["num_tiles_hit"], H, W, BW)
    n = b["num_intersects"]
    L = _lib.lib()
    T = ((W + 15) // 16) * ((H + 15) // 16)
    cap = n + 4096
    ws = torch.empty(L.gb_bin_tiles_workspace_bytes(G, T, cap), dtype=torch.uint8, device=cuda)
    bins = torch.empty(T, 2, dtype=torch.int32, device=cuda)
    order = torch.empty(L.gb_tile_schedule_ints(T), dtype=torch.int32, device=cuda)
    gids = torch.empty(cap, dtype=torch.int32, device=cuda)
    rec = torch.empty(cap, 12, device=cuda)
    ovf = torch.zeros(1, dtype=torch.int32, device=cuda)
    col3, op1 = u_t["diff_color"].contiguous(), u_t["opacity"].contiguous()
    _lib.check(L.gb_bin_tiles_pack(G, xys.data_ptr(), depths.data_ptr(), radii.data_ptr(), conics.data_ptr(),
                                   col3.data_ptr(), op1.data_ptr(), comp.data_ptr(), H, W, BW, cap, bins.data_ptr(),
                                   order.data_ptr(), 0, gids.data_ptr(), rec.data_ptr(), None, ovf.data_ptr(),
                                   ws.data_ptr(), _lib.stream_ptr(cuda)), "bin_tiles_pack")
    torch.cuda.synchronize()
    assert int(ovf) == 0
    assert np.array_equal(t2n(bins), b["tile_bins"]), "tile_bins"
    assert np.array_equal(t2n(gids[:n]), b["gaussian_ids_sorted"]), "gaussian_ids_sorted"


def _step_variant(cuda, G, split, side_stream, graph=False, ranked=True):
    """bench.gpu_step with the sync-free render as one autograd node or two, shade on the main or on the side stream."""
    import bench
    from goliath_b200 import synthetic
    from goliath_b200.gsplat import fused

    li = {k: v.to(cuda) for k, v in synthetic.lights(8).items()}
    c = synthetic.ring_camera(1, img_h=bench.H, img_w=bench.W)
    cam = dict(Rt=c["viewmat"][None].to(cuda), intr=(c["fx"], c["fy"], c["cx"], c["cy"]))
    leaves = {k: v.detach().requires_grad_() for k, v in bench.unpack(bench.packed_scene(G).to(cuda)).items()}
    old = (fused.SPLIT, bench.SHADE_STREAM, fused.RANKED)
    fused.SPLIT, bench.SHADE_STREAM, fused.RANKED = split, side_stream, ranked
    try:
        cap = max(8 * G, 1 << 20)
        if graph:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):  # warm-up outside the capture (workspace allocation, lazy module loads)
                    for v in leaves.values():
                        v.grad = None
                    bench.gpu_step(leaves, cam, li, capacity=cap)
            torch.cuda.current_stream().wait_stream(s)
            for v in leaves.values():
                v.grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                rgb, alpha, depth = bench.gpu_step(leaves, cam, li, capacity=cap)
            g.replay()
        else:
            rgb, alpha, depth = bench.gpu_step(leaves, cam, li, capacity=cap)
        torch.cuda.synchronize()
    finally:
        fused.SPLIT, bench.SHADE_STREAM, fused.RANKED = old
    return t2n(rgb), t2n(alpha), t2n(depth), {k: t2n(v.grad) for k, v in leaves.items()}


@pytest.mark.parametrize("graph", [False, True])
def test_split_nodes_and_side_stream_match_single_node(cuda, graph):
    """The two-node render with the shade on a side stream (colours waited for inside gb_bin_tiles_pack_ev, projection
    backward beside the shade backward) against the single-node render on one stream: same kernels, same inputs —
    images identical, gradients equal up to the order of the atomic adds.  Also under CUDA-graph capture (fork/join)."""
    G = 60_000
    ref = _step_variant(cuda, G, split=False, side_stream=False)
    for split, side, ranked in ((True, False, True), (True, True, True), (True, True, False), (True, False, False),
                                (False, True, False)):
        got = _step_variant(cuda, G, split, side, graph=graph, ranked=ranked)
        for name, a, b in zip(("rgb", "alpha", "depth"), got[:3], ref[:3]):
            assert np.array_equal(a, b), (name, split, side, ranked)
        for k, want in ref[3].items():
            assert_close(got[3][k], want, rtol=1e-4, atol=1e-5 * float(np.abs(want).max()), frac=0.9999,
                         what="grad %s split=%s side=%s ranked=%s" % (k, split, side, ranked))
