"""CPU, world_size 2, gloo: the N>1 host logic of the view-sharded path — the frame owner broadcasts
the packed decoded-Gaussian table, every rank renders its own view, gradients w.r.t. the table are all-reduced.
The per-view render is replaced by the CPU oracle here (tests may use it); the collectives and the sharding are the
code under test (bench.py uses the same pattern with NCCL)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from goliath_b200 import synthetic
    from goliath_b200.dist import owner_broadcast, reduce_grads, views_of_rank

    G, H, W = 800, 48, 40
    table = torch.zeros(G, 14)
    if rank == 0:  # only the owner holds the decoded Gaussians
        sc = synthetic.head_gaussians(G, seed=3)
        table = torch.cat([sc["means3d"], sc["quats"], sc["scales"] * 12, sc["opacity"], sc["colors"]], 1).contiguous()
    table = owner_broadcast(table, src=0)
    views = views_of_rank(rank, world, 4)
    assert views == [rank, rank + 2]
    grad = torch.zeros_like(table)
    checksum = 0.0
    for v in views:
        c = synthetic.ring_camera(v, img_h=H, img_w=W)
        f = 300.0
        t = table.numpy()
        p = oracle.project_fwd(t[:, 0:3], t[:, 7:10], 1.0, t[:, 3:7], c["viewmat"].numpy(), f, f, W / 2, H / 2, H, W, 16, 0.1)
        b = oracle.bin_and_sort(p["xys"], p["depths"], p["radii"], p["num_tiles_hit"], H, W, 16)
        opac = t[:, 10] * p["compensation"]
        img, Ts, fi = oracle.rasterize_fwd(H, W, 16, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"],
                                           t[:, 11:14], opac, np.zeros(3, np.float32))
        g = oracle.rasterize_bwd(H, W, 16, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"], t[:, 11:14],
                                 opac, np.zeros(3, np.float32), Ts, fi, np.ones((H, W, 3), np.float32),
                                 np.zeros((H, W), np.float32))
        grad[:, 11:14] += torch.from_numpy(g[2])
        checksum += float(img.sum())
    local = float(grad.sum())
    total = reduce_grads(grad)  # in place
    q.put((rank, float(table.sum()), checksum, float(total.sum()), local))
    dist.destroy_process_group()


def test_view_shard_broadcast_and_grad_allreduce():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    (_, t0, c0, tot0, g0), (_, t1, c1, tot1, g1) = res
    assert t0 == t1 and t0 != 0.0, "both ranks must hold the owner's table after the broadcast"
    assert c0 != c1, "ranks render different views"
    assert abs(tot0 - tot1) < 1e-3 * abs(tot0) and abs(tot0 - (g0 + g1)) < 1e-3 * abs(tot0), "all-reduce(sum) of grads"


def test_view_partition_of_the_baseline_configs():
    """BASELINE.json configs[2] / configs[4]: 16 views over 8 GPUs -> 2 per rank; 150 cameras over
    8 GPUs -> 19 on six ranks, 18 on two; every view exactly once."""
    from goliath_b200.dist import views_of_rank

    parts = [views_of_rank(r, 8, 16) for r in range(8)]
    assert all(len(p) == 2 for p in parts) and sorted(sum(parts, [])) == list(range(16))
    parts = [views_of_rank(r, 8, 150) for r in range(8)]
    assert sorted(len(p) for p in parts) == [18, 18, 19, 19, 19, 19, 19, 19]
    assert sorted(sum(parts, [])) == list(range(150))
    assert views_of_rank(0, 1, 5) == [0, 1, 2, 3, 4] and views_of_rank(3, 4, 2) == []


def _exchange_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from goliath_b200.dist import FrameExchange

    n = 1000
    ex = FrameExchange(n, "cpu", owner=0)
    frames = [torch.full((n,), float(i + 1)) for i in range(3)] if rank == 0 else [None] * 3  # only the owner has the tables
    static_in = torch.empty(n)
    seen, sums = [], []
    ex.post_input(0, frames[0])
    for i in range(3):
        ex.take_input(i, static_in)
        if i + 1 < 3:
            ex.post_input(i + 1, frames[i + 1])
        seen.append(float(static_in[0]))
        ex.post_grad(i, static_in * (rank + 1))        # rank r contributes (r+1) * table
        sums.append(float(ex.grad(i)[0]))
    ex.finish()
    # zero-copy variant: the step reads ex.inputs[i & 1] and writes ex.grads[i & 1] in place
    ex2 = FrameExchange(n, "cpu", owner=0)
    seen2, sums2 = [], []
    ex2.post_input(0, frames[0])
    for i in range(3):
        ex2.wait_input(i)
        if i + 1 < 3:
            ex2.post_input(i + 1, frames[i + 1])
        seen2.append(float(ex2.inputs[i & 1][0]))
        ex2.grads[i & 1].copy_(ex2.inputs[i & 1] * (rank + 1))
        ex2.release_input(i)
        ex2.post_grad_inplace(i)
        sums2.append(float(ex2.grad(i)[0]))
    ex2.finish()
    assert seen2 == seen and sums2 == sums, (seen2, sums2)
    bad = None
    if rank != 0:
        try:
            FrameExchange(n, "cpu", owner=1).post_input(0, None)
        except RuntimeError as e:
            bad = str(e)
    q.put((rank, seen, sums, ex.bytes_h2d, bad))
    dist.destroy_process_group()


def test_frame_exchange_owner_only_source_and_reduced_grads():
    """goliath_b200.dist.FrameExchange (the product API bench.py drives): non-owner ranks never provide a table, every
    rank receives the owner's, the gradient of every frame is the sum over the ranks."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, seen, sums, h2d, bad in res:
        assert seen == [1.0, 2.0, 3.0]
        assert sums == [3.0, 6.0, 9.0]                    # (1 + 2) * table
        assert h2d == (3 * 1000 * 4 if rank == 0 else 0)  # only the owner copies host bytes in
    assert res[1][4] is not None and "owner" in res[1][4]
