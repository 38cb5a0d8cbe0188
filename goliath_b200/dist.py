"""View-sharded multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on B200, gloo in
the CPU tests).  The path partitions by view — given the decoded Gaussians of a frame, project/bin/sort/blend of
different views are independent — so the only exchange steps are a broadcast of the decoded table from the frame
owner and a sum-reduction of its gradient.  The reference is single-GPU only: nothing to match."""
from typing import List

import torch
import torch.distributed as dist


def views_of_rank(rank: int, world: int, n_views: int) -> List[int]:
    """Round-robin view partition: rank r renders views {v : v mod world == r}."""
    return list(range(rank, n_views, world))


def owner_broadcast(decoded: torch.Tensor, src: int = 0) -> torch.Tensor:
    """Broadcast the packed decoded-Gaussian table of a frame from its owner (in place; returns the tensor)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(decoded, src=src)
    return decoded


def reduce_grads(grad: torch.Tensor) -> torch.Tensor:
    """Sum dL/d(decoded table) over the ranks (each rank contributes its views); in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
    return grad
