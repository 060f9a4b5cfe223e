"""Multi-GPU plumbing for the relation head: images are independent (every index in rel_pair_idxs[i] refers to
boxes of image i only), so the path shards by image with NO data-path collective.  One process per
GPU under torch.distributed; the only exchange is an optional fixed-layout gather of the results, replacing the
reference's pickled all_gather (pysgg/utils/comm.py:47-90, engine/inference.py:49-53).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def pair_count(n: int, max_pairs: int = 1 << 62) -> int:
    return max(1, min(n * (n - 1), max_pairs))


def shard_images(n_boxes: Sequence[int], world_size: int, max_pairs: int = 1 << 62) -> List[List[int]]:
    """Greedy longest-processing-time partition of image indices over ranks, balanced by candidate-pair count
    (the work of the head is proportional to R_i = N_i(N_i-1), not to the image count).  Deterministic; every
    rank computes the same table, so no communication is needed."""
    order = sorted(range(len(n_boxes)), key=lambda i: (-pair_count(n_boxes[i], max_pairs), i))
    load = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += pair_count(n_boxes[i], max_pairs)
    return [sorted(s) for s in shards]


def gather_rows(local: torch.Tensor, group=None) -> List[torch.Tensor]:
    """All-gather of per-rank row blocks [R_rank, C] with different R_rank: one size exchange + one padded
    all_gather_into_tensor-style exchange of a fixed-layout tensor (no pickling)."""
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes + [1])
    padded = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return [o[:s] for o, s in zip(out, sizes)]


def max_over_ranks(value: float, device) -> float:
    """Timing reduction of the bench contract: the slowest rank defines the step."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_gradients(params: Sequence[torch.Tensor], group=None) -> None:
    """The data-parallel gradient exchange of the training step (reference: DistributedDataParallel's bucketed NCCL
    all-reduce, tools/relation_train_net.py:372-380): the gradients of all trained parameters (about 17.6 M fp32
    values for the relation head) travel as ONE flat bucket — over NVSwitch the cost is latency, not links — and come
    back averaged over ranks, like DDP."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
