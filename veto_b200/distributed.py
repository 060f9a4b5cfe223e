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


def _flat_buckets(grads: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """The gradients grouped by the storage they live in: the relation head hands autograd VIEWS of one flat buffer
    (predictor._TrainStep), the depth backbone views of another (ops.depth_backbone_backward), so the exchange can sweep
    each buffer in place.  Returns one 1-D alias per storage covering the span its gradients occupy (padding between views
    is reduced along with them: harmless); a gradient that is not contiguous comes back as is."""
    spans = {}
    loose = []
    for g in grads:
        if not g.is_contiguous():
            loose.append(g)
            continue
        key = (g.untyped_storage().data_ptr(), g.dtype, g.device)
        lo, hi = g.storage_offset(), g.storage_offset() + g.numel()
        if key in spans:
            spans[key][1] = min(spans[key][1], lo)
            spans[key][2] = max(spans[key][2], hi)
        else:
            spans[key] = [g, lo, hi]
    out = loose
    for g, lo, hi in spans.values():
        out.append(torch.empty(0, dtype=g.dtype, device=g.device).set_(g.untyped_storage(), lo, (hi - lo,)))
    return out


def allreduce_gradients(params: Sequence[torch.Tensor], group=None, async_op: bool = False):
    """The data-parallel gradient exchange of the training step (reference: DistributedDataParallel's bucketed NCCL
    all-reduce of the relation head + depth backbone gradients, tools/relation_train_net.py:372-380), averaged over ranks
    like DDP.  The gradients of this library already live in one flat buffer per module, so each buffer is all-reduced IN
    PLACE — no concatenation, no copy back (over NVSwitch the cost is latency, not links).  async_op=True returns the NCCL
    work handles (finish with ``finish_gradient_sync``) so that the exchange of a module whose gradients are ready can run
    under the backward pass of the next one."""
    if not (dist.is_available() and dist.is_initialized()):
        return []
    world = dist.get_world_size(group)
    if world == 1:
        return []
    grads = [p.grad if isinstance(p, torch.nn.Parameter) or getattr(p, "grad", None) is not None else None for p in params]
    grads = [g for g in grads if g is not None]
    works = []
    for flat in _flat_buckets(grads):
        if hasattr(dist.ReduceOp, "AVG") and flat.is_cuda:
            works.append((dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group, async_op=True), None, world))
        else:   # gloo (CPU tests): sum, then divide when the work is finished
            works.append((dist.all_reduce(flat, group=group, async_op=True), flat, world))
    if async_op:
        return works
    finish_gradient_sync(works)
    return []


def allreduce_flat(flat: torch.Tensor, group=None):
    """Start the in-place averaging all-reduce of one flat gradient buffer; returns work handles for finish_gradient_sync
    (an empty list without a process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    world = dist.get_world_size(group)
    if hasattr(dist.ReduceOp, "AVG") and flat.is_cuda:
        return [(dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group, async_op=True), None, world)]
    return [(dist.all_reduce(flat, group=group, async_op=True), flat, world)]


def finish_gradient_sync(works) -> None:
    """Wait for the all-reduces started with async_op=True (the wait only orders the current stream behind NCCL's)."""
    for work, flat, world in works:
        work.wait()
        if flat is not None:
            flat.div_(world)
