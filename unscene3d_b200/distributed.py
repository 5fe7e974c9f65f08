"""Data-parallel plumbing: scenes shard one-per-rank with no data-path collective; the
only exchanges of a training step are the gradient all-reduce (the reference gets it from Lightning's DDP,
main_instance_segmentation.py:86-93) and the scalar `num_masks` all-reduce (models/criterion.py:258-260).

torch.distributed (NCCL over NVLink on the B200 box, gloo in CPU tests) carries both; gradients travel as a few
large flat buckets in reverse registration order (the order backward produces them), averaged over ranks.
"""
from typing import Iterable, List

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def scene_seed(base_seed: int, step: int, rank: int, world_size: int, scenes_per_rank: int = 1, slot: int = 0) -> int:
    """Deterministic, disjoint scene ids: global scene index of (step, rank, slot)."""
    return base_seed + (step * world_size + rank) * scenes_per_rank + slot


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the maximum over ranks (bench contract)."""
    rank, world = world_info()
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _buckets(params: List[torch.nn.Parameter], bucket_bytes: int):
    bucket, size = [], 0
    for p in params:
        bucket.append(p)
        size += p.grad.numel() * p.grad.element_size()
        if size >= bucket_bytes:
            yield bucket
            bucket, size = [], 0
    if bucket:
        yield bucket


def allreduce_gradients(parameters: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, async_op: bool = False):
    """Average .grad over all ranks.  Parameters are walked in REVERSE registration order (decoder first, stem last
    — the order in which backward finishes them) and flattened into ~bucket_bytes buckets, one all-reduce each;
    ~158 MB of fp32 gradients for Mask3D + Res16UNet34C = 3 buckets.  Returns the work handles when async_op."""
    rank, world = world_info()
    if world == 1:
        return []
    params = [p for p in reversed(list(parameters)) if p.grad is not None]
    handles = []
    for bucket in _buckets(params, bucket_bytes):
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
        handles.append((work, flat, bucket))
    if async_op:
        return handles
    finish_allreduce(handles)
    return []


def finish_allreduce(handles):
    _, world = world_info()
    for work, flat, bucket in handles:
        work.wait()
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
