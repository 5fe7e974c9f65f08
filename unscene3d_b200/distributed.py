"""Data-parallel plumbing: scenes shard one-per-rank with no data-path collective; the
only exchanges of a training step are the gradient all-reduce (the reference gets it from Lightning's DDP,
main_instance_segmentation.py:86-93) and the scalar `num_masks` all-reduce (models/criterion.py:258-260).

torch.distributed (NCCL over NVLink / NVSwitch on the B200 box, gloo in the CPU tests) carries both.

GradientReducer is the DDP half of the step:

  * the parameter list is walked in REVERSE registration order (the order in which backward finishes them: decoder first,
    stem last) and cut into flat buckets per dtype; the list is the model's, not "whoever has a gradient on this rank", so
    every rank builds the same buckets (a parameter without a gradient contributes zeros);
  * buckets are allocated once; after finish() `.grad` of every parameter is a VIEW into its bucket — no torch.cat before
    the collective, no copy back after it.  Gradients enter a bucket either by ONE multi-tensor copy when the bucket is
    complete (default: autograd keeps handing out fresh gradient tensors, so backward pays no per-parameter accumulate
    kernel) or, with `as_views=True`, by autograd accumulating in place into the views (gradient accumulation over several
    backward passes);
  * a post-accumulate-grad hook counts a bucket's parameters down and launches its all-reduce the moment the last one has
    its gradient (buckets go out in index order, so every rank issues the same sequence of collectives) — the collective
    runs on the communication stream of the process group while backward keeps producing the
    next bucket on the compute stream;
  * finish() launches whatever did not fill (unused parameters contribute zeros), makes the compute stream wait for the
    collectives and applies the 1 / world scale (fused into the collective as ReduceOp.AVG on NCCL).
"""
from typing import Iterable, List, Optional

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


def all_reduce_scalar(value, device=None, op=None) -> float:
    """The criterion's `num_masks` exchange (models/criterion.py:255-260): one scalar, summed over ranks."""
    rank, world = world_info()
    t = torch.as_tensor([float(value)], dtype=torch.float32, device=device)
    if world > 1:
        dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
    return float(t.item())


class _Bucket:
    __slots__ = ("flat", "params", "pending", "work", "launched", "nbytes", "views", "fired")

    def __init__(self, flat, params):
        self.flat, self.params = flat, params
        self.pending, self.work, self.launched = len(params), None, False
        self.nbytes = flat.numel() * flat.element_size()
        self.views, self.fired = [], set()


class GradientReducer:
    """Bucketed gradient averaging overlapped with backward (see the module docstring).

        reducer = GradientReducer(net.parameters())
        loss.backward()            # buckets all-reduce as they fill
        reducer.finish()           # compute stream now sees averaged .grad
        optimizer.step(); reducer.zero_grad()

    Works for world size 1 too (no collective; .grad still lives in the buckets)."""

    def __init__(self, parameters: Iterable[torch.nn.Parameter], bucket_bytes: int = 32 << 20, group=None, overlap: bool = True,
                 as_views: bool = False):
        self.rank, self.world = world_info()
        self.group, self.overlap, self.as_views = group, bool(overlap), bool(as_views)
        params = [p for p in parameters if p.requires_grad]
        # reverse registration order, stable per dtype / device: identical on every rank
        self.buckets: List[_Bucket] = []
        self._bucket_of = {}
        groups = {}
        for p in reversed(params):
            groups.setdefault((p.dtype, p.device), []).append(p)
        for (dtype, device), plist in groups.items():
            cur, size = [], 0
            for p in plist:
                cur.append(p)
                size += p.numel() * p.element_size()
                if size >= bucket_bytes:
                    self._make_bucket(cur, dtype, device)
                    cur, size = [], 0
            if cur:
                self._make_bucket(cur, dtype, device)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        backend = dist.get_backend(group) if self.world > 1 else None
        self._avg = backend == "nccl"
        self.enabled = True
        self._next = 0  # first bucket not launched yet in this backward pass

    def _make_bucket(self, plist, dtype, device):
        total = sum(p.numel() for p in plist)
        flat = torch.zeros(total, dtype=dtype, device=device)
        b = _Bucket(flat, list(plist))
        b.views = list(self._views(b))
        if self.as_views:
            for p, v in zip(plist, b.views):
                p.grad = v  # autograd accumulates into an existing .grad in place
        for p in plist:
            self._bucket_of[p] = b
        self.buckets.append(b)

    # ---- per step
    def _on_grad(self, p):
        b = self._bucket_of.get(p)
        if b is None:
            return
        b.pending -= 1
        b.fired.add(id(p))
        if b.pending == 0 and self.overlap:
            # collectives must be issued in the same order on every rank: bucket i goes out only after buckets 0..i-1 (a
            # bucket that never fills on some rank — an unused branch — holds the later ones back until finish())
            while self._next < len(self.buckets) and self.buckets[self._next].pending == 0:
                self._launch(self.buckets[self._next])
                self._next += 1

    def _gather(self, b: _Bucket):
        """Copy mode: bring the fresh gradient tensors of a bucket into its flat buffer with one multi-tensor copy."""
        if self.as_views:
            return
        src, dst = [], []
        complete = len(b.fired) == len(b.params)
        if not complete:
            b.flat.zero_()  # parameters without a gradient on this rank contribute zeros
        for p, v in zip(b.params, b.views):
            if id(p) in b.fired and p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)

    def _launch(self, b: _Bucket):
        if b.launched:
            return
        b.launched = True
        self._gather(b)
        if self.world == 1 or not self.enabled:
            return
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        # async_op: the collective is queued on the process group's stream behind everything the calling (backward) stream
        # has queued so far — i.e. behind the kernels that produced this bucket — and runs beside what backward queues next
        b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def finish(self):
        """After backward: launch the buckets that never filled, wait for every collective, scale."""
        for b in self.buckets:
            self._launch(b)
        for b in self.buckets:
            if b.work is not None:
                b.work.wait()  # NCCL: the current stream waits for the collective (no host block); gloo: host wait
                b.work = None
                if not self._avg:
                    b.flat.div_(self.world)
            if not self.as_views:
                for p, v in zip(b.params, b.views):
                    p.grad = v  # averaged gradient, a view of the bucket
            b.pending, b.launched = len(b.params), False
            b.fired.clear()
        self._next = 0

    def zero_grad(self):
        """Copy mode: drop the views (autograd hands out fresh tensors next time).  View mode: gradients stay views of the
        buckets and are zeroed in place (one memset per bucket)."""
        for b in self.buckets:
            if not self.as_views:
                for p in b.params:
                    p.grad = None
                continue
            b.flat.zero_()
            for p, view in zip(b.params, b.views):
                if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                    p.grad = view  # somebody called zero_grad(set_to_none=True): re-attach

    def _views(self, b):
        off = 0
        for p in b.params:
            n = p.numel()
            yield b.flat[off:off + n].view_as(p)
            off += n

    @property
    def total_bytes(self) -> int:
        return sum(b.nbytes for b in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def allreduce_gradients(parameters: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, async_op: bool = False):
    """One-shot form (no overlap): average the existing .grad tensors over all ranks after backward.  Every rank walks the
    same parameter list; a parameter without a gradient on this rank contributes zeros (and receives the average), so the
    buckets have the same layout everywhere.  Buckets are per dtype.  Kept for callers that own their .grad tensors; the
    training step uses GradientReducer."""
    rank, world = world_info()
    if world == 1:
        return []
    params = [p for p in reversed(list(parameters)) if p.requires_grad]
    by_dtype = {}
    for p in params:
        by_dtype.setdefault(p.dtype, []).append(p)
    handles = []
    for dtype, plist in by_dtype.items():
        bucket, size = [], 0
        for p in plist + [None]:
            if p is not None:
                bucket.append(p)
                size += p.numel() * p.element_size()
            if bucket and (p is None or size >= bucket_bytes):
                flat = torch.cat([(q.grad if q.grad is not None else torch.zeros_like(q)).reshape(-1) for q in bucket])
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
                handles.append((work, flat, bucket))
                bucket, size = [], 0
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
            n = p.numel()
            piece = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = piece.clone()
            else:
                p.grad.copy_(piece)
            off += n
