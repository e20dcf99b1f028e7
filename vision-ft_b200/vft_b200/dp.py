"""Data-parallel plumbing for the QLoRA step: bucketed sum-allreduce of the LoRA gradients.

The reference shards by batch through HF Accelerate -> torch DDP
(/root/reference/src/trainer/common.py:198 ``accelerator.prepare(self.model)``,
``no_sync`` on non-final accumulation steps :302-308).  The NF4 base is frozen and
replicated; the only exchange per optimizer step is the sum of the adapter
gradients (21-34 M bf16 parameters for AuraFlow at r = 16).

``LoraGradReducer`` registers post-accumulate-grad hooks on the trainable (adapter)
parameters, packs gradients into flat buckets in reverse registration order (the
order backward produces them), and launches one NCCL all-reduce per full bucket on
a side stream as soon as its last gradient lands, so the transfers overlap the
remaining backward kernels.  ``wait()`` joins the side stream and scatters the
averaged values back before the optimizer step.  Packing and unpacking are ONE
multi-tensor copy per bucket (``torch._foreach_copy_``) and the average is taken by
NCCL itself (``ReduceOp.AVG``): with 288 adapter tensors in an AuraFlow step, a copy
kernel per tensor in each direction was 11.5 ms of exposed time per step at 2 GPUs.
On CPU tensors (gloo; used by the world_size-2 tests) the same logic runs without
streams, with SUM followed by a division (gloo has no AVG).
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Iterable

import torch
import torch.distributed as dist


class _Bucket:
    def __init__(self, params: list[torch.nn.Parameter]):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        p0 = params[0]
        self.flat = torch.zeros(self.numel, dtype=p0.dtype, device=p0.device)
        self.pending = len(params)
        self.work = None
        self.event = None
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += p.numel()
        self.views = [self.flat[o : o + p.numel()].view(p.shape) for o, p in zip(self.offsets, params)]


class LoraGradReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 8 << 20, average: bool = True,
                 group=None, overlap: bool = True):
        # overlap=False: hooks only count; every bucket is reduced from wait(), after backward.  The fused GEMMs are
        # persistent kernels sized to the whole GPU: a collective that holds a few SMs while they launch costs them a
        # wave (measured: 11 ms per AuraFlow step at 2-4 GPUs), far more than the 67 MB all-reduce itself.
        self.overlap = overlap
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.average = average
        self.enabled = True
        plist = [p for p in params if p.requires_grad]
        # backward visits layers last-to-first: build buckets in that order so they fill contiguously in time
        plist = list(reversed(plist))
        self.buckets: list[_Bucket] = []
        cur: list[torch.nn.Parameter] = []
        cur_bytes = 0
        for p in plist:
            same = not cur or (cur[0].dtype == p.dtype and cur[0].device == p.device)
            if cur and (not same or cur_bytes + p.numel() * p.element_size() > bucket_bytes):
                self.buckets.append(_Bucket(cur))
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
        if cur:
            self.buckets.append(_Bucket(cur))
        self._where = {}
        for b in self.buckets:
            for i, p in enumerate(b.params):
                self._where[id(p)] = (b, i)
        self._cuda = bool(self.buckets) and self.buckets[0].flat.is_cuda
        self.stream = torch.cuda.Stream(device=self.buckets[0].flat.device) if self._cuda else None
        self._next = 0  # first bucket that has not been launched in this step
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in plist]

    # ------------------------------------------------------------------ hooks
    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self.enabled or self.world == 1:
            return
        b, _ = self._where[id(p)]
        b.pending -= 1
        # Collectives leave in BUCKET ORDER on every rank: a bucket that is complete is launched only once all buckets
        # before it have left (the order in which buckets complete can differ between ranks -- a branch not taken on
        # one of them -- and ranks issuing the same collectives in different orders deadlock or mix up buffers).
        while self.overlap and self._next < len(self.buckets) and self.buckets[self._next].pending == 0:
            self._launch(self.buckets[self._next])
            self._next += 1

    def _pack(self, b: _Bucket) -> None:
        """All gradients of the bucket -> its flat buffer, one multi-tensor copy (missing gradients count as zero)."""
        have = [(v, p.grad) for v, p in zip(b.views, b.params) if p.grad is not None]
        if len(have) != len(b.params):
            b.flat.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])

    def _launch(self, b: _Bucket) -> None:
        self._pack(b)
        op = dist.ReduceOp.AVG if (self._cuda and self.average) else dist.ReduceOp.SUM
        if self._cuda:
            b.event = torch.cuda.Event()
            b.event.record(torch.cuda.current_stream(b.flat.device))
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(b.event)
                b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)
        else:
            b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    # ------------------------------------------------------------------ API
    @contextmanager
    def no_sync(self):
        """Gradient-accumulation micro-steps: accumulate locally, exchange nothing (trainer/common.py:302-308)."""
        prev, self.enabled = self.enabled, False
        try:
            yield
        finally:
            self.enabled = prev

    def wait(self) -> None:
        """Block the current stream until every bucket is reduced; write the (averaged) sums back into .grad.

        EVERY bucket is reduced on EVERY rank at each call, whether or not a gradient reached it on this rank (as DDP
        does): which parameters receive a gradient can differ between ranks (a branch not taken, an empty ragged
        bucket), and a rank-local decision to skip a bucket would make the ranks issue different collective sequences.
        A parameter without a local gradient contributes zeros and RECEIVES the average (its .grad is created), so
        the replicas stay identical."""
        if self.world == 1:
            return
        for b in self.buckets[self._next:]:  # first pass: everything that has not left yet leaves now, in bucket order
            # deferred mode, or a bucket some (or all) of whose parameters received no gradient: they count as zero
            self._launch(b)
        self._next = 0
        for b in self.buckets:  # second pass: join and write back
            b.work.wait()
            if self._cuda:
                torch.cuda.current_stream(b.flat.device).wait_stream(self.stream)
            if self.average and not self._cuda:
                b.flat.div_(self.world)
            for v, p in zip(b.views, b.params):
                if p.grad is None:
                    p.grad = torch.empty_like(p)
            torch._foreach_copy_([p.grad for p in b.params], list(b.views))
            b.pending = len(b.params)
            b.work = None

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
