"""Data-parallel gradient all-reduce overlapped with the backward pass (BASELINE config 5: "DDP gradient allreduce").

One process per GPU.  The gradients already live in one flat buffer (`training.optim.FlatBuffers`) laid out in
reverse parameter order, so a bucket is a contiguous slice; a post-accumulate-grad hook on every parameter counts the
bucket down and, when its last gradient has landed, launches an asynchronous `all_reduce` on that slice (NCCL runs it
on its own stream over NVLink while the remaining backward kernels execute).  `finish()` waits for the outstanding
buckets; the 1/world_size of the mean is folded into the optimiser kernel (`FusedAdamW.grad_scale`).

Backend-agnostic (`nccl` on the GPUs, `gloo` in the CPU tests)."""
from __future__ import annotations

from typing import List

import torch.distributed as dist


class BucketedAllReduce:
    def __init__(self, flat, *, bucket_bytes: int = 64 << 20, group=None):
        self.flat = flat
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.buckets: List[tuple] = []   # (begin, end, n_params)
        self.bucket_of = {}
        order = sorted(flat.params, key=lambda p: flat.offsets[id(p)])
        begin, count = 0, 0
        cap = max(1, bucket_bytes // 4)
        self.bucket_elems = cap
        for p in order:
            o, n = flat.slice_of(p)
            end = o + (n + 3) // 4 * 4
            self.bucket_of[id(p)] = len(self.buckets)
            count += 1
            if end - begin >= cap:
                self.buckets.append((begin, end, count))
                begin, count = end, 0
        if count:
            self.buckets.append((begin, flat.numel, count))
        self._pending = [0] * len(self.buckets)
        self._handles: List = []
        self._armed = False
        self._record = None              # set of id(param) while `record()` is active
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in flat.params]

    def arm(self) -> None:
        """Call before each backward that should be reduced (the last micro-batch of an accumulation cycle)."""
        self._pending = [b[2] for b in self.buckets]
        self._handles = []
        self._armed = True

    def _on_grad(self, p) -> None:
        if self._record is not None:
            self._record.add(id(p))
        if not self._armed or self.world == 1:
            return
        i = self.bucket_of[id(p)]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            b, e, _ = self.buckets[i]
            self._handles.append(dist.all_reduce(self.flat.grad[b:e], op=dist.ReduceOp.SUM, group=self.group,
                                                 async_op=True))

    def finish(self) -> None:
        """Wait for the reductions; any bucket whose hooks did not all fire (unused parameters) is reduced now."""
        if not self._armed:
            return
        self._armed = False
        if self.world == 1:
            return
        for i, left in enumerate(self._pending):
            if left > 0:
                b, e, _ = self.buckets[i]
                self._handles.append(dist.all_reduce(self.flat.grad[b:e], op=dist.ReduceOp.SUM, group=self.group,
                                                     async_op=True))
        for h in self._handles:
            h.wait()
        self._handles = []

    # ---- staged form (CUDA-graph replayed steps whose backward is cut in two, `training.graph.BackwardCut`) --------
    def record(self, on: bool):
        """While on, remember which parameters received their gradient (direct writes and accumulate-grad hooks both
        end in `_on_grad`); returns the recorded id set when switched off."""
        got, self._record = self._record, (set() if on else None)
        return got

    def ranges_of(self, param_ids) -> List[tuple]:
        """Maximal contiguous flat ranges [(begin, end)] covered by these parameters, split at `bucket` elements."""
        spans = []
        for p in self.flat.params:
            if id(p) in param_ids:
                o, n = self.flat.slice_of(p)
                spans.append((o, o + (n + 3) // 4 * 4))
        spans.sort()
        merged: List[list] = []
        for b, e in spans:
            if merged and merged[-1][1] == b:
                merged[-1][1] = e
            else:
                merged.append([b, e])
        out = []
        for b, e in merged:
            while e - b > self.bucket_elems:
                out.append((b, b + self.bucket_elems))
                b += self.bucket_elems
            out.append((b, e))
        return out

    def complement(self, ranges) -> List[tuple]:
        out, pos = [], 0
        for b, e in sorted(ranges):
            if b > pos:
                out.append((pos, b))
            pos = max(pos, e)
        if pos < self.flat.numel:
            out.append((pos, self.flat.numel))
        return out

    def launch(self, ranges) -> None:
        """Asynchronous all-reduce of these flat gradient ranges (NCCL's stream waits for the work queued so far on
        the current stream, then runs beside whatever is queued next); `wait()` joins them."""
        if self.world == 1:
            return
        for b, e in ranges:
            self._handles.append(dist.all_reduce(self.flat.grad[b:e], op=dist.ReduceOp.SUM, group=self.group,
                                                 async_op=True))

    def wait(self) -> None:
        for h in self._handles:
            h.wait()
        self._handles = []

    def reduce_all(self) -> None:
        """Non-overlapped form (after a CUDA-graph replay of the backward): every bucket, then wait."""
        if self.world == 1:
            return
        hs = [dist.all_reduce(self.flat.grad[b:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
              for b, e, _ in self.buckets]
        for h in hs:
            h.wait()

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
