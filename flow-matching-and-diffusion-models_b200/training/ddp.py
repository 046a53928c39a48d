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
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in flat.params]

    def arm(self) -> None:
        """Call before each backward that should be reduced (the last micro-batch of an accumulation cycle)."""
        self._pending = [b[2] for b in self.buckets]
        self._handles = []
        self._armed = True

    def _on_grad(self, p) -> None:
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
