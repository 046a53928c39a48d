"""All conv-weight packs of a training step in ONE launch.

Every step re-packs each fp32 master weight into the bf16 K-major matrix the forward implicit GEMM reads and into the
mirrored/transposed matrix its data-gradient conv reads (~240 small launches).  The set of packs is the same every step
and the master weights sit at fixed addresses (`FlatBuffers`), so the first step RECORDS the requests (packing each one
individually into a persistent buffer) and every later step refreshes all of those buffers with one
`fm_weight_prepack_batch_bf16` launch at the start of the forward pass; the per-conv requests then just look their
matrix up.  A request whose source moved (parameters re-seated, `.to()`) or that was never recorded invalidates the plan
and the step falls back to individual packs while a new plan is recorded."""
from __future__ import annotations

import weakref
from typing import Dict, Optional

import torch

from .. import _lib, ops
from ..ops import _stream

BF16 = torch.bfloat16


def _stable_source(w: torch.Tensor) -> bool:
    """fp32 contiguous storage of a Parameter (itself or a view of one): its address survives optimiser steps."""
    if getattr(w, "_fm_params", None) is not None:      # `functions.fused_param`: a view over adjacent flat-buffer slices
        return w.dtype == torch.float32 and w.is_contiguous()
    base = w._base if w._base is not None else w
    return (isinstance(base, torch.nn.Parameter) and w.dtype == torch.float32 and w.is_contiguous())


class PackPlan:
    def __init__(self):
        self.mats: Dict[tuple, tuple] = {}     # key -> (PackedConvWeight, source pointers)
        self.entries = []                      # _lib.PackEntry, in request order
        self.sources = {}                      # id(parameter) -> (weakref, data_ptr) of every master weight used
        self.ready = False
        self._tables = None

    # ---- recording / lookup -------------------------------------------------------------------------------
    def forward_matrix(self, segs, weights) -> Optional[ops.PackedConvWeight]:
        """segs: [(weight index, c_begin, c_count)]; returns the persistent packed matrix, or None (not plannable)."""
        ws = [weights[wi] for wi, _, _ in segs]
        if not all(_stable_source(w) for w in ws):
            return None
        key = ("f",) + tuple((w.data_ptr(), tuple(w.shape), cb, cc) for w, (_, cb, cc) in zip(ws, segs))
        hit = self.mats.get(key)
        if hit is not None:
            return hit
        if self.ready:
            self.invalidate()
        pw = ops.pack_conv_weight([(w, cb, cc) for w, (_, cb, cc) in zip(ws, segs)])
        ktot, koff = pw.mat.shape[1], 0
        for w, (_, cb, cc) in zip(ws, segs):
            self._track(w)
            ks = 1 if w.dim() == 2 else int(w.shape[-1])
            self.entries.append(_lib.PackEntry(w.data_ptr(), pw.mat.data_ptr(), ktot, koff, w.shape[0], w.shape[1], cb,
                                               cc, ks, 0))
            koff += ks * ks * cc
        self.mats[key] = pw
        return pw

    def dgrad_matrix(self, w: torch.Tensor, cb: int, cc: int, build) -> Optional[ops.PackedConvWeight]:
        if not _stable_source(w):
            return None
        key = ("d", w.data_ptr(), tuple(w.shape), cb, cc)
        hit = self.mats.get(key)
        if hit is not None:
            return hit
        if self.ready:
            self.invalidate()
        pw = build()
        self._track(w)
        ks = 1 if w.dim() == 2 else int(w.shape[-1])
        self.entries.append(_lib.PackEntry(w.data_ptr(), pw.mat.data_ptr(), 0, 0, w.shape[0], w.shape[1], cb, cc, ks, 1))
        self.mats[key] = pw
        return pw

    def _track(self, w: torch.Tensor) -> None:
        fused = getattr(w, "_fm_params", None)
        for base in (fused if fused is not None else (w._base if w._base is not None else w,)):
            self.sources[id(base)] = (weakref.ref(base), base.data_ptr())

    def _sources_moved(self) -> bool:
        for ref, ptr in self.sources.values():
            p = ref()
            if p is None or p.data_ptr() != ptr:
                return True
        return False

    def invalidate(self) -> None:
        self.mats, self.entries, self.sources, self.ready, self._tables = {}, [], {}, False, None

    # ---- per-step refresh ---------------------------------------------------------------------------------
    def begin_step(self, device) -> None:
        """Call at the start of every training forward: refresh every recorded matrix from the master weights."""
        if self.sources and self._sources_moved():  # parameters re-seated (optimiser created later, .to(), ...)
            self.invalidate()
        if not self.ready:
            # the first forward + backward after (re)starting records the requests; the next forward freezes the plan
            if not self.entries:
                return
            self._finalize(device)
        ent, blk_e, blk_o, n_blocks = self._tables
        _lib.check(_lib.lib().fm_weight_prepack_batch_bf16(ent.data_ptr(), blk_e.data_ptr(), blk_o.data_ptr(), n_blocks,
                                                           _stream()), "weight_prepack_batch")

    def _finalize(self, device) -> None:
        per_block = int(_lib.lib().fm_weight_prepack_batch_block_elems())
        arr = (_lib.PackEntry * len(self.entries))(*self.entries)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
        blk_e, blk_o = [], []
        for i, e in enumerate(self.entries):
            pairs = e.Cout * e.Cseg  # a thread packs the ksize^2 taps of one (matrix row, channel) pair
            for off in range(0, pairs, per_block):
                blk_e.append(i)
                blk_o.append(off)
        self._tables = (raw, torch.tensor(blk_e, dtype=torch.int32, device=device),
                        torch.tensor(blk_o, dtype=torch.int64, device=device), len(blk_e))
        self.ready = True
