"""Differentiable forms of the B200 ops: `torch.autograd.Function`s whose forward AND backward are the hand-written
kernels (csrc/train.cu + the forward kernels), so `loss.backward()` of the reference's training step
(`src/pipelines/train/flow_matching_lib.py:138-182`) runs on them.  torch.autograd is used as the graph engine only
(gradient routing / accumulation); there is no eager or cuDNN fallback inside these functions.

Activations are bf16 channels_last, parameters fp32 (bf16-autocast semantics: fp32 master weights, bf16 matmul
inputs, fp32 accumulation; `flow_matching_lib.py:158-164`)."""
from __future__ import annotations

import contextlib
import math
from typing import Optional, Sequence

import torch
from torch.autograd import Function

from .. import _lib, ops
from ..ops import _ptr, _stream

BF16 = torch.bfloat16


# Direct parameter gradients.  The trainers keep every `.grad` as a view of one flat, pre-zeroed buffer
# (`training.optim.FlatBuffers`).  With DIRECT_PARAM_GRADS set (one backward per optimiser step, every parameter used
# once per step - true for both denoisers), the backward kernels write parameter gradients STRAIGHT into those views
# and the Functions return None for them: autograd's AccumulateGrad - one ATen `add` launch per parameter, ~450 per
# step, plus a temporary of the gradient's size - disappears from the step.  GRAD_READY_HOOK (the bucketed all-reduce's
# countdown) is then called by hand, since no accumulate-grad hook fires for a None gradient.
DIRECT_PARAM_GRADS = False
GRAD_READY_HOOK = None


def fused_param(params):
    """ONE tensor over several parameters that lie back to back in memory, or None.

    The trainers lay the flat parameter buffer out so that parameters which the step uses as one matrix - the
    `emb_layers` projections of all ResBlocks of a backward stage, the to_q / to_k / to_v weights of an attention block
    - are adjacent (`training.graph.flat_param_order`).  The concatenation is then a VIEW of the flat buffer (no
    `torch.cat` per step) and its gradient a view of the flat gradient buffer, which the backward kernels write
    directly (no per-parameter slice + AccumulateGrad).  Only under DIRECT_PARAM_GRADS (the trainers hold it over the
    forward AND the backward): the view is detached from autograd, its gradient exists only as that direct write."""
    if not DIRECT_PARAM_GRADS or not params:
        return None
    p0 = params[0]
    inner = tuple(p0.shape[1:])
    wp, gp, rows = p0.data_ptr(), None if p0.grad is None else p0.grad.data_ptr(), 0
    for q in params:
        g = q.grad
        if (not isinstance(q, torch.nn.Parameter) or not q.requires_grad or q.dtype != torch.float32
                or not q.is_contiguous() or tuple(q.shape[1:]) != inner or q.numel() % 4 or g is None
                or g.dtype != torch.float32 or not g.is_contiguous()
                or q.data_ptr() != wp or g.data_ptr() != gp):
            return None
        wp += 4 * q.numel()
        gp += 4 * q.numel()
        rows += q.shape[0]
    shape = (rows,) + inner
    strides = []
    acc = 1
    for d in reversed(shape):
        strides.append(acc)
        acc *= d
    strides = tuple(reversed(strides))
    w = torch.as_strided(p0.data, shape, strides)
    w._fm_grad_view = torch.as_strided(p0.grad, shape, strides)
    w._fm_params = tuple(params)
    return w


def _is_fused(t) -> bool:
    return getattr(t, "_fm_grad_view", None) is not None


def _grad_target(param, shape=None):
    """The `.grad` view of `param` to write into directly, or None (autograd accumulates the returned gradient)."""
    if _is_fused(param):
        if not DIRECT_PARAM_GRADS:
            raise RuntimeError("fmdm_b200.training: a forward that ran under DIRECT_PARAM_GRADS (fused parameter "
                               "views) must run its backward under it too")
        return param._fm_grad_view
    if not DIRECT_PARAM_GRADS or not isinstance(param, torch.nn.Parameter) or not param.requires_grad:
        return None
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() % 16:
        return None
    if shape is not None and tuple(g.shape) != tuple(shape):
        return None
    return g


def _grad_written(param) -> None:
    if GRAD_READY_HOOK is not None:
        for q in getattr(param, "_fm_params", (param,)):
            GRAD_READY_HOOK(q)


# The Cin <= 8 stem's weight gradient and the Cout = 1 head's data / weight gradients as tensor-core GEMMs over an im2col
# matrix (the forward stem's own trick).  FMDM_TRAIN_SMALL_CUDA=1 selects the first-generation CUDA-core kernels (A/B).
import os as _os

SMALL_CONVS_ON_TENSOR_CORES = _os.environ.get("FMDM_TRAIN_SMALL_CUDA") is None

_TICKETS = {}
_IN_SIDE = False


def _tickets(device) -> torch.Tensor:
    """The per-device (and per-stream: main / side) ticket buffer of the "last block finishes the job" kernels
    (`fm_ticket_ints`): zeroed once, left zero by every kernel, used by one stream at a time.  The trainers create
    both up front (`ensure_tickets`) so that none is first allocated inside a CUDA-graph capture."""
    key = (device.type, device.index, _IN_SIDE)
    t = _TICKETS.get(key)
    if t is None:
        t = _TICKETS[key] = torch.zeros(int(_lib.lib().fm_ticket_ints()), dtype=torch.int32, device=device)
    return t


def ensure_tickets(device) -> None:
    global _IN_SIDE
    was = _IN_SIDE
    try:
        for _IN_SIDE in (False, True):
            _tickets(device)
    finally:
        _IN_SIDE = was


# Side stream for work nothing downstream waits on (parameter gradients of small layers, see _ConvFn.backward).  The
# trainers switch it on while they run / capture a step and join it (`join_side_streams`) before anything reads the
# flat gradient buffer.  Tensors handed to the side stream are `record_stream`ed, so the caching allocator does not
# give their memory to a later main-stream allocation while the side branch still reads it.
SIDE_STREAMS = False
# per tensor: 32 Mi elements = 64 MB of bf16 - everything below the full-resolution level of LDCT-256 at B = 16 (measured
# at that config: 4 Mi no gain, 32 Mi -0.7 ms per step, 256 Mi -0.6 ms; the largest layers saturate the GPU on their own)
SIDE_STREAM_MAX_ELEMS = int(_os.environ.get("FMDM_SIDE_MAX_ELEMS", 1 << 25))
_SIDE = {}
_SIDE_USED = set()


@contextlib.contextmanager
def _side_branch(tensors):
    global _IN_SIDE
    dev = tensors[0].device
    key = (dev.type, dev.index)
    side = _SIDE.get(key)
    if side is None:
        side = _SIDE[key] = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    was, _IN_SIDE = _IN_SIDE, True
    try:
        with torch.cuda.stream(side):
            yield
    finally:
        _IN_SIDE = was
    for t in tensors:
        t.record_stream(side)
    _SIDE_USED.add(key)


def join_side_streams() -> None:
    """The current stream waits for every side branch opened since the last join."""
    for key in list(_SIDE_USED):
        dev = torch.device(key[0], key[1])
        torch.cuda.current_stream(dev).wait_stream(_SIDE[key])
    _SIDE_USED.clear()


# Gradient slots.  A tensor with several consumers (a ResBlock input feeds its GroupNorm AND its residual / skip
# conv; an encoder output feeds the next block AND a decoder block) would have its gradient summed by autograd with one
# ATen `add` launch per extra consumer.  Instead the FIRST consumer in forward order - always a GroupNorm or a conv
# here, and always the last of the consumers to run backward, because every other consumer sits downstream of it -
# opens a slot on the tensor; later consumers park their contribution in the slot (returning None to autograd) and
# the first consumer's backward kernel adds the parked tensors while it writes its own dx (`add0_a/b`, `add1` of
# fm_groupnorm_bwd_bf16, `residual` of the dgrad conv).  The result is bit-for-bit a sum of the same bf16 terms, in
# fp32, rounded once.
FUSE_GRAD_ACCUMULATION = True


class _Slot:
    __slots__ = ("parked", "expected", "contributed", "closed")

    def __init__(self):
        self.parked, self.expected, self.contributed, self.closed = [], 0, 0, False

    def pop_one(self):
        """A parked tensor for a parking consumer to fold into its own kernel (keeps the parked list short)."""
        return self.parked.pop() if self.parked else None

    def park(self, g: torch.Tensor) -> None:
        if self.closed:
            raise RuntimeError("fmdm_b200.training: a gradient was parked after its accumulating consumer ran backward "
                               "(consumer order violated); set training.functions.FUSE_GRAD_ACCUMULATION = False")
        self.parked.append(g)
        self.contributed += 1

    def take(self, limit: int):
        """The parked gradients, folded to at most `limit` tensors (more than that is rare: one ATen add each)."""
        if self.contributed != self.expected:
            raise RuntimeError(f"fmdm_b200.training: {self.expected} consumers registered on a gradient slot but "
                               f"{self.contributed} contributed before the accumulating backward ran")
        self.closed = True
        got, self.parked = self.parked, []
        while len(got) > limit:
            got = [got[0] + got[1]] + got[2:]
        return got


_OPEN_SLOTS = []


def assert_slots_drained() -> None:
    """After a backward pass: every parked gradient must have been consumed (a slot whose accumulating consumer never
    ran backward would silently drop gradients)."""
    left = [s for s in _OPEN_SLOTS if s.parked]
    del _OPEN_SLOTS[:]
    if left:
        raise RuntimeError(f"fmdm_b200.training: {len(left)} gradient slot(s) still hold parked gradients after the "
                           "backward pass; set training.functions.FUSE_GRAD_ACCUMULATION = False")


def _slot_eligible(t: torch.Tensor) -> bool:
    """Only tensors made inside the current forward carry slots: an activation (it has a grad_fn and belongs to one
    autograd graph) or a `BackwardCut` leaf.  A long-lived leaf (a parameter, an input reused across forwards) could
    be consumed by several independent backward passes, where a parked gradient would have no one to collect it."""
    return t.requires_grad and (t.grad_fn is not None or getattr(t, "_fm_fresh_leaf", False))


def _slot_open(t: torch.Tensor):
    """Called by the first consumer of `t`: returns the slot its backward must drain (None if fusion is off)."""
    if not FUSE_GRAD_ACCUMULATION or not _slot_eligible(t):
        return None
    slot = _Slot()
    t._fm_slot = slot
    _OPEN_SLOTS.append(slot)
    return slot


def _slot_join(t: torch.Tensor):
    """Called by a later consumer of `t`: the slot to park into, or None (autograd accumulates as usual)."""
    slot = getattr(t, "_fm_slot", None) if FUSE_GRAD_ACCUMULATION else None
    if slot is not None and slot.closed:   # a previous backward already drained it: not part of this graph
        slot = None
    if slot is not None:
        slot.expected += 1
    return slot


def _ws(n: int, device, dtype=torch.float32) -> torch.Tensor:
    return torch.empty((max(int(n), 1),), dtype=dtype, device=device)


def _rows_f32(t: torch.Tensor) -> torch.Tensor:
    """fp32 2-D tensor with unit inner stride and 16-byte aligned rows (column slices of a wider matrix pass through)."""
    t = t.detach()
    if t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0:
        return t
    return t.float().contiguous()


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    return ops.to_nhwc_bf16(t)


# --------------------------------------------------------------------------------------------------------------
# raw kernel wrappers
# --------------------------------------------------------------------------------------------------------------
def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, *, ksize: int, stride: int, c_begin: int) -> None:
    """Weight gradient of one source: writes dw[:, c_begin:c_begin+C_x] (fp32 OIHW, or [O][I] for 1x1; overwritten)."""
    lib = _lib.lib()
    b, cin, h, w = x.shape
    cout = dy.shape[1]
    ho, wo = dy.shape[2], dy.shape[3]
    n = int(lib.fm_conv_wgrad_workspace_elems(b, ho, wo, cin, cout, ksize))
    if n <= 0:
        raise RuntimeError(f"fmdm_b200.conv_wgrad: unsupported shape B={b} Ho={ho} Wo={wo} k={ksize}")
    ws = _ws(n, x.device)
    _lib.check(
        lib.fm_conv_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), ws.data_ptr(), b, h, w, cin, cout, ksize,
                               stride, dw.shape[1], c_begin, _stream()),
        "conv_wgrad",
    )


def colsum(dy: torch.Tensor, want_total: bool, total: Optional[torch.Tensor] = None):
    """(per-sample [B][C] fp32 sums over pixels, total [C] or None) of a bf16 NHWC gradient; `total` may be given."""
    lib = _lib.lib()
    b, c, h, w = dy.shape
    ws = _ws(lib.fm_colsum_workspace_elems(b, h * w, c), dy.device)
    out = torch.empty((b, c), dtype=torch.float32, device=dy.device)
    if total is None and want_total:
        total = torch.empty((c,), dtype=torch.float32, device=dy.device)
    _lib.check(lib.fm_colsum_bf16(dy.data_ptr(), ws.data_ptr(), out.data_ptr(), _ptr(total), b, h * w, c,
                                  _tickets(dy.device).data_ptr(), _stream()), "colsum")
    return out, total


def zero_insert2x(x: torch.Tensor) -> torch.Tensor:
    b, c, h, w = x.shape
    out = ops.empty_nhwc(b, c, 2 * h, 2 * w, x.device)
    _lib.check(_lib.lib().fm_zero_insert2x_bf16(x.data_ptr(), out.data_ptr(), b, h, w, c, _stream()), "zero_insert2x")
    return out


def sumpool2x2(x: torch.Tensor) -> torch.Tensor:
    b, c, h2, w2 = x.shape
    out = ops.empty_nhwc(b, c, h2 // 2, w2 // 2, x.device)
    _lib.check(_lib.lib().fm_sumpool2x2_bf16(x.data_ptr(), out.data_ptr(), b, h2 // 2, w2 // 2, c, _stream()),
               "sumpool2x2")
    return out


# the pack plan of the model whose training forward is running (training.packplan); None = pack per call
ACTIVE_PLAN = None


def _dgrad_weight(w: torch.Tensor, c_begin: int, c_count: int, plan=None) -> ops.PackedConvWeight:
    """Packed weights of the data-gradient conv: in/out channels swapped, taps mirrored (one launch)."""
    if plan is not None:
        pw = plan.dgrad_matrix(w, c_begin, c_count, lambda: _dgrad_weight(w, c_begin, c_count))
        if pw is not None:
            return pw
    w32 = w.detach()
    if w32.dtype != torch.float32 or not w32.is_contiguous():
        w32 = w32.float().contiguous()
    cout, cin_total = w32.shape[0], w32.shape[1]
    ks = 1 if w32.dim() == 2 else int(w32.shape[-1])
    mat = torch.empty((c_count, ks * ks * cout), dtype=BF16, device=w32.device)
    _lib.check(_lib.lib().fm_weight_prepack_dgrad_bf16(mat.data_ptr(), w32.data_ptr(), cout, cin_total, c_begin,
                                                       c_count, ks, _stream()), "weight_prepack_dgrad")
    return ops.PackedConvWeight(mat, [cout], [ks], c_count)


# --------------------------------------------------------------------------------------------------------------
# conv
# --------------------------------------------------------------------------------------------------------------
class _ConvFn(Function):
    """out = conv(virtual concat of srcs) + bias + addvec[b] + residual   (ops.conv2d), all segments 'same' padding."""

    @staticmethod
    def forward(ctx, meta, *args):
        stride, segs, nsrc, nw = meta[:4]  # segs: per source (weight index, c_begin, c_count)
        srcs = [_nhwc(a) for a in args[:nsrc]]
        weights = list(args[nsrc:nsrc + nw])
        bias, addvec, residual = args[nsrc + nw:nsrc + nw + 3]
        plan = ACTIVE_PLAN
        pw = plan.forward_matrix(segs, weights) if plan is not None else None
        if pw is None:
            pw = ops.pack_conv_weight([(weights[wi], cb, cc) for wi, cb, cc in segs])
        ctx.plan = plan
        if residual is not None:
            residual = _nhwc(residual)
        # want_stats: the epilogue also emits the GroupNorm partial statistics of the output (`out._fm_stats`), so a
        # GroupNorm that consumes this tensor skips its statistics pass (as on the inference path)
        out = ops.conv2d(srcs, pw, stride=stride, bias=None if bias is None else bias.detach().float().contiguous(),
                         addvec=None if addvec is None else _rows_f32(addvec), residual=residual, want_stats=True)
        ctx.meta = meta
        ctx.flags = (bias is not None, addvec is not None, residual is not None)
        ctx.params = weights  # the tensors as passed (Parameters / views of them): the pack plan keys on their storage
        ctx.bias_param = bias
        ctx.save_for_backward(*srcs, *weights)
        return out

    @staticmethod
    def backward(ctx, dy):
        stride, segs, nsrc, nw, src_slots, res_slot = ctx.meta
        has_bias, has_addvec, has_res = ctx.flags
        saved = ctx.saved_tensors
        srcs, weights = saved[:nsrc], saved[nsrc:]
        dy = _nhwc(dy)
        need = ctx.needs_input_grad  # index 0 is `meta`
        grads = [None] * (nsrc + nw + 3)
        dyz = None
        for i, (wi, cb, cc) in enumerate(segs):
            if not need[1 + i]:
                continue
            pw = _dgrad_weight(ctx.params[wi], cb, cc, ctx.plan)
            role, slot = src_slots[i]
            # gradients other consumers of this source left behind ride in as the dgrad conv's `residual`
            extra = None
            if role == "acc":
                got = slot.take(1)
                extra = got[0] if got else None
            elif role == "park":
                extra = slot.pop_one()
            if stride == 1:
                dx = ops.conv2d([dy], pw, residual=extra)
            else:
                if dyz is None:
                    dyz = zero_insert2x(dy)
                dx = ops.conv2d([dyz], pw, residual=extra)
            if role == "park":
                slot.park(dx)
            else:
                grads[i] = dx
        dws, direct, jobs = {}, set(), []
        for i, (wi, cb, cc) in enumerate(segs):
            if not (need[1 + nsrc + wi] or _is_fused(ctx.params[wi])):
                continue
            w = weights[wi]
            if wi not in dws:
                target = _grad_target(ctx.params[wi], w.shape)
                if target is not None:      # the flat gradient view (pre-zeroed): written in place, nothing returned
                    dws[wi] = target
                    direct.add(wi)
                else:
                    covered = sum(c for j, _, c in segs if j == wi)
                    dws[wi] = (torch.empty if covered == w.shape[1] else torch.zeros)(
                        w.shape, dtype=torch.float32, device=w.device)
            jobs.append((i, wi, cb, 1 if w.dim() == 2 else int(w.shape[-1])))
        need_bias = has_bias and (need[1 + nsrc + nw] or _is_fused(ctx.bias_param))
        need_addvec = has_addvec and need[1 + nsrc + nw + 1]
        btarget = _grad_target(ctx.bias_param, (dy.shape[1],)) if need_bias else None
        tagged = getattr(dy, "_fm_colsum", None)
        part = tagged[0] if tagged is not None and tagged[1] == dy._version else None
        if part is not None and tuple(part.shape[::2]) != (dy.shape[0], dy.shape[1]):
            part = None

        def bias_grads():
            """(per-sample column sums, their total): from the GroupNorm backward's partials if dy carries them."""
            if part is not None:
                # dy is the dx of a GroupNorm backward that already summed its columns per row block
                per_sample = torch.empty((dy.shape[0], dy.shape[1]), dtype=torch.float32, device=dy.device)
                total = btarget if btarget is not None else (
                    torch.empty((dy.shape[1],), dtype=torch.float32, device=dy.device) if has_bias else None)
                _lib.check(_lib.lib().fm_colsum_finish_f32(part.data_ptr(), per_sample.data_ptr(), _ptr(total),
                                                           dy.shape[0], part.shape[1], dy.shape[1], part.stride(1),
                                                           _tickets(dy.device).data_ptr(), _stream()),
                           "colsum_finish")
                return per_sample, total
            return colsum(dy, has_bias, total=btarget)

        # Nothing downstream waits for a parameter gradient, so on small tensors - where the data-gradient chain is a
        # string of latency-bound launches that leave most SMs idle - the weight-gradient GEMMs, their split-K folds
        # and the bias column sums run on a side stream beside that chain (a parallel branch of the step graph).
        # (not when dy itself is handed back to autograd as the residual's gradient: the engine may accumulate into it)
        on_side = (SIDE_STREAMS and dy.is_cuda and jobs and len(direct) == len(dws)
                   and not (has_res and need[1 + nsrc + nw + 2] and res_slot is None)
                   and dy.numel() <= SIDE_STREAM_MAX_ELEMS and all(srcs[i].numel() <= SIDE_STREAM_MAX_ELEMS
                                                                   for i, _, _, _ in jobs))
        bias_on_side = on_side and need_bias and btarget is not None and not need_addvec
        with _side_branch([dy] + [srcs[i] for i, _, _, _ in jobs] + ([part] if part is not None and bias_on_side else
                                                                     [])) if on_side else contextlib.nullcontext():
            for i, wi, cb, ks in jobs:
                conv_wgrad(dy, srcs[i], dws[wi], ksize=ks, stride=stride, c_begin=cb)
            if bias_on_side:
                bias_grads()
        for wi, dw in dws.items():
            if wi in direct:
                _grad_written(ctx.params[wi])
            else:
                grads[nsrc + wi] = dw
        if bias_on_side:
            _grad_written(ctx.bias_param)
        elif need_bias or need_addvec:
            per_sample, total = bias_grads()
            if btarget is not None:
                _grad_written(ctx.bias_param)
            elif has_bias:
                grads[nsrc + nw] = total
            if has_addvec:
                grads[nsrc + nw + 1] = per_sample
        if has_res and need[1 + nsrc + nw + 2]:
            if res_slot is not None:
                res_slot.park(dy)
            else:
                grads[nsrc + nw + 2] = dy
        return (None, *grads)


def conv(srcs: Sequence[torch.Tensor], weights: Sequence[tuple], *, bias=None, stride: int = 1, addvec=None,
         residual=None) -> torch.Tensor:
    """srcs[i] is convolved with weights[i] = (param, c_begin, c_count): the channel slice of an OIHW (or [O][I])
    fp32 parameter; several sources may slice the same parameter (virtual concat) or different ones (conv2 + skip)."""
    uniq, segs = [], []
    for w, cb, cc in weights:
        for j, u in enumerate(uniq):
            if u is w:
                segs.append((j, int(cb), int(cc)))
                break
        else:
            uniq.append(w)
            segs.append((len(uniq) - 1, int(cb), int(cc)))
    # gradient slots (see _Slot): a source this conv consumes first accumulates, a source / residual that already has
    # an accumulating consumer parks
    src_slots = []
    for t in srcs:
        joined = _slot_join(t)
        if joined is not None:
            src_slots.append(("park", joined))
        else:
            opened = _slot_open(t)
            src_slots.append(("acc", opened) if opened is not None else (None, None))
    res_slot = _slot_join(residual) if residual is not None else None
    meta = (int(stride), tuple(segs), len(srcs), len(uniq), tuple(src_slots), res_slot)
    return _ConvFn.apply(meta, *srcs, *uniq, bias, addvec, residual)


# --------------------------------------------------------------------------------------------------------------
# GroupNorm (+ scale-shift) (+ SiLU)
# --------------------------------------------------------------------------------------------------------------
class _GroupNormFn(Function):
    """GroupNorm over the virtual channel concat of one or two sources; the (materialised) result is one tensor."""

    @staticmethod
    def forward(ctx, x0, x1, gamma, beta, scale_shift, groups, eps, silu, slots):
        lib = _lib.lib()
        ctx.slots = slots
        x0 = _nhwc(x0)
        x1 = None if x1 is None else _nhwc(x1)
        b, c0, h, w = x0.shape
        c1 = 0 if x1 is None else x1.shape[1]
        c = c0 + c1
        g32 = gamma.detach().float().contiguous()
        b32 = beta.detach().float().contiguous()
        ss = None if scale_shift is None else _rows_f32(scale_shift)
        st = _stream()
        stats = torch.empty((b, groups, 2), dtype=torch.float32, device=x0.device)
        fused = [getattr(s, "_fm_stats", None) for s in ((x0,) if x1 is None else (x0, x1))]
        if all(f is not None for f in fused) and (c // groups) % 4 == 0:
            # the convs that wrote the sources left channel-quad (sum, sumsq) partials: fold them, no read pass
            p1 = fused[1] if len(fused) == 2 else (None, 0)
            _lib.check(lib.fm_groupnorm_finalize_partials(fused[0][0].data_ptr(), fused[0][1], c0, _ptr(p1[0]), p1[1],
                                                          c1, b, h * w, groups, float(eps), stats.data_ptr(), st),
                       "groupnorm_finalize_partials")
        else:
            n = int(lib.fm_groupnorm_workspace_elems(b, h * w, c, groups))
            if n <= 0:
                raise RuntimeError(f"fmdm_b200.group_norm: unsupported shape B={b} HW={h * w} C={c} groups={groups}")
            ws = _ws(n, x0.device)
            _lib.check(lib.fm_groupnorm_stats_bf16(x0.data_ptr(), c0, _ptr(x1), c1, b, h * w, groups, float(eps),
                                                   ws.data_ptr(), stats.data_ptr(), st), "groupnorm_stats")
        out = ops.empty_nhwc(b, c, h, w, x0.device)
        _lib.check(lib.fm_groupnorm_apply_bf16(x0.data_ptr(), c0, _ptr(x1), c1, b, h * w, groups, stats.data_ptr(),
                                               g32.data_ptr(), b32.data_ptr(), _ptr(ss),
                                               0 if ss is None else ss.stride(0), int(silu), out.data_ptr(), st),
                   "groupnorm_apply")
        ctx.cfg = (groups, bool(silu), ss is not None, x1 is not None)
        ctx.affine = (gamma, beta)
        ctx.save_for_backward(x0, stats, g32, b32, *([x1] if x1 is not None else []), *([ss] if ss is not None else []))
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.lib()
        groups, silu, has_ss, has_x1 = ctx.cfg
        x0, stats, g32, b32, *rest = ctx.saved_tensors
        x1 = rest.pop(0) if has_x1 else None
        ss = rest.pop(0) if has_ss else None
        dout = _nhwc(dout)
        b, c0, h, w = x0.shape
        c1 = 0 if x1 is None else x1.shape[1]
        c = c0 + c1
        ws = _ws(lib.fm_groupnorm_bwd_workspace_elems(b, h * w, c), x0.device)
        dx0 = ops.empty_nhwc(b, c0, h, w, x0.device)
        dx1 = ops.empty_nhwc(b, c1, h, w, x0.device) if has_x1 else None
        gamma, beta = ctx.affine
        tg, tb = _grad_target(gamma, (c,)), _grad_target(beta, (c,))
        direct = tg is not None and tb is not None   # dgamma / dbeta land in the flat gradient, no AccumulateGrad
        dgb = None if direct else torch.empty((2, c), dtype=torch.float32, device=x0.device)
        dss = torch.empty((b, 2 * c), dtype=torch.float32, device=x0.device) if has_ss else None
        # gradients the other consumers of x0 / x1 parked (`_Slot`): added by the apply pass while it writes dx
        (role0, slot0), (role1, slot1) = ctx.slots
        add0 = slot0.take(2) if role0 == "acc" else ([slot0.pop_one()] if role0 == "park" else [])
        add1 = slot1.take(1) if role1 == "acc" else ([slot1.pop_one()] if role1 == "park" else [])
        add0 = [_nhwc(t) for t in add0 if t is not None]
        add1 = [_nhwc(t) for t in add1 if t is not None]
        # also emit the column sums of dx (first stage); if x came straight out of a conv, that conv's backward turns
        # them into its bias / embedding-add gradients without another pass over dx
        nblk = int(lib.fm_groupnorm_bwd_blocks(b, h * w))
        colpart = torch.empty((b, nblk, c), dtype=torch.float32, device=x0.device) if nblk > 0 else None
        _lib.check(
            lib.fm_groupnorm_bwd_bf16(x0.data_ptr(), c0, _ptr(x1), c1, dout.data_ptr(), stats.data_ptr(),
                                      g32.data_ptr(), b32.data_ptr(), _ptr(ss), 0 if ss is None else ss.stride(0),
                                      int(silu), b, h * w, groups, ws.data_ptr(), dx0.data_ptr(), _ptr(dx1),
                                      tg.data_ptr() if direct else dgb.data_ptr(), _ptr(dss), _ptr(colpart),
                                      tb.data_ptr() if direct else None,
                                      add0[0].data_ptr() if add0 else None, add0[1].data_ptr() if len(add0) > 1 else None,
                                      add1[0].data_ptr() if add1 else None, _tickets(x0.device).data_ptr(), _stream()),
            "groupnorm_bwd",
        )
        if colpart is not None:
            # valid only for this exact tensor state: autograd may accumulate another branch's gradient into dx in
            # place, which bumps `_version` and invalidates the sums
            dx0._fm_colsum = (colpart[:, :, :c0], dx0._version)
            if has_x1:
                dx1._fm_colsum = (colpart[:, :, c0:], dx1._version)
        if role0 == "park":
            slot0.park(dx0)
            dx0 = None
        if has_x1 and role1 == "park":
            slot1.park(dx1)
            dx1 = None
        if direct:
            _grad_written(gamma)
            _grad_written(beta)
            return dx0, dx1, None, None, dss, None, None, None, None
        return dx0, dx1, dgb[0], dgb[1], dss, None, None, None, None


def group_norm(x, gamma, beta, *, groups: int, eps: float, silu: bool, scale_shift=None) -> torch.Tensor:
    """`x`: a tensor, or a pair of tensors read as their channel concat (never materialised)."""
    x0, x1 = (x[0], x[1]) if isinstance(x, (tuple, list)) else (x, None)
    slots = []
    for t in (x0, x1):
        if t is None:
            slots.append((None, None))
            continue
        joined = _slot_join(t)
        if joined is not None:
            slots.append(("park", joined))
        else:
            opened = _slot_open(t)
            slots.append(("acc", opened) if opened is not None else (None, None))
    return _GroupNormFn.apply(x0, x1, gamma, beta, scale_shift, int(groups), float(eps), bool(silu), tuple(slots))


# --------------------------------------------------------------------------------------------------------------
# self-attention over a fused qkv projection (DiffusersAttentionND layout: NHWC [b][T][3C], head h at h*dh)
# --------------------------------------------------------------------------------------------------------------
class _AttentionFn(Function):
    @staticmethod
    def forward(ctx, qkv, heads):
        qkv = _nhwc(qkv)
        b, c3, hh, ww = qkv.shape
        c = c3 // 3
        t, hd = hh * ww, c // heads
        att = ops.empty_nhwc(b, c, hh, ww, qkv.device)
        flat = qkv.permute(0, 2, 3, 1).reshape(-1)
        ops.attention(flat, flat[c:], flat[2 * c:], att.permute(0, 2, 3, 1).reshape(-1), batch=b, heads=heads, tq=t,
                      tk=t, head_dim=hd, q_strides=(t * 3 * c, hd, 3 * c), kv_strides=(t * 3 * c, hd, 3 * c),
                      o_strides=(t * c, hd, c))
        ctx.heads = heads
        ctx.save_for_backward(qkv, att)
        return att

    @staticmethod
    def backward(ctx, datt):
        qkv, att = ctx.saved_tensors
        heads = ctx.heads
        datt = _nhwc(datt)
        b, c3, hh, ww = qkv.shape
        c = c3 // 3
        t, hd = hh * ww, c // heads
        dqkv = ops.empty_nhwc(b, c3, hh, ww, qkv.device)
        esz = 2
        q, dq = qkv.data_ptr(), dqkv.data_ptr()
        _lib.check(
            _lib.lib().fm_attention_bwd_bf16(q, q + c * esz, q + 2 * c * esz, att.data_ptr(), datt.data_ptr(), dq,
                                             dq + c * esz, dq + 2 * c * esz, b, heads, t, hd, t * 3 * c, hd, 3 * c,
                                             t * c, hd, c, 1.0 / math.sqrt(hd), _stream()),
            "attention_bwd",
        )
        return dqkv, None


def attention_qkv(qkv: torch.Tensor, heads: int) -> torch.Tensor:
    return _AttentionFn.apply(qkv, int(heads))


class _CrossAttentionFn(Function):
    """Cross-attention of `DiffusersAttentionND(context_dim)` (`attention.py:232-274`): q NHWC [b][T][C] from the image
    tokens, kv [b][Tc][2C] (K | V, head h at h*hd) from the context path; head_dim 8."""

    @staticmethod
    def forward(ctx, q, kv, heads):
        q = _nhwc(q)
        b, c, hh, ww = q.shape
        t, hd, tc = hh * ww, c // heads, kv.shape[1]
        assert kv.dtype == BF16 and kv.is_contiguous() and kv.shape[2] == 2 * c
        att = ops.empty_nhwc(b, c, hh, ww, q.device)
        kvf = kv.view(-1)
        ops.attention(q.permute(0, 2, 3, 1).reshape(-1), kvf, kvf[c:], att.permute(0, 2, 3, 1).reshape(-1), batch=b,
                      heads=heads, tq=t, tk=tc, head_dim=hd, q_strides=(t * c, hd, c),
                      kv_strides=(tc * 2 * c, hd, 2 * c), o_strides=(t * c, hd, c))
        ctx.heads = heads
        ctx.save_for_backward(q, kv, att)
        return att

    @staticmethod
    def backward(ctx, datt):
        q, kv, att = ctx.saved_tensors
        heads = ctx.heads
        datt = _nhwc(datt)
        b, c, hh, ww = q.shape
        t, hd, tc = hh * ww, c // heads, kv.shape[1]
        dq = ops.empty_nhwc(b, c, hh, ww, q.device)
        dkv = torch.empty_like(kv)
        k0, d0 = kv.data_ptr(), dkv.data_ptr()
        _lib.check(
            _lib.lib().fm_attention_bwd_cross_bf16(q.data_ptr(), k0, k0 + 2 * c, att.data_ptr(), datt.data_ptr(),
                                                   dq.data_ptr(), d0, d0 + 2 * c, b, heads, t, tc, hd, t * c, hd, c,
                                                   tc * 2 * c, hd, 2 * c, t * c, hd, c, 1.0 / math.sqrt(hd), _stream()),
            "attention_bwd_cross",
        )
        return dq, dkv, None


def cross_attention(q: torch.Tensor, kv: torch.Tensor, heads: int) -> torch.Tensor:
    return _CrossAttentionFn.apply(q, kv, int(heads))


class _ContextKVFn(Function):
    """The cross-attention context path (`fm_context_kv_bf16`, token-major): GroupNorm over the context tokens + the
    key | value projection; gradients to the projection and to the context GroupNorm's affine (the context is data)."""

    @staticmethod
    def forward(ctx, tokens, gamma, beta, weight, bias, groups, eps):
        tokens = tokens.detach().float().contiguous()
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        w32 = weight.detach().float().contiguous()
        bias32 = None if bias is None else bias.detach().float().contiguous()
        kv, stats = ops.context_kv(tokens, g32, b32, w32, bias32, groups=groups, eps=eps, channel_major=False,
                                   return_stats=True)
        ctx.groups = int(groups)
        ctx.params = (gamma, beta, weight, bias)
        ctx.save_for_backward(tokens, stats, g32, b32, w32)
        return kv

    @staticmethod
    def backward(ctx, dkv):
        lib = _lib.lib()
        tokens, stats, g32, b32, w32 = ctx.saved_tensors
        gamma, beta, weight, bias = ctx.params
        dkv = dkv.to(BF16).contiguous()
        b, cc, tc = tokens.shape
        o = w32.shape[0]
        ws = _ws(lib.fm_context_kv_bwd_workspace_elems(b, cc, tc, o), tokens.device)
        tw, tb = _grad_target(weight, w32.shape), (_grad_target(bias, (o,)) if bias is not None else None)
        tg, tbe = _grad_target(gamma, (cc,)), _grad_target(beta, (cc,))
        dw = tw if tw is not None else torch.empty_like(w32)
        db = tb if tb is not None else (torch.empty((o,), dtype=torch.float32, device=tokens.device)
                                        if bias is not None else None)
        dg = tg if tg is not None else torch.empty((cc,), dtype=torch.float32, device=tokens.device)
        dbe = tbe if tbe is not None else torch.empty((cc,), dtype=torch.float32, device=tokens.device)
        _lib.check(lib.fm_context_kv_bwd_f32(tokens.data_ptr(), stats.data_ptr(), g32.data_ptr(), b32.data_ptr(),
                                             w32.data_ptr(), dkv.data_ptr(), ws.data_ptr(), dw.data_ptr(), _ptr(db),
                                             dg.data_ptr(), dbe.data_ptr(), b, cc, tc, o, ctx.groups, _stream()),
                   "context_kv_bwd")
        for param, target in ((weight, tw), (bias, tb), (gamma, tg), (beta, tbe)):
            if target is not None:
                _grad_written(param)
        return (None, None if tg is not None else dg, None if tbe is not None else dbe, None if tw is not None else dw,
                None if (tb is not None or bias is None) else db, None, None)


def context_kv(tokens, gamma, beta, weight, bias, *, groups: int, eps: float) -> torch.Tensor:
    return _ContextKVFn.apply(tokens, gamma, beta, weight, bias, int(groups), float(eps))


LINEAR_ATTENTION_EPS = 1e-6   # LinearQKVAttention's default (`attention.py:58`)


class _AttentionRawFn(Function):
    """SpatialSelfAttention's raw-reshape head split (`attention.py:111-115`): the channel-major (b, 3*inner, T) qkv
    buffer re-read as (b, heads, T, 3*dh); returns (b, heads, T, dh) contiguous.  `linear`: LinearQKVAttention."""

    @staticmethod
    def forward(ctx, qkv_cm, heads, dh, linear):
        assert qkv_cm.dtype == BF16 and qkv_cm.is_contiguous() and qkv_cm.dim() == 3
        b, c3, t = qkv_cm.shape
        inner = heads * dh
        assert c3 == 3 * inner
        att = torch.empty((b, heads, t, dh), dtype=BF16, device=qkv_cm.device)
        flat = qkv_cm.view(-1)
        kw = dict(batch=b, heads=heads, tq=t, tk=t, head_dim=dh, q_strides=(3 * inner * t, t * 3 * dh, 3 * dh),
                  kv_strides=(3 * inner * t, t * 3 * dh, 3 * dh), o_strides=(heads * t * dh, t * dh, dh))
        if linear:
            ops.linear_attention(flat, flat[dh:], flat[2 * dh:], att, eps=LINEAR_ATTENTION_EPS, **kw)
        else:
            ops.attention(flat, flat[dh:], flat[2 * dh:], att, **kw)
        ctx.cfg = (heads, dh, bool(linear))
        ctx.save_for_backward(qkv_cm, att)
        return att

    @staticmethod
    def backward(ctx, datt):
        qkv_cm, att = ctx.saved_tensors
        heads, dh, linear = ctx.cfg
        b, c3, t = qkv_cm.shape
        inner = heads * dh
        datt = datt.to(BF16).contiguous()
        dqkv = torch.empty_like(qkv_cm)
        q, dq = qkv_cm.data_ptr(), dqkv.data_ptr()
        qs, os_ = (3 * inner * t, t * 3 * dh, 3 * dh), (heads * t * dh, t * dh, dh)
        if linear:
            _lib.check(
                _lib.lib().fm_linear_attention_bwd_bf16(q, q + 2 * dh, q + 4 * dh, datt.data_ptr(), dq, dq + 2 * dh,
                                                        dq + 4 * dh, b, heads, t, t, dh, *qs, *qs, *os_,
                                                        LINEAR_ATTENTION_EPS, _stream()), "linear_attention_bwd")
        else:
            _lib.check(
                _lib.lib().fm_attention_bwd_bf16(q, q + 2 * dh, q + 4 * dh, att.data_ptr(), datt.data_ptr(), dq,
                                                 dq + 2 * dh, dq + 4 * dh, b, heads, t, dh, *qs, *os_,
                                                 1.0 / math.sqrt(dh), _stream()), "attention_bwd")
        return dqkv, None, None, None


def attention_raw(qkv_cm: torch.Tensor, heads: int, dim_head: int, linear: bool = False) -> torch.Tensor:
    return _AttentionRawFn.apply(qkv_cm, int(heads), int(dim_head), bool(linear))


class _CrossAttentionRawFn(Function):
    """SpatialCrossAttention's raw-reshape head split (`attention.py:176-186`): q_cm (b, inner, T) re-read as
    (b, heads, T, dh), kv_cm (b, 2*inner, Tc) as (b, heads, Tc, 2*dh) = K | V; returns (b, heads, T, dh)."""

    @staticmethod
    def forward(ctx, q_cm, kv_cm, heads, dh, linear):
        assert q_cm.dtype == BF16 and q_cm.is_contiguous() and kv_cm.dtype == BF16 and kv_cm.is_contiguous()
        b, inner, t = q_cm.shape
        tc = kv_cm.shape[-1]
        assert inner == heads * dh and kv_cm.shape[1] == 2 * inner
        att = torch.empty((b, heads, t, dh), dtype=BF16, device=q_cm.device)
        kvf = kv_cm.view(-1)
        kw = dict(batch=b, heads=heads, tq=t, tk=tc, head_dim=dh, q_strides=(inner * t, t * dh, dh),
                  kv_strides=(2 * inner * tc, tc * 2 * dh, 2 * dh), o_strides=(heads * t * dh, t * dh, dh))
        if linear:
            ops.linear_attention(q_cm.view(-1), kvf, kvf[dh:], att, eps=LINEAR_ATTENTION_EPS, **kw)
        else:
            ops.attention(q_cm.view(-1), kvf, kvf[dh:], att, **kw)
        ctx.cfg = (heads, dh, bool(linear))
        ctx.save_for_backward(q_cm, kv_cm, att)
        return att

    @staticmethod
    def backward(ctx, datt):
        q_cm, kv_cm, att = ctx.saved_tensors
        heads, dh, linear = ctx.cfg
        b, inner, t = q_cm.shape
        tc = kv_cm.shape[-1]
        datt = datt.to(BF16).contiguous()
        dq, dkv = torch.empty_like(q_cm), torch.empty_like(kv_cm)
        k0, d0 = kv_cm.data_ptr(), dkv.data_ptr()
        qs, ks, os_ = (inner * t, t * dh, dh), (2 * inner * tc, tc * 2 * dh, 2 * dh), (heads * t * dh, t * dh, dh)
        if linear:
            _lib.check(
                _lib.lib().fm_linear_attention_bwd_bf16(q_cm.data_ptr(), k0, k0 + 2 * dh, datt.data_ptr(), dq.data_ptr(),
                                                        d0, d0 + 2 * dh, b, heads, t, tc, dh, *qs, *ks, *os_,
                                                        LINEAR_ATTENTION_EPS, _stream()), "linear_attention_bwd")
        else:
            _lib.check(
                _lib.lib().fm_attention_bwd_cross_bf16(q_cm.data_ptr(), k0, k0 + 2 * dh, att.data_ptr(), datt.data_ptr(),
                                                       dq.data_ptr(), d0, d0 + 2 * dh, b, heads, t, tc, dh, *qs, *ks,
                                                       *os_, 1.0 / math.sqrt(dh), _stream()), "attention_bwd_cross")
        return dq, dkv, None, None, None


def cross_attention_raw(q_cm, kv_cm, heads: int, dim_head: int, linear: bool = False) -> torch.Tensor:
    return _CrossAttentionRawFn.apply(q_cm, kv_cm, int(heads), int(dim_head), bool(linear))


class _TransposeFn(Function):
    """[B][R][C] -> [B][C][R] (bf16, contiguous); the backward is the same kernel."""

    @staticmethod
    def forward(ctx, x):
        return ops.transpose_bf16(x.to(BF16).contiguous())

    @staticmethod
    def backward(ctx, dy):
        return ops.transpose_bf16(dy.to(BF16).contiguous())


def transpose(x: torch.Tensor) -> torch.Tensor:
    return _TransposeFn.apply(x)


# --------------------------------------------------------------------------------------------------------------
# tiny fp32 Linear (time MLP, per-block embedding projections): y = f(x) W^T + b, f = SiLU if silu_in
# --------------------------------------------------------------------------------------------------------------
class _LinearFn(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, silu_in):
        x32 = x.detach().float().contiguous()
        w32 = weight.detach().float().contiguous()
        y = ops.linear_f32(x32, w32, None if bias is None else bias.detach().float().contiguous(), silu_in=silu_in)
        ctx.silu_in = bool(silu_in)
        ctx.has_bias = bias is not None
        ctx.wb = (weight, bias)
        ctx.save_for_backward(x32, w32)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.float().contiguous()
        dx = dw = db = None
        if x.shape[0] <= 32:
            # small-batch kernels: dW (+db) in one launch, dX as a split-K pair
            lib = _lib.lib()
            b, i = x.shape
            o = w.shape[0]
            want_dx = ctx.needs_input_grad[0]
            ws = _ws(lib.fm_linear_bwd_workspace_elems(b, i, o), x.device) if want_dx else None
            dx = torch.empty_like(x) if want_dx else None
            wp, bp = ctx.wb
            tw = _grad_target(wp, w.shape)
            tb = _grad_target(bp, (o,)) if ctx.has_bias else None
            dw = tw if tw is not None else torch.empty_like(w)
            db = tb if tb is not None else (torch.empty((o,), dtype=torch.float32, device=x.device)
                                            if ctx.has_bias else None)
            _lib.check(lib.fm_linear_bwd_f32(x.data_ptr(), w.data_ptr(), dy.data_ptr(), _ptr(ws), _ptr(dx),
                                             dw.data_ptr(), _ptr(db), b, i, o, int(ctx.silu_in), _stream()),
                       "linear_bwd")
            if tw is not None:
                _grad_written(wp)
            if tb is not None:
                _grad_written(bp)
            return dx, None if tw is not None else dw, None if tb is not None else db, None
        dyt = dy.t().contiguous()                                              # [O][B]
        if ctx.needs_input_grad[0]:
            dx = ops.linear_f32(dy, w.t().contiguous())                        # [B][I] = dy W
            if ctx.silu_in:
                out = torch.empty_like(dx)
                _lib.check(_lib.lib().fm_silu_bwd_f32(x.data_ptr(), dx.data_ptr(), out.data_ptr(), dx.numel(),
                                                      _stream()), "silu_bwd")
                dx = out
        if ctx.needs_input_grad[1]:
            # dW^T[i][o] = sum_b f(x[b][i]) dy[b][o]: the same kernel with the batch as the reduction axis
            dw = ops.linear_f32(x.t().contiguous(), dyt, silu_in=ctx.silu_in).t().contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            ones = torch.ones((1, dy.shape[0]), dtype=torch.float32, device=dy.device)
            db = ops.linear_f32(ones, dyt).reshape(-1)
        return dx, dw, db, None


def linear(x, weight, bias=None, *, silu_in: bool = False) -> torch.Tensor:
    return _LinearFn.apply(x, weight, bias, bool(silu_in))


class _SplitColsFn(Function):
    """Column slices of a [B][sum(sizes)] matrix as views; the backward gathers the slice gradients with one concat
    (autograd's own slice backward would pad and add one full-width matrix per slice)."""

    @staticmethod
    def forward(ctx, x, sizes):
        ctx.sizes = sizes
        ctx.rows = x.shape[0]
        outs, off = [], 0
        for n in sizes:
            outs.append(x[:, off:off + n])
            off += n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        ref = next(g for g in grads if g is not None)
        parts = [g if g is not None else torch.zeros((ctx.rows, n), dtype=ref.dtype, device=ref.device)
                 for g, n in zip(grads, ctx.sizes)]
        return torch.cat(parts, 1), None


def split_cols(x: torch.Tensor, sizes: Sequence[int]):
    return _SplitColsFn.apply(x, tuple(int(n) for n in sizes))


# --------------------------------------------------------------------------------------------------------------
# stem / head convs, nearest upsample
# --------------------------------------------------------------------------------------------------------------
class _StemFn(Function):
    @staticmethod
    def forward(ctx, x0, x1, weight, bias, in_scale, in_shift):
        x0 = x0.detach().float().contiguous()
        x1 = None if x1 is None else x1.detach().float().contiguous()
        w32 = weight.detach().float().contiguous()
        out = ops.conv_stem(x0, x1, w32, None if bias is None else bias.detach().float().contiguous(),
                            in_scale=in_scale, in_shift=in_shift, want_stats=True,
                            tensor_cores=SMALL_CONVS_ON_TENSOR_CORES)
        ctx.cfg = (float(in_scale), float(in_shift), x1 is not None, bias is not None, tuple(weight.shape))
        ctx.wb = (weight, bias)
        ctx.save_for_backward(x0, *([x1] if x1 is not None else []))
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.lib()
        in_scale, in_shift, has_x1, has_bias, wshape = ctx.cfg
        x0, *rest = ctx.saved_tensors
        x1 = rest[0] if has_x1 else None
        dy = _nhwc(dy)
        b, c0, h, w = x0.shape
        c1 = x1.shape[1] if x1 is not None else 0
        cout = wshape[0]
        wp, bp = ctx.wb
        tw = _grad_target(wp, wshape)
        tb = _grad_target(bp, (cout,)) if has_bias else None
        dw = tw if tw is not None else torch.empty(wshape, dtype=torch.float32, device=dy.device)
        if SMALL_CONVS_ON_TENSOR_CORES and 9 * (c0 + c1) <= 72:
            # dW[co][ci*9 + tap] = sum_p dY[p][co] * im2col(x)[p][ci*9 + tap]: the 1x1 weight-gradient GEMM over the
            # stem's im2col matrix (tcgen05), instead of a CUDA-core kernel that re-reads x through shared memory
            cols = ops.stem_im2col(x0, x1, in_scale=in_scale, in_shift=in_shift)
            dw2 = torch.empty((cout, cols.shape[1]), dtype=torch.float32, device=dy.device)
            conv_wgrad(dy, cols, dw2, ksize=1, stride=1, c_begin=0)
            dw.view(cout, -1).copy_(dw2[:, :9 * (c0 + c1)])
        else:
            ws = _ws(lib.fm_conv_stem_wgrad_workspace_elems(c0 + c1, cout), dy.device)
            _lib.check(lib.fm_conv_stem_wgrad_f32(x0.data_ptr(), c0, _ptr(x1), c1, in_scale, in_shift, dy.data_ptr(),
                                                  ws.data_ptr(), dw.data_ptr(), b, h, w, cout, _stream()), "stem_wgrad")
        db = colsum(dy, True, total=tb)[1] if has_bias else None
        if tw is not None:
            _grad_written(wp)
        if tb is not None:
            _grad_written(bp)
        return None, None, None if tw is not None else dw, None if tb is not None else db, None, None


def conv_stem(x0, x1, weight, bias, *, in_scale: float = 1.0, in_shift: float = 0.0) -> torch.Tensor:
    return _StemFn.apply(x0, x1, weight, bias, float(in_scale), float(in_shift))


def sum_f32(x: torch.Tensor, *, t1=None, t2=None, mode: int = 0, scale: float = 1.0) -> torch.Tensor:
    ws = _ws(1024, x.device, torch.float64)
    out = torch.empty((1,), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().fm_sum_f32(x.data_ptr(), _ptr(t1), _ptr(t2), ws.data_ptr(), out.data_ptr(), x.numel(), mode,
                                     float(scale), _stream()), "sum_f32")
    return out


class _HeadFn(Function):
    """3x3 head conv to ONE output channel: bf16 NHWC -> fp32 NCHW."""

    @staticmethod
    def forward(ctx, a, weight, bias):
        a = _nhwc(a)
        w32 = weight.detach().float().contiguous()
        out = ops.conv_head(a, w32, None if bias is None else bias.detach().float().contiguous())
        ctx.has_bias = bias is not None
        ctx.wb = (weight, bias)
        ctx.save_for_backward(a, w32)
        return out

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.lib()
        a, w32 = ctx.saved_tensors
        dy = dy.float().contiguous()
        b, cin, h, w = a.shape
        wp, bp = ctx.wb
        tw = _grad_target(wp, w32.shape)
        dw = tw if tw is not None else torch.empty_like(w32)
        want_da = ctx.needs_input_grad[0]
        if SMALL_CONVS_ON_TENSOR_CORES:
            # both gradients are GEMMs over the 3x3 neighbourhoods of dY (one im2col, [p][16] bf16, column k = the tap):
            #   dA[p][c]  = sum_k cols[p][k] * w[c][8 - k]            -> the 1x1 implicit GEMM (forward conv kernel)
            #   dW[c][k]  = sum_q a[q][c] * cols[q][8 - k]            -> the 1x1 weight-gradient GEMM, columns flipped
            cols = ops.stem_im2col(dy, None)
            kp = cols.shape[1]
            da = None
            if want_da:
                wd = torch.zeros((cin, kp), dtype=torch.float32, device=a.device)
                wd[:, :9] = w32.reshape(cin, 9).flip(1)
                da = ops.conv2d([cols], ops.pack_conv_weight([(wd, 0, kp)]))
            dw2 = torch.empty((cin, kp), dtype=torch.float32, device=a.device)
            conv_wgrad(a, cols, dw2, ksize=1, stride=1, c_begin=0)
            dw.view(cin, 9).copy_(dw2[:, :9].flip(1))
        else:
            ws = _ws(lib.fm_conv_head_bwd_workspace_elems(cin), a.device)
            da = ops.empty_nhwc(b, cin, h, w, a.device) if want_da else None
            _lib.check(lib.fm_conv_head_bwd_f32(a.data_ptr(), dy.data_ptr(), w32.data_ptr(), ws.data_ptr(), _ptr(da),
                                                dw.data_ptr(), b, h, w, cin, _stream()), "head_bwd")
        db = sum_f32(dy) if ctx.has_bias else None
        if tw is not None:
            _grad_written(wp)
        return da, None if tw is not None else dw, db


def conv_head(a, weight, bias) -> torch.Tensor:
    if weight.shape[0] != 1:
        raise RuntimeError("fmdm_b200.training: the head conv backward supports out_channels == 1")
    return _HeadFn.apply(a, weight, bias)


class _UpsampleFn(Function):
    @staticmethod
    def forward(ctx, x):
        return ops.upsample_nearest2x(_nhwc(x))

    @staticmethod
    def backward(ctx, dy):
        return sumpool2x2(_nhwc(dy))


def upsample_nearest2x(x) -> torch.Tensor:
    return _UpsampleFn.apply(x)


# --------------------------------------------------------------------------------------------------------------
# fused velocity-target MSE (flow_matching_lib.py:163-164): mean((pred - (noise - clean))^2)
# --------------------------------------------------------------------------------------------------------------
class _MseFn(Function):
    @staticmethod
    def forward(ctx, pred, t1, t2):
        pred = pred.float().contiguous()
        t1 = t1.detach().float().contiguous()
        t2 = None if t2 is None else t2.detach().float().contiguous()
        ctx.save_for_backward(pred, t1, *([t2] if t2 is not None else []))
        return sum_f32(pred, t1=t1, t2=t2, mode=1, scale=1.0 / pred.numel()).reshape(())

    @staticmethod
    def backward(ctx, g):
        pred, t1, *rest = ctx.saved_tensors
        t2 = rest[0] if rest else None
        g = g.float().reshape(1).contiguous()
        dpred = torch.empty_like(pred)
        _lib.check(_lib.lib().fm_mse_bwd_f32(pred.data_ptr(), t1.data_ptr(), _ptr(t2), g.data_ptr(), dpred.data_ptr(),
                                             pred.numel(), _stream()), "mse_bwd")
        return dpred, None, None


def mse_loss(pred: torch.Tensor, target: torch.Tensor, minus: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mean((pred - (target - minus))^2); `minus=None` is plain F.mse_loss(pred, target)."""
    return _MseFn.apply(pred, target, minus)
