"""Thin tensor-level wrappers over the C-ABI kernels.

PyTorch is plumbing here: it owns device memory (caching allocator) and the current stream; every function below
launches exactly the hand-written sm_100a kernels through ctypes.  Activations are logical NCHW tensors stored
channels_last in bf16 ("NHWC bf16"); the sampler state and model prediction are fp32 NCHW.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib

BF16 = torch.bfloat16


class pdl:
    """`with ops.pdl(True): ...` - launches inside carry the programmatic-dependent-launch attribute (`fm_set_pdl`);
    a CUDA graph captured inside keeps the programmatic edges for all its replays."""

    def __init__(self, on: bool):
        self.on, self._was = bool(on), 0

    def __enter__(self):
        self._was = int(_lib.lib().fm_set_pdl(int(self.on)))
        return self

    def __exit__(self, *exc):
        _lib.lib().fm_set_pdl(self._was)
        return False


def _stream() -> int:
    """Raw handle of torch's current stream on the CURRENT device.  Every entry point checks (`require_cuda`) that its
    tensors live on that device: the C-ABI launches on the current device, so a tensor of another GPU would otherwise
    be read through a foreign pointer on the wrong stream.  Callers switch devices with `torch.cuda.device(...)`
    (`BaseUNetND.forward`, `sample_with_scheduler` and the trainers do it from their input tensor)."""
    return torch.cuda.current_stream().cuda_stream


# ---- optional per-kernel timing (bench.py's live roofline measurement; CUDA events on the launching stream) ----
_PROFILE = None


class profile:
    """with ops.profile() as rec: ...  -> rec.rows = [(tag, work, milliseconds)] after exit (synchronises)."""

    def __enter__(self):
        global _PROFILE
        self._events = []
        _PROFILE = self._events
        self.rows = []
        return self

    def __exit__(self, *exc):
        global _PROFILE
        _PROFILE = None
        torch.cuda.synchronize()
        self.rows = [(tag, work, e0.elapsed_time(e1)) for tag, work, e0, e1 in self._events]
        return False


def _prof_begin():
    if _PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(tag: str, work: float, e0) -> None:
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    _PROFILE.append((tag, work, e0, e1))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"fmdm_b200.{what}: tensor is on {t.device}; the B200 hot path has no CPU implementation "
            "(use the oracle/ package or the reference for CPU runs)."
        )
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(
            f"fmdm_b200.{what}: tensor is on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
            "kernels launch on the current device's stream - wrap the call in `torch.cuda.device(tensor.device)` or "
            "call `torch.cuda.set_device` first."
        )


def to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """Logical NCHW tensor -> bf16 channels_last storage (no copy if already so)."""
    require_cuda(x, "to_nhwc_bf16")
    if x.dtype == BF16 and x.dim() == 4 and is_nhwc(x):
        return x
    return x.to(dtype=BF16).contiguous(memory_format=torch.channels_last)


def is_nhwc(x: torch.Tensor) -> bool:
    if x.dim() != 4:
        return False
    b, c, h, w = x.shape
    return x.stride() == (h * w * c, 1, w * c, c) or x.permute(0, 2, 3, 1).is_contiguous()


def empty_nhwc(b: int, c: int, h: int, w: int, device) -> torch.Tensor:
    return torch.empty((b, h, w, c), dtype=BF16, device=device).permute(0, 3, 1, 2)


def _check_act(x: torch.Tensor, what: str) -> None:
    require_cuda(x, what)
    if x.dtype != BF16 or not is_nhwc(x):
        raise RuntimeError(f"fmdm_b200.{what}: expected a bf16 channels_last activation, got {x.dtype} {x.stride()}")


# --------------------------------------------------------------------------------------------------------------
# conv2d implicit GEMM
# --------------------------------------------------------------------------------------------------------------
class PackedConvWeight:
    """K-major bf16 weight matrix [Cout][Ktot] for fm_conv2d_igemm_bf16, K = (segment, tap, channel).

    split: the matrix holds every segment twice - the bf16 weights, then their bf16 rounding residuals
    (`fm_weight_prepack_lo_bf16`); `conv2d` reads each source through both, i.e. multiplies by w_hi + w_lo."""

    def __init__(self, mat: torch.Tensor, seg_channels: Sequence[int], seg_ksize: Sequence[int], cout: int,
                 split: bool = False):
        self.mat = mat
        self.seg_channels = tuple(int(c) for c in seg_channels)
        self.seg_ksize = tuple(int(k) for k in seg_ksize)
        self.cout = int(cout)
        self.split = bool(split)


def pack_conv_weight(parts: Sequence[tuple], split: bool = False) -> PackedConvWeight:
    """parts: [(weight_oihw_fp32 [Cout][Cin][k][k] (or [Cout][Cin] for linear), c_begin, c_count), ...]

    Each part becomes one K segment reading `c_count` input channels starting at `c_begin` of that weight.
    split: split-bf16 weights (see `PackedConvWeight`); at most FM_CONV_MAX_SEG / 2 parts.
    """
    lib = _lib.lib()
    cout = parts[0][0].shape[0]
    seg_c, seg_k = [], []
    ktot = 0
    for w, c_begin, c_count in parts:
        ks = 1 if w.dim() == 2 else int(w.shape[-1])
        if w.shape[0] != cout:
            raise ValueError("all parts must share Cout")
        seg_c.append(int(c_count))
        seg_k.append(ks)
        ktot += ks * ks * int(c_count)
    dev = parts[0][0].device
    require_cuda(parts[0][0], "pack_conv_weight")
    split = bool(split) and 2 * len(parts) <= _lib.FM_CONV_MAX_SEG
    kall = ktot * (2 if split else 1)
    mat = torch.empty((cout, kall), dtype=BF16, device=dev)
    koff = 0
    keep = []
    for fn in ((lib.fm_weight_prepack_bf16, lib.fm_weight_prepack_lo_bf16) if split else (lib.fm_weight_prepack_bf16,)):
        for (w, c_begin, c_count), ks in zip(parts, seg_k):
            w32 = w.detach().to(dtype=torch.float32).contiguous()
            keep.append(w32)
            cin_total = w32.shape[1]
            _lib.check(
                fn(mat.data_ptr(), kall, koff, w32.data_ptr(), cout, cin_total, int(c_begin), int(c_count), ks,
                   _stream()),
                "weight_prepack",
            )
            koff += ks * ks * int(c_count)
    return PackedConvWeight(mat, seg_c, seg_k, cout, split)


def conv2d(
    srcs: Sequence[torch.Tensor],
    weight: PackedConvWeight,
    *,
    stride: int = 1,
    bias: Optional[torch.Tensor] = None,
    addvec: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    out: Optional[torch.Tensor] = None,
    want_stats: bool = False,
    norm: Optional[Sequence] = None,
    upsample_out: bool = False,
    prof_tag: Optional[str] = None,
) -> torch.Tensor:
    """Implicit-GEMM conv over the virtual channel concat of `srcs` (one K segment per source).

    upsample_out: the result is stored nearest-2x upsampled ([B, Cout, 2Ho, 2Wo]): the F.interpolate of a following
    UpsampleND folded into this conv's store (4 TMA stores of each staged tile).

    norm: per source, None or a `NormTable` slice `(table, channel_offset)` — that source is read through the fused
    operand transform act(a*x+b) (GroupNorm apply + SiLU folded into the conv; see `conv_operand_norm_ok`).

    want_stats: also emit, from the epilogue, the GroupNorm partial statistics of the output; they ride on the
    returned tensor (`out._fm_stats`) and let the consumer `group_norm` skip its statistics pass."""
    lib = _lib.lib()
    if len(srcs) != len(weight.seg_channels) or len(srcs) * (2 if weight.split else 1) > _lib.FM_CONV_MAX_SEG:
        raise ValueError(f"conv2d: {len(srcs)} sources for {len(weight.seg_channels)} weight segments")
    b, _, h, w = srcs[0].shape
    p = _lib.ConvParams()
    seg_channels, seg_ksize = weight.seg_channels, weight.seg_ksize
    if weight.split:  # every source once per weight half (hi segments first, then the residual segments)
        srcs, seg_channels, seg_ksize = list(srcs) * 2, seg_channels * 2, seg_ksize * 2
        norm = None if norm is None else list(norm) * 2
    for i, (s, c, ks) in enumerate(zip(srcs, seg_channels, seg_ksize)):
        _check_act(s, "conv2d")
        if s.shape != (b, c, h, w):
            raise ValueError(f"conv2d: segment {i} has shape {tuple(s.shape)}, expected {(b, c, h, w)}")
        p.seg[i].src = s.data_ptr()
        p.seg[i].C = c
        p.seg[i].ksize = ks
        p.seg[i].upsample = 0
        nt = None if norm is None else norm[i]
        if nt is not None:
            table, c_off = nt
            if table.ab.shape[0] != b or c_off + c > table.ab.shape[2]:
                raise ValueError("conv2d: norm table does not cover this segment")
            p.seg[i].norm_a = table.ab.data_ptr() + 4 * c_off
            p.seg[i].norm_b = table.ab.data_ptr() + 4 * (table.ab.shape[2] + c_off)
            p.seg[i].norm_stride = 2 * table.ab.shape[2]
            p.seg[i].norm_act = 1 if table.silu else 0
    p.nseg = len(srcs)
    p.B, p.H, p.W = b, h, w
    p.stride = stride
    p.Cout = weight.cout
    ho, wo = (h + stride - 1) // stride, (w + stride - 1) // stride
    oh, ow = (2 * ho, 2 * wo) if upsample_out else (ho, wo)
    if out is None:
        out = empty_nhwc(b, weight.cout, oh, ow, srcs[0].device)
    else:
        _check_act(out, "conv2d(out)")
        if tuple(out.shape) != (b, weight.cout, oh, ow):
            raise ValueError("conv2d: out shape mismatch")
    p.weight = weight.mat.data_ptr()
    p.bias = _ptr(bias)
    if addvec is not None:
        if addvec.dtype != torch.float32 or addvec.dim() != 2 or addvec.shape[0] != b or addvec.stride(1) != 1:
            raise ValueError("conv2d: addvec must be fp32 [B][>=Cout] with unit inner stride")
        p.addvec = addvec.data_ptr()
        p.addvec_stride = addvec.stride(0)
    if residual is not None:
        _check_act(residual, "conv2d(residual)")
        if tuple(residual.shape) != (b, weight.cout, ho, wo):
            raise ValueError("conv2d: residual shape mismatch")
        p.residual = residual.data_ptr()
    p.out = out.data_ptr()
    p.gn_stats = None
    p.out_upsample = int(bool(upsample_out))
    stats_ws = None
    rpi = C.c_int32(0)
    if want_stats and weight.cout % 4 == 0 and lib.fm_conv_stats_rows(C.byref(p), C.byref(rpi)) == 0:
        stats_ws = torch.empty((b * rpi.value, weight.cout // 4, 2), dtype=torch.float32, device=srcs[0].device)
        p.gn_stats = stats_ws.data_ptr()
    e0 = _prof_begin()
    _lib.check(lib.fm_conv2d_igemm_bf16(C.byref(p), _stream()), "conv2d_igemm_bf16")
    if e0 is not None:
        kind = lib.fm_conv_kernel_kind(C.byref(p))
        tag = prof_tag or {1: "conv_rolling", 2: "conv_rolling_xf"}.get(kind, "conv_tile")
        _prof_end(tag, 2.0 * b * ho * wo * weight.cout * weight.mat.shape[1], e0)
    if stats_ws is not None:
        # with upsample_out the partials describe the [Ho][Wo] result; each value appears 4x in `out`, so the
        # consumer's per-pixel count (out.H * out.W) must be divided by 4: not wired, no consumer norms an upsample
        if not upsample_out:
            out._fm_stats = (stats_ws, rpi.value)
    return out


STEM_TENSOR_MIN_PIXELS = 1 << 18  # B*H*W from which conv_in runs as im2col + 1x1 tensor-core GEMM


def stem_pack(weight_oihw: torch.Tensor) -> PackedConvWeight:
    """conv_in's weight as the K-major matrix of the tensor-core stem: w.reshape(Cout, Cin*9), zero-padded to Kp."""
    cout, cin = weight_oihw.shape[0], weight_oihw.shape[1]
    kp = (9 * cin + 7) // 8 * 8
    w2 = torch.zeros((cout, kp), dtype=torch.float32, device=weight_oihw.device)
    w2[:, :9 * cin] = weight_oihw.detach().to(torch.float32).reshape(cout, 9 * cin)
    return pack_conv_weight([(w2, 0, kp)])


def stem_im2col(x0, x1, *, in_scale=1.0, in_shift=0.0, kp: Optional[int] = None) -> torch.Tensor:
    """The 3x3 neighbourhoods of fp32 NCHW (x0 [, x1]) as a bf16 NHWC tensor [B][Kp][H][W] (logical NCHW), column
    ci*9 + kh*3 + kw = scale * x[ci][y+kh-1][x+kw-1] + shift, zero beyond 9*Cin (`fm_stem_im2col_bf16`)."""
    require_cuda(x0, "stem_im2col")
    x0 = x0.to(torch.float32).contiguous()
    b, c0, h, w = x0.shape
    c1 = 0
    if x1 is not None:
        x1 = x1.to(torch.float32).contiguous()
        c1 = x1.shape[1]
    if kp is None:
        kp = (9 * (c0 + c1) + 7) // 8 * 8
    if kp > 72:
        raise RuntimeError(f"fmdm_b200.stem_im2col: {c0 + c1} input channels (at most 8)")
    cols = empty_nhwc(b, kp, h, w, x0.device)
    e0 = _prof_begin()
    _lib.check(_lib.lib().fm_stem_im2col_bf16(x0.data_ptr(), c0, _ptr(x1), c1, float(in_scale), float(in_shift),
                                              cols.data_ptr(), b, h, w, kp, _stream()), "stem_im2col")
    _prof_end("conv_stem_im2col", 2.0 * b * h * w * kp, e0)
    return cols


def conv_stem(x0, x1, weight_oihw, bias, *, in_scale=1.0, in_shift=0.0, want_stats: bool = True,
              packed: Optional[PackedConvWeight] = None, tensor_cores: bool = False) -> torch.Tensor:
    """fp32 NCHW (x0 [, x1]) -> bf16 NHWC, 3x3 s1 p1; fuses the conditioning concat and the 2x-1 centering.

    want_stats: also emit the GroupNorm partial statistics of the output (`out._fm_stats`, as `conv2d` does).
    packed (`stem_pack(weight)`): on large inputs the conv runs on the tensor cores (bf16-rounded inputs and weights):
    `fm_conv_stem_tc_f32_bf16` in one launch for Cout 64 / 128 and Cin <= 3, otherwise `fm_stem_im2col_bf16` writes the
    3x3 neighbourhoods as a [B][H][W][Kp] bf16 tensor and the 1x1 implicit GEMM contracts it; store-bandwidth bound
    instead of fp32-FMA bound.  Small inputs (and packed=None) keep the fp32 CUDA-core kernel.
    tensor_cores=True allows the one-launch kernel without a packed matrix (the training forward)."""
    lib = _lib.lib()
    require_cuda(x0, "conv_stem")
    x0 = x0.to(torch.float32).contiguous()
    b, c0, h, w = x0.shape
    c1 = 0
    if x1 is not None:
        x1 = x1.to(torch.float32).contiguous()
        c1 = x1.shape[1]
    cout = weight_oihw.shape[0]
    tc_rows = 0
    if (packed is not None or tensor_cores) and b * h * w >= STEM_TENSOR_MIN_PIXELS:
        tc_rows = int(lib.fm_conv_stem_tc_stats_rows(b, h, w, c0 + c1, cout))
    if tc_rows > 0:  # one launch: mma.sync fragments gathered from a shared-memory halo tile
        out = empty_nhwc(b, cout, h, w, x0.device)
        stats_ws = None
        if want_stats:
            stats_ws = torch.empty((b * tc_rows, cout // 4, 2), dtype=torch.float32, device=x0.device)
        e0 = _prof_begin()
        _lib.check(
            lib.fm_conv_stem_tc_f32_bf16(
                x0.data_ptr(), c0, _ptr(x1), c1, float(in_scale), float(in_shift), weight_oihw.data_ptr(), _ptr(bias),
                out.data_ptr(), b, h, w, cout, _ptr(stats_ws), _stream(),
            ),
            "conv_stem_tc",
        )
        _prof_end("conv_stem", 2.0 * b * h * w * (c0 + c1) * 9 * cout, e0)
        if stats_ws is not None:
            out._fm_stats = (stats_ws, tc_rows)
        return out
    if packed is not None and b * h * w >= STEM_TENSOR_MIN_PIXELS and 9 * (c0 + c1) <= 72:
        kp = packed.seg_channels[0]
        cols = empty_nhwc(b, kp, h, w, x0.device)
        e0 = _prof_begin()
        _lib.check(
            lib.fm_stem_im2col_bf16(x0.data_ptr(), c0, _ptr(x1), c1, float(in_scale), float(in_shift), cols.data_ptr(),
                                    b, h, w, kp, _stream()),
            "stem_im2col",
        )
        _prof_end("conv_stem_im2col", 2.0 * b * h * w * kp, e0)
        return conv2d([cols], packed, bias=bias, want_stats=want_stats, prof_tag="conv_stem")
    out = empty_nhwc(b, cout, h, w, x0.device)
    stats_ws, rows = None, 0
    if want_stats and cout % 8 == 0:
        rows = int(lib.fm_conv_stem_stats_rows(b, h, w, cout))
        if rows > 0:
            stats_ws = torch.empty((b * rows, cout // 4, 2), dtype=torch.float32, device=x0.device)
    e0 = _prof_begin()
    _lib.check(
        lib.fm_conv_stem_f32_bf16(
            x0.data_ptr(), c0, _ptr(x1), c1, float(in_scale), float(in_shift), weight_oihw.data_ptr(), _ptr(bias),
            out.data_ptr(), b, h, w, cout, _ptr(stats_ws), _stream(),
        ),
        "conv_stem",
    )
    _prof_end("conv_stem", 2.0 * b * h * w * (c0 + c1) * 9 * cout, e0)
    if stats_ws is not None:
        out._fm_stats = (stats_ws, rows)
    return out


def conv_head(x: torch.Tensor, weight_oihw, bias, norm: Optional["NormTable"] = None) -> torch.Tensor:
    """bf16 NHWC -> fp32 NCHW, 3x3 s1 p1, Cout <= 4; `norm` folds the output GroupNorm (+SiLU) into the load."""
    lib = _lib.lib()
    _check_act(x, "conv_head")
    b, cin, h, w = x.shape
    cout = weight_oihw.shape[0]
    out = torch.empty((b, cout, h, w), dtype=torch.float32, device=x.device)
    if norm is not None and tuple(norm.ab.shape) != (b, 2, cin):
        raise ValueError("conv_head: norm table shape mismatch")
    e0 = _prof_begin()
    _lib.check(
        lib.fm_conv_head_bf16_f32(x.data_ptr(), weight_oihw.data_ptr(), _ptr(bias), out.data_ptr(), b, h, w, cin, cout,
                                  None if norm is None else norm.ab.data_ptr(),
                                  0 if norm is None else int(norm.silu), _stream()),
        "conv_head",
    )
    _prof_end("conv_head", 2.0 * b * h * w * cin * 9 * cout, e0)
    return out


# --------------------------------------------------------------------------------------------------------------
# GroupNorm (+SiLU, +scale/shift)
# --------------------------------------------------------------------------------------------------------------
class NormTable:
    """Per-(sample, channel) affine form of a GroupNorm: ab[n][0][c]*x + ab[n][1][c], then SiLU if `silu`."""

    def __init__(self, ab: torch.Tensor, silu: bool):
        self.ab, self.silu = ab, bool(silu)


def conv_operand_norm_ok(h: int, w: int, stride: int, ksizes: Sequence[int], channels: Sequence[int]) -> bool:
    """Can a conv over these sources fold the GroupNorm apply into its operand path?"""
    if any(c % 64 for c in channels):
        return False
    return bool(_lib.lib().fm_conv_operand_norm_supported(h, w, stride, int(any(k == 3 for k in ksizes))))


def group_norm_table(
    srcs: Sequence[torch.Tensor],
    groups: int,
    eps: float,
    gamma: torch.Tensor,
    beta: torch.Tensor,
    *,
    silu: bool,
    scale_shift: Optional[torch.Tensor] = None,
) -> NormTable:
    """GroupNorm statistics of the virtual concat of `srcs` -> the a*x+b table a consumer conv applies on the fly
    (nothing the size of the activation is written)."""
    lib = _lib.lib()
    if not 1 <= len(srcs) <= 2:
        raise ValueError("group_norm_table: one or two sources")
    for s in srcs:
        _check_act(s, "group_norm_table")
    x0 = srcs[0]
    x1 = srcs[1] if len(srcs) == 2 else None
    b, c0, h, w = x0.shape
    c1 = x1.shape[1] if x1 is not None else 0
    ctot = c0 + c1
    if scale_shift is not None and (scale_shift.dtype != torch.float32 or scale_shift.shape != (b, 2 * ctot)
                                    or scale_shift.stride(1) != 1):
        raise ValueError("group_norm_table: scale_shift must be fp32 [B][2C] with unit inner stride")
    ss_stride = 0 if scale_shift is None else scale_shift.stride(0)
    ab = torch.empty((b, 2, ctot), dtype=torch.float32, device=x0.device)
    st = _stream()
    e0 = _prof_begin()
    fused = [getattr(s, "_fm_stats", None) for s in srcs]
    if all(f is not None for f in fused) and (ctot // groups) % 4 == 0:
        p1 = fused[1] if len(fused) == 2 else (None, 0)
        _lib.check(
            lib.fm_groupnorm_finalize_partials_affine(
                fused[0][0].data_ptr(), fused[0][1], c0, _ptr(p1[0]), p1[1], c1, b, h * w, groups, float(eps),
                gamma.data_ptr(), beta.data_ptr(), _ptr(scale_shift), ss_stride, None, ab.data_ptr(), st),
            "groupnorm_finalize_partials_affine",
        )
        _prof_end("groupnorm_table", 0.0, e0)
    else:
        ws_elems = int(lib.fm_groupnorm_workspace_elems(b, h * w, ctot, groups))
        if ws_elems <= 0:
            raise RuntimeError(f"fmdm_b200.group_norm_table: unsupported shape B={b} HW={h * w} C={ctot} groups={groups}")
        ws = torch.empty((ws_elems,), dtype=torch.float32, device=x0.device)
        stats = torch.empty((b, groups, 2), dtype=torch.float32, device=x0.device)
        _lib.check(
            lib.fm_groupnorm_stats_bf16(x0.data_ptr(), c0, _ptr(x1), c1, b, h * w, groups, float(eps),
                                        ws.data_ptr(), stats.data_ptr(), st),
            "groupnorm_stats",
        )
        _lib.check(
            lib.fm_groupnorm_affine_f32(stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(scale_shift),
                                        ss_stride, b, ctot, groups, ab.data_ptr(), st),
            "groupnorm_affine",
        )
        _prof_end("groupnorm_stats", 2.0 * b * ctot * h * w, e0)  # one read pass, bf16
    return NormTable(ab, silu)


def group_norm(
    srcs: Sequence[torch.Tensor],
    groups: int,
    eps: float,
    gamma: torch.Tensor,
    beta: torch.Tensor,
    *,
    silu: bool,
    scale_shift: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """GroupNorm over the virtual concat of one or two NHWC bf16 sources; returns the concatenated result."""
    lib = _lib.lib()
    if not 1 <= len(srcs) <= 2:
        raise ValueError("group_norm: one or two sources")
    for s in srcs:
        _check_act(s, "group_norm")
    x0 = srcs[0]
    x1 = srcs[1] if len(srcs) == 2 else None
    b, c0, h, w = x0.shape
    c1 = x1.shape[1] if x1 is not None else 0
    ctot = c0 + c1
    st = _stream()
    e0 = _prof_begin()
    fused = [getattr(s, "_fm_stats", None) for s in srcs]
    if scale_shift is not None and (scale_shift.dtype != torch.float32 or scale_shift.shape != (b, 2 * ctot)
                                    or scale_shift.stride(1) != 1):
        raise ValueError("group_norm: scale_shift must be fp32 [B][2C] with unit inner stride")
    if all(f is not None for f in fused):
        p1 = fused[1] if len(fused) == 2 else (None, 0)
        if lib.fm_groupnorm_apply_partials_supported(fused[0][1], c0, p1[1], c1, groups):
            # small partial tables: the apply kernel folds them itself - one kernel for the whole GroupNorm (+SiLU)
            out = empty_nhwc(b, ctot, h, w, x0.device)
            _lib.check(
                lib.fm_groupnorm_apply_partials_bf16(
                    x0.data_ptr(), c0, _ptr(x1), c1, b, h * w, groups, fused[0][0].data_ptr(), fused[0][1],
                    _ptr(p1[0]), p1[1], float(eps), gamma.data_ptr(), beta.data_ptr(), _ptr(scale_shift),
                    0 if scale_shift is None else scale_shift.stride(0), int(silu), out.data_ptr(), st,
                ),
                "groupnorm_apply_partials",
            )
            _prof_end("groupnorm", 4.0 * b * ctot * h * w, e0)
            return out
    stats = torch.empty((b, groups, 2), dtype=torch.float32, device=x0.device)
    if all(f is not None for f in fused) and (ctot // groups) % 4 == 0:
        # statistics were produced by the convs that wrote the sources: fold their partial sums, no extra read pass
        p1 = fused[1] if len(fused) == 2 else (None, 0)
        _lib.check(
            lib.fm_groupnorm_finalize_partials(fused[0][0].data_ptr(), fused[0][1], c0, _ptr(p1[0]), p1[1], c1, b,
                                               h * w, groups, float(eps), stats.data_ptr(), st),
            "groupnorm_finalize_partials",
        )
    else:
        ws_elems = int(lib.fm_groupnorm_workspace_elems(b, h * w, ctot, groups))
        if ws_elems <= 0:
            raise RuntimeError(f"fmdm_b200.group_norm: unsupported shape B={b} HW={h * w} C={ctot} groups={groups}")
        ws = torch.empty((ws_elems,), dtype=torch.float32, device=x0.device)
        _lib.check(
            lib.fm_groupnorm_stats_bf16(x0.data_ptr(), c0, _ptr(x1), c1, b, h * w, groups, float(eps),
                                        ws.data_ptr(), stats.data_ptr(), st),
            "groupnorm_stats",
        )
    out = empty_nhwc(b, ctot, h, w, x0.device)
    _lib.check(
        lib.fm_groupnorm_apply_bf16(
            x0.data_ptr(), c0, _ptr(x1), c1, b, h * w, groups, stats.data_ptr(), gamma.data_ptr(),
            beta.data_ptr(), _ptr(scale_shift), 0 if scale_shift is None else scale_shift.stride(0), int(silu),
            out.data_ptr(), st,
        ),
        "groupnorm_apply",
    )
    _prof_end("groupnorm", 4.0 * b * ctot * h * w, e0)  # algorithmic bytes: read once + write once, bf16
    return out


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.lib()
    _check_act(x, "upsample_nearest2x")
    b, c, h, w = x.shape
    out = empty_nhwc(b, c, 2 * h, 2 * w, x.device)
    e0 = _prof_begin()
    _lib.check(lib.fm_upsample_nearest2x_bf16(x.data_ptr(), out.data_ptr(), b, h, w, c, _stream()), "upsample")
    _prof_end("upsample", 2.0 * b * c * h * w * 5, e0)
    return out


def transpose_bf16(x: torch.Tensor) -> torch.Tensor:
    """[B][R][C] contiguous bf16 -> [B][C][R]."""
    lib = _lib.lib()
    require_cuda(x, "transpose")
    assert x.dtype == BF16 and x.dim() == 3 and x.is_contiguous()
    b, r, c = x.shape
    out = torch.empty((b, c, r), dtype=BF16, device=x.device)
    _lib.check(lib.fm_transpose_bf16(x.data_ptr(), out.data_ptr(), b, r, c, _stream()), "transpose")
    return out


def attention(q, k, v, out, *, batch, heads, tq, tk, head_dim, q_strides, kv_strides, o_strides, scale=None):
    """Strided softmax(QK^T)V; q/k/v/out are bf16 tensors (views give the base pointers), strides in elements."""
    lib = _lib.lib()
    for t in (q, k, v, out):
        require_cuda(t, "attention")
        assert t.dtype == BF16
    if scale is None:
        scale = 1.0 / math.sqrt(head_dim)
    e0 = _prof_begin()
    _lib.check(
        lib.fm_attention_bf16(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), batch, heads, tq, tk, head_dim,
            *[int(s) for s in q_strides], *[int(s) for s in kv_strides], *[int(s) for s in o_strides], float(scale),
            _stream(),
        ),
        "attention",
    )
    _prof_end("attention", 4.0 * batch * heads * tq * tk * head_dim, e0)
    return out


def linear_attention(q, k, v, out, *, batch, heads, tq, tk, head_dim, q_strides, kv_strides, o_strides, eps=1e-6):
    """LinearQKVAttention on strided bf16 operands (same conventions as `attention`)."""
    lib = _lib.lib()
    for t in (q, k, v, out):
        require_cuda(t, "linear_attention")
        assert t.dtype == BF16
    _lib.check(
        lib.fm_linear_attention_bf16(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), batch, heads, tq, tk, head_dim,
            *[int(s) for s in q_strides], *[int(s) for s in kv_strides], *[int(s) for s in o_strides], float(eps),
            _stream(),
        ),
        "linear_attention",
    )
    return out


def context_kv(ctx: torch.Tensor, gamma, beta, weight, bias, *, groups: int, eps: float,
               channel_major: bool, return_stats: bool = False):
    """GroupNorm over the context tokens + key/value projection: ctx fp32 [B][Cc][Tc] -> bf16 [B][Tc][O]
    (or [B][O][Tc] when channel_major); weight fp32 [O][Cc]."""
    lib = _lib.lib()
    require_cuda(ctx, "context_kv")
    ctx = ctx.to(torch.float32).contiguous()
    b, cc, tc = ctx.shape
    o = weight.shape[0]
    assert weight.dtype == torch.float32 and weight.is_contiguous() and weight.shape[1] == cc
    out = torch.empty((b, o, tc) if channel_major else (b, tc, o), dtype=BF16, device=ctx.device)
    ws = torch.empty((b, groups, 2), dtype=torch.float32, device=ctx.device)
    _lib.check(
        lib.fm_context_kv_bf16(ctx.data_ptr(), gamma.data_ptr(), beta.data_ptr(), weight.data_ptr(), _ptr(bias),
                               ws.data_ptr(), out.data_ptr(), b, cc, tc, o, int(groups), float(eps),
                               int(bool(channel_major)), _stream()),
        "context_kv",
    )
    return (out, ws) if return_stats else out


# --------------------------------------------------------------------------------------------------------------
# time embedding path
# --------------------------------------------------------------------------------------------------------------
def timestep_embedding(t: Optional[torch.Tensor], dim: int, max_period: float = 10000.0, *, flip_sin_to_cos=True,
                       freq_shift: float = 0.0, batch: Optional[int] = None, t_table=None, step_dev=None):
    lib = _lib.lib()
    if t is not None:
        require_cuda(t, "timestep_embedding")
        t = t.to(torch.float32).contiguous()
        batch = t.shape[0]
        dev = t.device
    else:
        dev = t_table.device
    out = torch.empty((batch, dim), dtype=torch.float32, device=dev)
    _lib.check(
        lib.fm_timestep_embedding_f32(_ptr(t), _ptr(t_table), _ptr(step_dev), out.data_ptr(), batch, dim,
                                      float(max_period), int(bool(flip_sin_to_cos)), float(freq_shift), _stream()),
        "timestep_embedding",
    )
    return out


def linear_f32(x, weight, bias=None, bias2=None, *, silu_in=False, silu_out=False) -> torch.Tensor:
    lib = _lib.lib()
    require_cuda(x, "linear_f32")
    x = x.to(torch.float32).contiguous()
    b, i = x.shape
    o = weight.shape[0]
    assert weight.dtype == torch.float32 and weight.is_contiguous() and weight.shape[1] == i
    y = torch.empty((b, o), dtype=torch.float32, device=x.device)
    _lib.check(
        lib.fm_linear_f32(x.data_ptr(), weight.data_ptr(), _ptr(bias), _ptr(bias2), y.data_ptr(), b, i, o,
                          int(silu_in), int(silu_out), _stream()),
        "linear_f32",
    )
    return y


# --------------------------------------------------------------------------------------------------------------
# scheduler steps
# --------------------------------------------------------------------------------------------------------------
def sched_flowmatch(x, v, coef, step, x_out=None, step_dev=None):
    lib = _lib.lib()
    require_cuda(x, "sched_flowmatch")
    assert x.dtype == torch.float32 and v.dtype == torch.float32 and x.is_contiguous() and v.is_contiguous()
    if x_out is None:
        x_out = torch.empty_like(x)
    _lib.check(
        lib.fm_sched_flowmatch_f32(x_out.data_ptr(), x.data_ptr(), v.data_ptr(), coef.data_ptr(), _ptr(step_dev),
                                   int(step), x.numel(), _stream()),
        "sched_flowmatch",
    )
    return x_out


def sched_ddim(x, eps, coef, step, clip, clip_range, x_out=None, step_dev=None):
    lib = _lib.lib()
    require_cuda(x, "sched_ddim")
    assert x.dtype == torch.float32 and eps.dtype == torch.float32 and x.is_contiguous() and eps.is_contiguous()
    if x_out is None:
        x_out = torch.empty_like(x)
    _lib.check(
        lib.fm_sched_ddim_f32(x_out.data_ptr(), x.data_ptr(), eps.data_ptr(), coef.data_ptr(), _ptr(step_dev),
                              int(step), int(bool(clip)), float(clip_range), x.numel(), _stream()),
        "sched_ddim",
    )
    return x_out


def sched_ddpm(x, eps, noise, coef, step, clip, clip_range, x_out=None, step_dev=None):
    lib = _lib.lib()
    require_cuda(x, "sched_ddpm")
    for t in (x, eps, noise):
        assert t.dtype == torch.float32 and t.is_contiguous()
    if x_out is None:
        x_out = torch.empty_like(x)
    _lib.check(
        lib.fm_sched_ddpm_f32(x_out.data_ptr(), x.data_ptr(), eps.data_ptr(), noise.data_ptr(), coef.data_ptr(),
                              _ptr(step_dev), int(step), int(bool(clip)), float(clip_range), x.numel(), _stream()),
        "sched_ddpm",
    )
    return x_out


def sched_dpmpp2m(x, eps, m_prev, coef, step, x_out=None, m_cur=None, step_dev=None):
    lib = _lib.lib()
    require_cuda(x, "sched_dpmpp2m")
    for t in (x, eps, m_prev):
        assert t.dtype == torch.float32 and t.is_contiguous()
    if x_out is None:
        x_out = torch.empty_like(x)
    if m_cur is None:
        m_cur = torch.empty_like(x)
    _lib.check(
        lib.fm_sched_dpmpp2m_f32(x_out.data_ptr(), m_cur.data_ptr(), x.data_ptr(), eps.data_ptr(), m_prev.data_ptr(),
                                 coef.data_ptr(), _ptr(step_dev), int(step), x.numel(), _stream()),
        "sched_dpmpp2m",
    )
    return x_out, m_cur


def sched_unipc(x, eps, last, m1, m2, coef, step, x_out=None, step_dev=None):
    """One UniPC step (corrector + predictor); `last`, `m1`, `m2` are updated in place."""
    lib = _lib.lib()
    require_cuda(x, "sched_unipc")
    for t in (x, eps, last, m1, m2):
        assert t.dtype == torch.float32 and t.is_contiguous()
    if x_out is None:
        x_out = torch.empty_like(x)
    _lib.check(
        lib.fm_sched_unipc_f32(x_out.data_ptr(), last.data_ptr(), m1.data_ptr(), m2.data_ptr(), x.data_ptr(),
                               eps.data_ptr(), coef.data_ptr(), _ptr(step_dev), int(step), x.numel(), _stream()),
        "sched_unipc",
    )
    return x_out


def sched_add_noise(x0, noise, a, b):
    lib = _lib.lib()
    require_cuda(x0, "sched_add_noise")
    x0 = x0.to(torch.float32).contiguous()
    noise = noise.to(torch.float32).contiguous()
    out = torch.empty_like(x0)
    bsz = x0.shape[0]
    _lib.check(
        lib.fm_sched_add_noise_f32(out.data_ptr(), x0.data_ptr(), noise.data_ptr(), a.data_ptr(), b.data_ptr(), bsz,
                                   x0.numel() // max(bsz, 1), _stream()),
        "sched_add_noise",
    )
    return out


def counter_add(ctr: torch.Tensor, delta: int) -> None:
    _lib.check(_lib.lib().fm_counter_add(ctr.data_ptr(), int(delta), _stream()), "counter_add")


def clamp_f32(x: torch.Tensor, lo: float, hi: float) -> torch.Tensor:
    require_cuda(x, "clamp")
    x = x.contiguous()
    y = torch.empty_like(x)
    _lib.check(_lib.lib().fm_clamp_f32(y.data_ptr(), x.data_ptr(), float(lo), float(hi), x.numel(), _stream()),
               "clamp")
    return y


def launch_count() -> int:
    return int(_lib.lib().fm_launch_count())
