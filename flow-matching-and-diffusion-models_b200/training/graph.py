"""Differentiable forward of the denoiser for the training step (BASELINE config 5).

`BaseUNetND.forward` routes here when autograd is recording and the model is in training mode: the same modules and
parameters (`state_dict` unchanged), but every op is one of `training.functions` so that `loss.backward()` runs the
hand-written backward kernels.  The inference-only fusions that have no backward (operand-transform GroupNorm, the
batched time-embedding projection, the nearest-2x folded into a producer's store, CUDA-graph replay) are not used:
the normalised activations are materialised because the weight-gradient GEMMs need them anyway.

Reference call chain restated: `unet_diffusers_nd.py:146-191`, `legacy_unet.py:60-160`, `residual.py:92-121`,
`attention.py:220-274`, `upsampling.py:24-62`."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .._runtime import out_of_scope
from . import functions as F


def supported(model) -> bool:
    from ..models.unet.unet import EfficientUNetND
    from ..models.unet.unet_diffusers_nd import UNetDiffusersND

    if isinstance(model, UNetDiffusersND):
        if model.spatial_dims != 2:
            return False
        if model.cross_attention_dim is None:
            return True
        # conditioning: "attention" (`flow_matching_lib.py:159-164`): cross-attention trains at head_dim 8 with a
        # context of <= 16 channels (the LDCT latents have 4)
        from ..nn.blocks.attention import CONTEXT_DIM_MAX, DiffusersAttentionND

        cross = [m for m in model.modules() if isinstance(m, DiffusersAttentionND) and m.context_dim is not None]
        return all(m.head_dim == 8 and m.context_dim <= CONTEXT_DIM_MAX and m.dropout == 0 for m in cross)
    if isinstance(model, EfficientUNetND):
        from ..nn.blocks.attention import CONTEXT_DIM_MAX, SpatialCrossAttention, SpatialSelfAttention

        if model.spatial_dims != 2 or model.dropout != 0:
            return False
        for m in model.modules():   # softmax or linear attention, self or over a context of <= 16 channels
            if isinstance(m, (SpatialSelfAttention, SpatialCrossAttention)) and m.dim_head not in (8, 16, 32, 64):
                return False
            if isinstance(m, SpatialCrossAttention) and m.context_dim > CONTEXT_DIM_MAX:
                return False
        return True
    return False


def _gn(norm: nn.GroupNorm, x, *, silu: bool, scale_shift=None):
    return F.group_norm(x, norm.weight, norm.bias, groups=norm.num_groups, eps=norm.eps, silu=silu,
                        scale_shift=scale_shift)


class BackwardCut:
    """Splits the backward pass of one training forward into two stages so that the data-parallel gradient all-reduce
    of the first stage's parameters runs while the second stage still computes (`training.step`): every tensor that
    flows from the early part of the network (stage 2 of the backward: stem, time MLP, the first down blocks) into
    the late part (stage 1: the remaining down blocks, mid block, up path, head) is replaced by a detached leaf;
    `loss.backward()` then stops at the leaves, and `finish()` continues from their gradients.

    A UNet's parameters sit almost entirely in its low-resolution levels while its backward time sits in the
    high-resolution ones, so cutting after the high-resolution down blocks leaves ~95 % of the gradient bytes ready
    with ~40 % of the backward still to run."""

    def __init__(self):
        self.outer, self.inner, self._by_id = [], [], {}

    def cross(self, t: torch.Tensor) -> torch.Tensor:
        leaf = self._by_id.get(id(t))
        if leaf is None:
            leaf = t.detach().requires_grad_(True)
            leaf._fm_fresh_leaf = True                   # made in this forward: may carry a gradient slot
            stats = getattr(t, "_fm_stats", None)        # GroupNorm partial statistics the producer conv left
            if stats is not None:
                leaf._fm_stats = stats
            self._by_id[id(t)] = leaf
            self.outer.append(t)
            self.inner.append(leaf)
        return leaf

    def finish(self) -> None:
        """Stage 2: backward of the early part from the gradients `loss.backward()` left on the leaves."""
        outs = [o for o, l in zip(self.outer, self.inner) if l.grad is not None]
        grads = [l.grad for l in self.inner if l.grad is not None]
        self.outer, self.inner, self._by_id = [], [], {}
        if outs:
            torch.autograd.backward(outs, grads)


def finish_backward(model) -> None:
    """Run the second backward stage of `model`'s last training forward, if that forward was cut."""
    cut = model.__dict__.pop("_fm_cut_state", None)
    if cut is not None:
        cut.finish()


def _embedding_blocks(modules):
    from ..nn.blocks.residual import ResBlockND

    return [m for mod in modules for m in mod.modules() if isinstance(m, ResBlockND) and m.uses_embedding
            and (m.use_scale_shift_norm or m.add_embedding_to_hidden)]


def _cut_groups(model, cut_at):
    """(early, late) module lists of a two-stage backward (`BackwardCut`); without a cut everything is "late"."""
    if hasattr(model, "down_blocks"):
        down = list(model.down_blocks)
        rest = ([model.mid_block] if getattr(model, "mid_block", None) is not None else []) + list(model.up_blocks)
        if cut_at is None or not 1 <= cut_at < len(down):
            return [], down + rest
        return down[:cut_at], down[cut_at:] + rest
    blocks_in = list(model.input_blocks)
    rest = [model.middle_block] + list(model.output_blocks)
    if cut_at is None or not 1 <= cut_at < len(blocks_in):
        return [], blocks_in + rest
    return blocks_in[:cut_at], blocks_in[cut_at:] + rest


def flat_param_order(model, cut_at=None):
    """Layout of the trainers' flat parameter / gradient buffers: reverse registration order (gradients complete
    roughly front to back during the backward), except that parameters the step uses as ONE matrix are made adjacent
    so that `functions.fused_param` can view them in place: the `emb_layers` weights (then biases) of all ResBlocks of
    a backward stage in block order - the late stage's at the front of the buffer, the early stage's at the end - and
    the to_q / to_k / to_v weights (then biases) of every attention block."""
    from ..nn.blocks.attention import DiffusersAttentionND

    params = [p for p in model.parameters() if p.requires_grad]
    early, late = _cut_groups(model, cut_at)
    front, back, taken = [], [], set()

    def emb_group(mods, dest):
        groups = {}
        for blk in _embedding_blocks(mods):
            groups.setdefault(bool(blk.emb_activation_before_proj), []).append(blk)
        for blks in groups.values():
            if any(b.emb_layers.bias is None for b in blks):
                continue
            for plist in ([b.emb_layers.weight for b in blks], [b.emb_layers.bias for b in blks]):
                if all(q.requires_grad and id(q) not in taken for q in plist):
                    dest.extend(plist)
                    taken.update(id(q) for q in plist)

    emb_group(late, front)
    emb_group(early, back)
    trios = {}
    for m in model.modules():
        if isinstance(m, DiffusersAttentionND):
            cross = getattr(m, "context_dim", None) is not None   # cross-attention: only K | V are one matrix
            wts = [m.to_k.weight, m.to_v.weight] if cross else [m.to_q.weight, m.to_k.weight, m.to_v.weight]
            bss = [m.to_k.bias, m.to_v.bias] if cross else [m.to_q.bias, m.to_k.bias, m.to_v.bias]
            for plist in (wts, bss):
                if all(q is not None and q.requires_grad and id(q) not in taken for q in plist):
                    for q in plist:
                        trios[id(q)] = plist
                    taken.update(id(q) for q in plist)
    middle, emitted = [], set()
    for q in reversed(params):
        if id(q) in trios:
            if id(q) not in emitted:
                middle.extend(trios[id(q)])
                emitted.update(id(t) for t in trios[id(q)])
        elif id(q) not in taken:
            middle.append(q)
    return front + middle + back


class EmbProjections:
    """Every ResBlock's `emb_layers` projection of the time embedding in ONE launch (`residual.py:99-108` runs one
    Linear per block on the same input): the weights are concatenated per step, the result is split into per-block
    column views, and the backward is one dX / dW / db GEMM over the concatenation.  `blocks`: restrict to these
    ResBlocks (one projection batch per backward stage, see `BackwardCut`)."""

    def __init__(self, model, emb: torch.Tensor, blocks=None):
        if blocks is None:
            blocks = _embedding_blocks([model])
        self.slices = {}
        groups = {}
        for blk in blocks:
            groups.setdefault(bool(blk.emb_activation_before_proj), []).append(blk)
        for silu_in, blks in groups.items():
            sizes = [(b.emb_layers.weight.shape[0] + 3) // 4 * 4 for b in blks]  # 16-byte aligned column slices
            # under a trainer the projection weights of a stage are adjacent in the flat parameter buffer
            # (`flat_param_order`): the concatenation is a view, its gradient a view of the flat gradient
            wf = bf = None
            if emb.shape[0] <= 32 and all(n == b.emb_layers.weight.shape[0] for b, n in zip(blks, sizes)) \
                    and all(b.emb_layers.bias is not None for b in blks):
                wf = F.fused_param([b.emb_layers.weight for b in blks])
                bf = F.fused_param([b.emb_layers.bias for b in blks]) if wf is not None else None
            if wf is not None and bf is not None:
                e_all = F.linear(emb, wf, bf, silu_in=silu_in)
                for b, e in zip(blks, F.split_cols(e_all, sizes)):
                    self.slices[id(b)] = e
                continue
            ws, bs = [], []
            for b, n in zip(blks, sizes):
                w, bias = b.emb_layers.weight, b.emb_layers.bias
                pad = n - w.shape[0]
                if pad:
                    w = torch.cat([w, w.new_zeros(pad, w.shape[1])], 0)
                    bias = torch.cat([bias, bias.new_zeros(pad)], 0)
                ws.append(w)
                bs.append(bias)
            e_all = F.linear(emb, torch.cat(ws, 0), torch.cat(bs, 0), silu_in=silu_in)
            for b, e in zip(blks, F.split_cols(e_all, sizes)):
                self.slices[id(b)] = e[:, :b.emb_layers.weight.shape[0]]

        self.raw = emb

    def get(self, blk):
        return self.slices.get(id(blk))


def resblock(blk, x: torch.Tensor, emb) -> torch.Tensor:
    """`residual.py:92-121` (GroupNorm / SiLU variant).  `emb`: the time embedding or an `EmbProjections`."""
    if not (blk.spatial_dims == 2 and blk.norm_type == "gn" and blk.act_name in ("silu", "swish")) \
            or blk.dropout > 0:
        out_of_scope(f"training ResBlockND(norm={blk.norm_type}, act={blk.act_name}, dropout={blk.dropout})")
        raise RuntimeError("fmdm_b200.training: unsupported ResBlockND variant")
    c, oc = blk.channels, blk.out_channels
    xs = list(x) if isinstance(x, (tuple, list)) else [x]   # (hidden, skip): the decoder concat stays virtual
    if len(xs) == 2 and any(t.shape[1] % 8 for t in xs):
        xs = [torch.cat(xs, 1)]                               # the two-source kernels take 8-channel granules
    h = _gn(blk.norm1, xs if len(xs) == 2 else xs[0], silu=True)
    addvec = scale_shift = None
    if blk.uses_embedding:
        if emb is None:
            raise ValueError("ResBlockND expects `emb` when emb_channels is set.")
        e = emb.get(blk) if isinstance(emb, EmbProjections) else None
        if e is None:
            raw = emb.raw if isinstance(emb, EmbProjections) else emb
            e = F.linear(raw, blk.emb_layers.weight, blk.emb_layers.bias, silu_in=blk.emb_activation_before_proj)
        if blk.use_scale_shift_norm:
            scale_shift = e
        elif blk.add_embedding_to_hidden:
            addvec = e
    h = F.conv([h], [(blk.conv1.conv.weight, 0, c)], bias=blk.conv1.conv.bias, addvec=addvec)
    h = _gn(blk.norm2, h, silu=True, scale_shift=scale_shift)
    w2, b2 = blk.conv2.conv.weight, blk.conv2.conv.bias
    if isinstance(blk.skip_connection, nn.Identity):
        return F.conv([h], [(w2, 0, oc)], bias=b2, residual=xs[0] if len(xs) == 1 else torch.cat(xs, 1))
    skip = blk.skip_connection.conv
    ws = skip.weight if skip.weight.shape[-1] == 3 else skip.weight.reshape(oc, c)
    bias = b2
    if skip.bias is not None:
        bias = skip.bias if b2 is None else b2 + skip.bias
    segs, off = [(w2, 0, oc)], 0
    for t in xs:
        segs.append((ws, off, t.shape[1]))
        off += t.shape[1]
    return F.conv([h] + xs, segs, bias=bias)


def cross_attention(att, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
    """`attention.py:232-274` with `context_dim`: q from the image tokens, k / v from the normalised context."""
    from ..nn.blocks.attention import CONTEXT_DIM_MAX, context_tokens

    b, c, hh, ww = x.shape
    if context is None:
        raise ValueError("DiffusersAttentionND cross-attention requires a non-empty context tensor.")
    if att.head_dim != 8 or att.context_dim > CONTEXT_DIM_MAX or att.dropout > 0:
        out_of_scope("training DiffusersAttentionND cross-attention (head_dim != 8 / context_dim > 16 / dropout)")
    n = _gn(att.group_norm, x, silu=False)
    q = F.conv([n], [(att.to_q.weight, 0, c)], bias=att.to_q.bias)
    wkv = F.fused_param([att.to_k.weight, att.to_v.weight])
    bkv = F.fused_param([att.to_k.bias, att.to_v.bias]) if wkv is not None else None
    if wkv is None or bkv is None:
        wkv = torch.cat([att.to_k.weight, att.to_v.weight], 0)
        bkv = torch.cat([att.to_k.bias, att.to_v.bias], 0)
    cn = att.context_norm
    kv = F.context_kv(context_tokens(context, att.context_dim), cn.weight, cn.bias, wkv, bkv, groups=cn.num_groups,
                      eps=cn.eps)
    a = F.cross_attention(q, kv, att.heads)
    return F.conv([a], [(att.to_out[0].weight, 0, c)], bias=att.to_out[0].bias, residual=x)


def attention(att, x: torch.Tensor, context=None) -> torch.Tensor:
    """`attention.py:220-274` (self-attention; with `context_dim`: cross-attention over `context`)."""
    b, c, hh, ww = x.shape
    if att.context_dim is not None:
        return cross_attention(att, x, context)
    if att.head_dim not in (8, 16, 32, 64) or att.dropout > 0:
        out_of_scope("training DiffusersAttentionND (dropout / head_dim)")
        raise RuntimeError("fmdm_b200.training: unsupported attention variant")
    gn = att.group_norm
    n = _gn(gn, x, silu=False)
    w = F.fused_param([att.to_q.weight, att.to_k.weight, att.to_v.weight])
    bias = F.fused_param([att.to_q.bias, att.to_k.bias, att.to_v.bias]) if w is not None else None
    if w is None or bias is None:
        w = torch.cat([att.to_q.weight, att.to_k.weight, att.to_v.weight], 0)
        bias = torch.cat([att.to_q.bias, att.to_k.bias, att.to_v.bias], 0)
    qkv = F.conv([n], [(w, 0, c)], bias=bias)
    a = F.attention_qkv(qkv, att.heads)
    return F.conv([a], [(att.to_out[0].weight, 0, c)], bias=att.to_out[0].bias, residual=x)


def spatial_self_attention(att, x: torch.Tensor) -> torch.Tensor:
    """`attention.py:82-117` (CompVis block: GN -> Conv1d qkv -> raw-reshape head split -> SDPA -> Conv1d -> + x)."""
    b, c, hh, ww = x.shape
    if att.dim_head not in (8, 16, 32, 64):
        out_of_scope("training SpatialSelfAttention (dim_head)")
        raise RuntimeError("fmdm_b200.training: unsupported SpatialSelfAttention variant")
    t, inner = hh * ww, att.inner_dim
    n = _gn(att.norm, x, silu=False)
    qkv = F.conv([n], [(att.qkv.weight.squeeze(-1), 0, c)], bias=att.qkv.bias)          # NHWC == [b][T][3*inner]
    qkv_cm = F.transpose(qkv.permute(0, 2, 3, 1).reshape(b, t, 3 * inner))              # [b][3*inner][T]
    a = F.attention_raw(qkv_cm, att.heads, att.dim_head, linear=att.use_linear)          # [b][heads][T][dh]
    h_tc = F.transpose(a.view(b, inner, t))                                              # raw reshape, [b][T][inner]
    h = h_tc.view(b, hh, ww, inner).permute(0, 3, 1, 2)
    return F.conv([h], [(att.proj_out.weight.squeeze(-1), 0, inner)], bias=att.proj_out.bias, residual=x)


def spatial_cross_attention(att, x: torch.Tensor, context) -> torch.Tensor:
    """`attention.py:120-189` (CompVis cross-attention: GN -> Conv1d q -> raw reshape; context GN -> Conv1d kv -> raw
    reshape -> softmax or linear attention with Tq != Tk -> raw reshape back -> Conv1d -> + x)."""
    from ..nn.blocks.attention import CONTEXT_DIM_MAX, context_tokens

    if context is None:
        raise ValueError("SpatialCrossAttention requires a non-empty context tensor.")
    b, c, hh, ww = x.shape
    if att.dim_head not in (8, 16, 32, 64) or att.context_dim > CONTEXT_DIM_MAX:
        out_of_scope("training SpatialCrossAttention (dim_head / context_dim > 16)")
    t, inner = hh * ww, att.inner_dim
    n = _gn(att.norm, x, silu=False)
    q = F.conv([n], [(att.q_proj.weight.squeeze(-1), 0, c)], bias=att.q_proj.bias)       # NHWC == [b][T][inner]
    q_cm = F.transpose(q.permute(0, 2, 3, 1).reshape(b, t, inner))                       # [b][inner][T]
    cn = att.context_norm
    kv = F.context_kv(context_tokens(context, att.context_dim), cn.weight, cn.bias, att.kv_proj.weight.squeeze(-1),
                      att.kv_proj.bias, groups=cn.num_groups, eps=cn.eps)                # [b][Tc][2*inner]
    kv_cm = F.transpose(kv)                                                              # [b][2*inner][Tc]
    a = F.cross_attention_raw(q_cm, kv_cm, att.heads, att.dim_head, linear=att.use_linear)
    h_tc = F.transpose(a.view(b, inner, t))                                              # raw reshape, [b][T][inner]
    h = h_tc.view(b, hh, ww, inner).permute(0, 3, 1, 2)
    return F.conv([h], [(att.proj_out.weight.squeeze(-1), 0, inner)], bias=att.proj_out.bias, residual=x)


def downsample(down, x):
    if not down.use_conv:
        out_of_scope("training DownsampleND(use_conv=False)")
        raise RuntimeError("fmdm_b200.training: unsupported downsampler")
    conv = down.op.conv
    return F.conv([x], [(conv.weight, 0, down.channels)], bias=conv.bias, stride=2)


def upsample(up, x):
    y = F.upsample_nearest2x(x)
    if not up.use_conv:
        return y
    conv = up.conv.conv
    return F.conv([y], [(conv.weight, 0, up.channels)], bias=conv.bias)


def _sequential(seq, x, emb, context_ca=None):
    """`TimestepEmbedSequential.forward` (`unet.py:18-39`) over differentiable ops."""
    from ..nn.blocks.attention import SpatialCrossAttention, SpatialSelfAttention
    from ..nn.blocks.residual import ResBlockND
    from ..nn.ops.upsampling import DownsampleND, UpsampleND

    for layer in seq:
        if isinstance(layer, ResBlockND):
            x = resblock(layer, x, emb)
        elif isinstance(layer, SpatialSelfAttention):
            x = spatial_self_attention(layer, x)
        elif isinstance(layer, SpatialCrossAttention):
            x = spatial_cross_attention(layer, x, context_ca)
        elif isinstance(layer, DownsampleND):
            x = downsample(layer, x)
        elif isinstance(layer, UpsampleND):
            x = upsample(layer, x)
        else:
            out_of_scope(f"training {type(layer).__name__} inside EfficientUNetND")
            raise RuntimeError(f"fmdm_b200.training: no training path for {type(layer).__name__}")
    return x


def efficient_unet_forward(model, x: torch.Tensor, t, context=None, context_ca=None) -> torch.Tensor:
    """`EfficientUNetND.forward` (`unet.py:295-326`) with autograd."""
    ops.require_cuda(x, "training.efficient_unet_forward")
    t = model._normalize_timesteps(t, x)
    feats = ops.timestep_embedding(t, model.model_channels, 10000.0, flip_sin_to_cos=False, freq_shift=0.0)
    emb = F.linear(feats, model.time_embed[0].weight, model.time_embed[0].bias)
    emb = F.linear(emb, model.time_embed[2].weight, model.time_embed[2].bias, silu_in=True)
    cut_at = model.__dict__.get("_fm_backward_cut")
    blocks_in = list(model.input_blocks)
    if cut_at is not None and not 1 <= cut_at < len(blocks_in):
        cut_at = None
    raw_emb = emb
    early_mods, late_mods = _cut_groups(model, cut_at)   # the same grouping `flat_param_order` lays the weights out by
    emb = EmbProjections(model, raw_emb, _embedding_blocks(late_mods if cut_at is None else early_mods))
    stem = model.input_blocks[0][0].conv
    cin = x.shape[1] + (context.shape[1] if context is not None else 0)
    if cin != stem.in_channels:
        raise ValueError(f"EfficientUNetND expected {stem.in_channels} input channels, got {cin}")
    if cin > 4:
        out_of_scope(f"training stem conv with {cin} input channels")
        raise RuntimeError("fmdm_b200.training: the stem backward supports up to 4 input channels")
    h = F.conv_stem(x, context, stem.weight, stem.bias)
    hs = [h]
    for i, block in enumerate(blocks_in[1:], start=1):
        if i == cut_at:
            cut = model.__dict__["_fm_cut_state"] = BackwardCut()
            hs = [cut.cross(t) for t in hs]
            h = hs[-1]
            emb = EmbProjections(model, cut.cross(raw_emb), _embedding_blocks(late_mods))
        h = _sequential(block, h, emb, context_ca)
        hs.append(h)
    h = _sequential(model.middle_block, h, emb, context_ca)
    for block in model.output_blocks:
        h = _sequential(block, (h, hs.pop()), emb, context_ca)   # `unet.py:321-322`, concat kept virtual
    h = _gn(model.out[0], h, silu=True)
    head = model.out[2].conv
    return F.conv_head(h, head.weight, head.bias)


def unet_forward(model, x: torch.Tensor, t, context=None, context_ca=None) -> torch.Tensor:
    """`UNetDiffusersND.forward` / `EfficientUNetND.forward` with autograd: returns the fp32 NCHW prediction."""
    from ..models.unet.unet import EfficientUNetND
    from .packplan import PackPlan

    if not supported(model):
        out_of_scope(f"training {type(model).__name__}")
        raise RuntimeError("fmdm_b200.training: no training path for this denoiser variant (2-D, self-attention only)")
    # every conv-weight pack of the step in one launch (recorded on the first step, see training.packplan)
    plan = model.__dict__.get("_fm_pack_plan")
    if plan is None:
        plan = model.__dict__["_fm_pack_plan"] = PackPlan()
    plan.begin_step(x.device)
    F.ACTIVE_PLAN = plan
    model.__dict__.pop("_fm_cut_state", None)
    del F._OPEN_SLOTS[:]
    if isinstance(model, EfficientUNetND):
        return efficient_unet_forward(model, x, t, context, context_ca)
    ops.require_cuda(x, "training.unet_forward")
    t = model._normalize_timesteps(t, x)
    feats = ops.timestep_embedding(t, model.time_proj_dim, 10000.0, flip_sin_to_cos=model.flip_sin_to_cos,
                                   freq_shift=float(model.freq_shift))
    te = model.time_embedding
    emb = F.linear(feats, te.linear_1.weight, te.linear_1.bias)
    raw_emb = F.linear(emb, te.linear_2.weight, te.linear_2.bias, silu_in=True)  # SiLU between the two layers
    down = list(model.down_blocks)
    cut_at = model.__dict__.get("_fm_backward_cut")     # number of down blocks in the second backward stage
    if cut_at is not None and not 1 <= cut_at < len(down):
        cut_at = None
    early_mods, late_mods = _cut_groups(model, cut_at)   # the same grouping `flat_param_order` lays the weights out by
    emb = EmbProjections(model, raw_emb, _embedding_blocks(late_mods if cut_at is None else early_mods))

    scale, shift = (2.0, -1.0) if model.center_input_sample else (1.0, 0.0)
    cin = x.shape[1] + (context.shape[1] if context is not None else 0)
    if cin != model.conv_in.in_channels:
        raise ValueError(f"UNetDiffusersND expected {model.conv_in.in_channels} input channels, got {cin}")
    if cin > 4:
        out_of_scope(f"training stem conv with {cin} input channels")
        raise RuntimeError("fmdm_b200.training: the stem backward supports up to 4 input channels")
    sample = F.conv_stem(x, context, model.conv_in.weight, model.conv_in.bias, in_scale=scale, in_shift=shift)

    skips = [sample]
    for bi, block in enumerate(down):
        if bi == cut_at:
            cut = model.__dict__["_fm_cut_state"] = BackwardCut()
            skips = [cut.cross(t) for t in skips]
            sample = skips[-1]
            emb = EmbProjections(model, cut.cross(raw_emb), _embedding_blocks(late_mods))
        for i, res in enumerate(block.resnets):
            sample = resblock(res, sample, emb)
            if block.attentions is not None:
                sample = attention(block.attentions[i], sample, context_ca)
            skips.append(sample)
        if block.downsamplers is not None:
            for d in block.downsamplers:
                sample = downsample(d, sample)
            skips.append(sample)
    mid = model.mid_block
    if mid is not None:
        sample = resblock(mid.resnets[0], sample, emb)
        if mid.attentions is not None:
            sample = attention(mid.attentions[0], sample, context_ca)
        sample = resblock(mid.resnets[1], sample, emb)
    for block in model.up_blocks:
        for i, res in enumerate(block.resnets):
            skip = skips.pop()
            sample = resblock(res, (sample, skip), emb)  # `legacy_unet.py:150`, concat kept virtual
            if block.attentions is not None:
                sample = attention(block.attentions[i], sample, context_ca)
        if block.upsamplers is not None:
            for u in block.upsamplers:
                sample = upsample(u, sample)
    sample = _gn(model.conv_norm_out, sample, silu=True)
    return F.conv_head(sample, model.conv_out.weight, model.conv_out.bias)
