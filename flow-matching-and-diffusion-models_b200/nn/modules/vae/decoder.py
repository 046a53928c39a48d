"""Decoder — drop-in for `src/nn/modules/vae/decoder.py:20-160` (the Stable-Diffusion-style VAE decoder that turns the
latent sampler's output back into images, BASELINE config 3) on the B200 kernels.

Same constructor keywords, child names and `state_dict` keys (`conv_in.conv`, `mid_block1`, `mid_attn.{norm,qkv,
proj_out}`, `mid_block2`, `ups.{i}.blocks.{j}`, `ups.{i}.attns.{j}`, `ups.{i}.up.conv.conv`, `norm_out`,
`conv_out.conv`).  Kernel schedule: stem kernel (tiny Cin; optionally with the 1x1 `post_quant_conv` of
`AutoencoderKL.decode` composed into its weights) -> ResBlockND / SpatialSelfAttention / UpsampleND exactly as in the
denoisers (GroupNorm folded into the convs on rows >= 65 px, nearest-2x folded into the producer's store) -> head
kernel with `norm_out` + SiLU folded into its load."""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import ops
from ...._runtime import ParamCache, f32, out_of_scope
from ...blocks.attention import SpatialSelfAttention
from ...blocks.residual import ResBlockND
from ...ops.convolution import ConvND
from ...ops.normalization import fused_group_norm, fused_group_norm_table
from ...ops.upsampling import UpsampleND


class Decoder(nn.Module):
    def __init__(self, out_ch: int = 3, base_ch: int = 128, ch_mult: Tuple[int, ...] = (1, 2, 4, 4),
                 down_channels: Optional[Tuple[int, ...]] = None, num_res_blocks: int = 2,
                 attn_resolutions: Tuple[int, ...] = (), resolution: int = 256, z_channels: int = 4,
                 dropout: float = 0.0, use_attention: bool = True, attn_heads: Optional[int] = None,
                 attn_dim_head: Optional[int] = None, tanh_out: bool = False, spatial_dims: int = 2,
                 emb_channels: Optional[int] = None, use_scale_shift_norm: bool = False,
                 norm_groups: Optional[int] = None, block_factory=None) -> None:
        super().__init__()
        if emb_channels is None and use_scale_shift_norm:
            raise ValueError("use_scale_shift_norm requires emb_channels to be provided.")
        self.resolution, self.tanh_out, self.spatial_dims = resolution, tanh_out, spatial_dims
        self.emb_channels, self.use_attention = emb_channels, use_attention
        self.attn_heads, self.attn_dim_head = attn_heads, attn_dim_head
        self.use_scale_shift_norm = use_scale_shift_norm and emb_channels is not None
        channels = tuple(down_channels) if down_channels is not None else tuple(base_ch * m for m in ch_mult)
        block_in = channels[-1]
        curr_res = resolution // (2 ** (len(channels) - 1))

        def block(cin, cout):
            factory = block_factory or ResBlockND
            return factory(channels=cin, emb_channels=emb_channels, dropout=dropout, out_channels=cout, use_conv=False,
                           use_scale_shift_norm=self.use_scale_shift_norm, spatial_dims=spatial_dims)

        self.conv_in = ConvND(spatial_dims, z_channels, block_in, 3, padding=1)
        self.mid_block1 = block(block_in, block_in)
        self.mid_attn = self._build_attention_layer(block_in) if use_attention else nn.Identity()
        self.mid_block2 = block(block_in, block_in)
        ups: List[nn.Module] = []
        in_ch = block_in
        for idx, out_ch_stage in enumerate(reversed(channels)):
            blocks, attns = [], []
            for _ in range(num_res_blocks + 1):
                blocks.append(block(in_ch, out_ch_stage))
                in_ch = out_ch_stage
                if use_attention and (curr_res in attn_resolutions):
                    attns.append(self._build_attention_layer(in_ch))
            stage = nn.Module()
            stage.blocks = nn.ModuleList(blocks)
            stage.attns = nn.ModuleList(attns)
            if idx != len(channels) - 1:
                stage.up = UpsampleND(spatial_dims, in_ch, use_conv=True)
                curr_res *= 2
            ups.insert(0, stage)  # stored in encoder order, executed reversed (reference :116)
        self.ups = nn.ModuleList(ups)
        groups = norm_groups if norm_groups is not None else max(1, math.gcd(in_ch, 32))
        self.norm_out = nn.GroupNorm(groups, in_ch)
        self.conv_out = ConvND(spatial_dims, in_ch, out_ch, 3, padding=1)
        self._cache = ParamCache()

    def _build_attention_layer(self, channels: int) -> nn.Module:
        heads = self.attn_heads if self.attn_heads is not None else 1
        if self.attn_dim_head is not None:
            dim_head = self.attn_dim_head
        elif heads == 1:
            dim_head = channels
        else:
            dim_head = max(1, channels // heads)
        return SpatialSelfAttention(dim=channels, heads=heads, dim_head=dim_head, use_linear=False,
                                    use_efficient_attn=True)

    # ------------------------------------------------------------------------------------------------------
    def _stem(self, z: torch.Tensor, pre: Optional[nn.Conv2d] = None, z_scale: float = 1.0) -> torch.Tensor:
        """conv_in(pre(z * z_scale)) with the 1x1 conv `pre` (AutoencoderKL.post_quant_conv) composed into conv_in's
        weights.  Exact under zero padding: pre's bias rides on a constant-one extra input channel (a padded pixel
        contributes neither pre(z) nor pre's bias, exactly as in conv_in(pad(pre(z))))."""
        ops.require_cuda(z, "Decoder.forward")
        ci = self.conv_in.conv
        if pre is None and z_scale == 1.0:
            return self.conv_in(z)
        if z.shape[1] + 1 > 8 or ci.out_channels % 8:
            out_of_scope(f"Decoder stem with {z.shape[1]} latent channels / {ci.out_channels} features")

        def build():
            w_in = ci.weight.detach().float()                       # [O][M][3][3]
            if pre is None:
                wz = w_in * z_scale
                wb = torch.zeros_like(w_in[:, :1])
            else:
                w_pq = pre.weight.detach().float()[:, :, 0, 0]      # [M][I]
                wz = torch.einsum("omhw,mi->oihw", w_in, w_pq) * z_scale
                b_pq = pre.bias.detach().float() if pre.bias is not None else torch.zeros(w_pq.shape[0], device=w_in.device)
                wb = torch.einsum("omhw,m->ohw", w_in, b_pq)[:, None]
            return torch.cat([wz, wb], 1).contiguous()

        deps = [ci.weight] + ([pre.weight, pre.bias] if pre is not None else [])
        w = self._cache.get(f"stem:{z_scale}", deps, build)
        ones = torch.ones((z.shape[0], 1) + tuple(z.shape[2:]), dtype=torch.float32, device=z.device)
        return ops.conv_stem(z.float(), ones, w, f32(ci.bias))

    def _run(self, h: torch.Tensor) -> torch.Tensor:
        emb = None
        if self.emb_channels is not None:
            emb = torch.zeros(h.size(0), self.emb_channels, dtype=torch.float32, device=h.device)
        h = self.mid_block1(h, emb)
        h = self.mid_attn(h)
        h = self.mid_block2(h, emb)
        for stage in reversed(self.ups):
            has_up = hasattr(stage, "up")
            fuse_up = has_up and stage.up.can_fuse_into_producer()
            last = len(stage.blocks) - 1
            for i, blk in enumerate(stage.blocks):
                attn_follows = i < len(stage.attns)
                up_here = fuse_up and i == last and not attn_follows and isinstance(blk, ResBlockND) and blk._fast_ok()
                h = blk(h, emb, upsample_out=True) if up_here else blk(h, emb)
                if attn_follows:
                    h = stage.attns[i](h)
                    up_here = False
            if has_up:
                h = stage.up(h, upsampled=True) if up_here else stage.up(h)
        co = self.conv_out.conv
        if co.out_channels <= 4 and h.shape[1] % 8 == 0:
            tab = fused_group_norm_table(self.norm_out, [h], silu=True)
            if tab is not None:
                out = ops.conv_head(ops.to_nhwc_bf16(h), f32(co.weight), f32(co.bias), norm=tab)
                return torch.tanh(out) if self.tanh_out else out
        h = fused_group_norm(self.norm_out, [h], silu=True)
        out = self.conv_out(h).float().contiguous()
        return torch.tanh(out) if self.tanh_out else out

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        if self.spatial_dims != 2:
            out_of_scope("Decoder with spatial_dims != 2")
        return self._run(self._stem(z))
