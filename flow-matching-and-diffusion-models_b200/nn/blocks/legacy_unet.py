"""Diffusers-compatible down / up / mid blocks (`src/nn/blocks/legacy_unet.py`), same constructor arguments and
children (`resnets`, `attentions`, `downsamplers`, `upsamplers`).  The up path hands each ResBlockND the pair
(hidden, skip) instead of a `torch.cat` copy (`legacy_unet.py:150`): GroupNorm and the skip conv read both tensors
as K/channel segments."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..ops.upsampling import DownsampleND, UpsampleND
from .attention import DiffusersAttentionND
from .residual import ResBlockND


def _resnet(spatial_dims, cin, cout, temb_channels, eps, groups, dropout, time_scale_shift):
    return ResBlockND(spatial_dims=spatial_dims, channels=cin, emb_channels=temb_channels, out_channels=cout,
                      dropout=dropout, use_conv=False, use_scale_shift_norm=(time_scale_shift == "scale_shift"),
                      norm_type="gn", norm_groups=groups, norm_eps=eps, zero_init_last_conv=False,
                      emb_activation_before_proj=True, add_embedding_to_hidden=True)


def _attn(channels, attention_head_dim, cross_attention_dim, eps, groups):
    heads = max(1, channels // max(attention_head_dim, 1))
    return DiffusersAttentionND(channels, heads=heads, context_dim=cross_attention_dim, eps=eps,
                                norm_num_groups=groups)


class DownBlock2DCompat(nn.Module):
    def __init__(self, spatial_dims: int, num_layers: int, in_channels: int, out_channels: int, temb_channels: int,
                 add_downsample: bool, eps: float, groups: int, dropout: float, time_scale_shift: str,
                 with_attention: bool = False, attention_head_dim: int = 8, cross_attention_dim: int | None = None):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if with_attention else None
        for layer in range(num_layers):
            cin = in_channels if layer == 0 else out_channels
            self.resnets.append(_resnet(spatial_dims, cin, out_channels, temb_channels, eps, groups, dropout,
                                        time_scale_shift))
            if with_attention:
                self.attentions.append(_attn(out_channels, attention_head_dim, cross_attention_dim, eps, groups))
        self.downsamplers = nn.ModuleList([DownsampleND(spatial_dims, out_channels, use_conv=True)]) \
            if add_downsample else None

    def forward(self, hidden_states: torch.Tensor, temb: torch.Tensor, context: torch.Tensor | None = None):
        outputs = []
        for idx, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb)
            if self.attentions is not None:
                hidden_states = self.attentions[idx](hidden_states, context=context)
            outputs.append(hidden_states)
        if self.downsamplers is not None:
            for down in self.downsamplers:
                hidden_states = down(hidden_states)
            outputs.append(hidden_states)
        return hidden_states, tuple(outputs)


class UpBlock2DCompat(nn.Module):
    def __init__(self, spatial_dims: int, num_layers: int, in_channels: int, out_channels: int,
                 prev_output_channel: int, temb_channels: int, add_upsample: bool, eps: float, groups: int,
                 dropout: float, time_scale_shift: str, with_attention: bool = False, attention_head_dim: int = 8,
                 cross_attention_dim: int | None = None):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList() if with_attention else None
        for layer in range(num_layers):
            skip_ch = in_channels if layer == num_layers - 1 else out_channels
            hid_ch = prev_output_channel if layer == 0 else out_channels
            self.resnets.append(_resnet(spatial_dims, hid_ch + skip_ch, out_channels, temb_channels, eps, groups,
                                        dropout, time_scale_shift))
            if with_attention:
                self.attentions.append(_attn(out_channels, attention_head_dim, cross_attention_dim, eps, groups))
        self.upsamplers = nn.ModuleList([UpsampleND(spatial_dims, out_channels, use_conv=True)]) \
            if add_upsample else None

    def forward(self, hidden_states: torch.Tensor, res_hidden_states_tuple, temb: torch.Tensor,
                context: torch.Tensor | None = None):
        pending = list(res_hidden_states_tuple)
        # the nearest-2x of the upsampler is folded into the store of whichever conv produces its input (the last
        # ResBlock's conv2 or the last attention's output projection)
        fuse_up = (self.upsamplers is not None and len(self.upsamplers) == 1 and hidden_states.is_cuda
                   and self.upsamplers[0].can_fuse_into_producer() and context is None)
        last = len(self.resnets) - 1
        for idx, resnet in enumerate(self.resnets):
            skip = pending.pop()
            up_here = fuse_up and idx == last and self.attentions is None
            hidden_states = resnet((hidden_states, skip), temb, upsample_out=up_here) if up_here \
                else resnet((hidden_states, skip), temb)  # virtual concat, hidden first
            if self.attentions is not None:
                up_here = fuse_up and idx == last
                hidden_states = self.attentions[idx](hidden_states, context=context, upsample_out=True) if up_here \
                    else self.attentions[idx](hidden_states, context=context)
        if self.upsamplers is not None:
            for up in self.upsamplers:
                hidden_states = up(hidden_states, upsampled=True) if fuse_up else up(hidden_states)
        return hidden_states


class UNetMidBlock2DCompat(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, temb_channels: int, eps: float, groups: int,
                 dropout: float, time_scale_shift: str, add_attention: bool = True, attention_head_dim: int = 8,
                 cross_attention_dim: int | None = None):
        super().__init__()
        self.resnets = nn.ModuleList([
            _resnet(spatial_dims, in_channels, in_channels, temb_channels, eps, groups, dropout, time_scale_shift)
            for _ in range(2)
        ])
        self.attentions = nn.ModuleList([_attn(in_channels, attention_head_dim, cross_attention_dim, eps, groups)]) \
            if add_attention else None

    def forward(self, hidden_states: torch.Tensor, temb: torch.Tensor, context: torch.Tensor | None = None):
        hidden_states = self.resnets[0](hidden_states, temb)
        if self.attentions is not None:
            hidden_states = self.attentions[0](hidden_states, context=context)
        return self.resnets[1](hidden_states, temb)
