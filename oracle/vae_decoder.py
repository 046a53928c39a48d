"""ORACLE (test infrastructure, never imported by the product path).

Functional fp32 restatement of `AutoencoderKL.decode` -> `Decoder.forward` -> `raw_output_to_image`, driven by a
reference-format `state_dict` (keys `post_quant_conv.conv.*`, `decoder.*`) and the `model` block of the VAE JSON config.

Pinned: tests/test_oracle_vae.py checks it against the reference's own `AutoencoderKL` imported from
/root/reference/src when that tree is present, and against `tests/golden/vae_decode_*.pt` (generated from the reference
module by `oracle/make_golden_vae.py`) everywhere else.

Reference lines followed:
  AutoencoderKL.decode        src/models/vae/kl.py:126-130   (LATENT_SCALE = 0.18215, :19)
  Decoder.__init__ / forward  src/nn/modules/vae/decoder.py:24-160
  ResBlockND.forward          src/nn/blocks/residual.py:84-120   (emb = None)
  SpatialSelfAttention        src/nn/blocks/attention.py:102-117
  UpsampleND                  src/nn/ops/upsampling.py:8-30
  raw_output_to_image         src/models/autoencoder/base.py:21-28
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from .denoiser import _conv, resblock, spatial_self_attention

SD = Dict[str, torch.Tensor]
LATENT_SCALE = 0.18215


def decoder_channels(cfg: dict):
    if cfg.get("down_channels") is not None:
        return tuple(cfg["down_channels"])
    return tuple(int(cfg.get("base_ch", 128)) * m for m in cfg.get("ch_mult", (1, 2, 4, 4)))


def decoder_forward(sd: SD, cfg: dict, z: torch.Tensor, prefix: str = "decoder.") -> torch.Tensor:
    channels = decoder_channels(cfg)
    heads = cfg.get("attn_heads") if cfg.get("attn_heads") is not None else 1
    use_attn = bool(cfg.get("use_attention", True))
    nres = int(cfg.get("num_res_blocks", 2))
    attn_res = tuple(cfg.get("attn_resolutions", ()))
    h = _conv(sd, prefix + "conv_in.conv", z)
    h = resblock(sd, prefix + "mid_block1.", h, None)
    if use_attn:
        h = spatial_self_attention(sd, prefix + "mid_attn.", h, heads)
    h = resblock(sd, prefix + "mid_block2.", h, None)
    res = int(cfg.get("resolution", 256)) // (2 ** (len(channels) - 1))
    # stages are stored with ups.insert(0, ...): the first executed stage has the highest index
    for idx in range(len(channels)):
        stage = len(channels) - 1 - idx
        p = f"{prefix}ups.{stage}."
        na = 0
        for i in range(nres + 1):
            h = resblock(sd, f"{p}blocks.{i}.", h, None)
            if use_attn and res in attn_res:
                h = spatial_self_attention(sd, f"{p}attns.{na}.", h, heads)
                na += 1
        if idx != len(channels) - 1:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = _conv(sd, p + "up.conv.conv", h)
            res *= 2
    c = h.shape[1]
    groups = cfg.get("norm_groups") if cfg.get("norm_groups") is not None else max(1, math.gcd(c, 32))
    h = F.silu(F.group_norm(h, groups, sd[prefix + "norm_out.weight"], sd[prefix + "norm_out.bias"], 1e-5))
    h = _conv(sd, prefix + "conv_out.conv", h)
    return torch.tanh(h) if cfg.get("tanh_out", False) else h


def kl_decode(sd: SD, cfg: dict, z: torch.Tensor, denorm: bool = False) -> torch.Tensor:
    if denorm:
        z = z / LATENT_SCALE
    z = F.conv2d(z, sd["post_quant_conv.conv.weight"], sd["post_quant_conv.conv.bias"])
    return decoder_forward(sd, cfg, z)


def raw_output_to_image(x: torch.Tensor, recon_type: str = "l1") -> torch.Tensor:
    if str(recon_type).lower() in ("bce", "focal", "bce_focal"):
        return torch.sigmoid(x)
    return (x.clamp(-1.0, 1.0) + 1.0) * 0.5
