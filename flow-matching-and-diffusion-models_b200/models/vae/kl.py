"""AutoencoderKL, decode side — drop-in for `src/models/vae/kl.py:20-130` + `src/models/autoencoder/base.py:14-28` on
the path BASELINE config 3 uses: latents from the flow-matching sampler -> `decode` -> `raw_output_to_image`.

Same constructor keywords; `decoder.*` and `post_quant_conv.conv.*` state_dict keys are the reference's, so a reference
checkpoint loads with `load_state_dict` (its `encoder.*` / `quant_conv.*` entries are ignored: encoding is training-side
and out of scope here, `encode` raises)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from ..._runtime import out_of_scope
from ...nn.modules.vae import Decoder
from ...nn.ops.convolution import ConvND

LATENT_SCALE: float = 0.18215


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels: int = 3, out_channels: int = 3, resolution: int = 256, base_ch: int = 128,
                 ch_mult: Tuple[int, ...] = (1, 2, 4, 4), down_channels: Optional[Tuple[int, ...]] = None,
                 num_res_blocks: int = 2, attn_resolutions: Tuple[int, ...] = (), z_channels: int = 4,
                 embed_dim: int = 4, dropout: float = 0.0, use_attention: bool = True, attn_heads: int = 4,
                 attn_dim_head: int = 64, spatial_dims: int = 2, emb_channels: Optional[int] = None,
                 use_scale_shift_norm: bool = False, norm_groups: Optional[int] = None,
                 codebook_size: Optional[int] = None, num_embeddings: Optional[int] = None,
                 ckpt_path: Optional[str] = None, double_z: bool = True, block_factory=None) -> None:
        super().__init__()
        self.spatial_dims = spatial_dims
        self.decoder = Decoder(out_ch=out_channels, base_ch=base_ch, ch_mult=ch_mult, down_channels=down_channels,
                               num_res_blocks=num_res_blocks, attn_resolutions=attn_resolutions, resolution=resolution,
                               z_channels=z_channels, dropout=dropout, use_attention=use_attention,
                               attn_heads=attn_heads, attn_dim_head=attn_dim_head, tanh_out=False,
                               spatial_dims=spatial_dims, emb_channels=emb_channels,
                               use_scale_shift_norm=use_scale_shift_norm, norm_groups=norm_groups,
                               block_factory=block_factory)
        self.post_quant_conv = ConvND(spatial_dims, embed_dim, z_channels, 1, padding=0)
        self.embed_dim, self.num_embeddings, self.codebook_size = embed_dim, num_embeddings, codebook_size
        if ckpt_path:
            state = torch.load(ckpt_path, map_location="cpu", weights_only=True)
            self.load_state_dict(state["model"] if isinstance(state, dict) and "model" in state else state)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts a full reference AutoencoderKL checkpoint: encoder-side tensors are dropped."""
        kept = {k: v for k, v in state_dict.items() if not (k.startswith("encoder.") or k.startswith("quant_conv."))}
        return super().load_state_dict(kept, strict=strict, assign=assign)

    # range helpers of BaseAutoencoder (base.py:18-28)
    def image_to_model_range(self, x: torch.Tensor) -> torch.Tensor:
        return x * 2.0 - 1.0

    def model_to_image_range(self, x: torch.Tensor) -> torch.Tensor:
        return (x.clamp(-1.0, 1.0) + 1.0) * 0.5

    def raw_output_to_image(self, x: torch.Tensor, recon_type: str = "l1") -> torch.Tensor:
        if str(recon_type).lower() in {"bce", "focal", "bce_focal"}:
            return torch.sigmoid(x)
        return self.model_to_image_range(x)

    def encode(self, x: torch.Tensor, normalize: bool = False):
        out_of_scope("AutoencoderKL.encode (training-side encoder)")

    def decode(self, z: torch.Tensor, denorm: bool = False) -> torch.Tensor:
        """z: (B, embed_dim, h, w) fp32 latents -> raw (B, out_channels, 8h, 8w) fp32 (kl.py:126-130)."""
        dec = self.decoder
        h = dec._stem(z, pre=self.post_quant_conv.conv, z_scale=(1.0 / LATENT_SCALE) if denorm else 1.0)
        return dec._run(h)

    def forward(self, x: torch.Tensor, sample_posterior: bool = True):
        out_of_scope("AutoencoderKL.forward (encode + decode)")
