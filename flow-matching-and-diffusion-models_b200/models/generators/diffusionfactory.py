"""DiffusionUNetFactory — `model.unet` JSON -> denoiser (`src/models/generators/diffusionfactory.py:25-130`).

Reads the reference's configs unchanged.  `unet_impl` in {diffusers_nd, diffusers_exact_nd, exact_nd, diffusers}
selects UNetDiffusersND, anything else EfficientUNetND; `conditioning == "concatenate"` widens the stem by the
conditioning channels."""
from __future__ import annotations

from typing import Any, Dict, Optional

from ..unet import EfficientUNetND, UNetDiffusersND

__all__ = ["DiffusionUNetFactory"]

_DIFFUSERS_IMPLS = frozenset({"diffusers_nd", "diffusers_exact_nd", "exact_nd", "diffusers"})


def _tuple_or(value, fallback):
    if value is None:
        return fallback
    return (value,) if isinstance(value, int) else tuple(value)


class DiffusionUNetFactory:
    DEFAULT_BLOCK_CHANNELS = (128, 128, 256, 256, 512, 512)

    def build(self, model_cfg: Dict[str, Any], conditioning: Optional[str] = None, channels: Optional[int] = None):
        cfg = dict(model_cfg or {})
        if str(cfg.get("unet_impl", "efficient_nd")).lower() in _DIFFUSERS_IMPLS:
            return self._build_diffusers_nd(cfg, conditioning, channels)
        return self._build_efficient_nd(cfg, conditioning, channels)

    # -- EfficientUNetND ------------------------------------------------------------------------------------
    def _build_efficient_nd(self, cfg, conditioning=None, channels=None):
        mode = (conditioning or "").lower()
        widths = _tuple_or(cfg.get("block_out_channels"), self.DEFAULT_BLOCK_CHANNELS)
        base = int(cfg.get("model_channels", widths[0] if widths else 128))
        n_in = cfg.get("in_channels", channels or 1)
        n_cond = cfg.get("conditioning_channels", channels or n_in)
        if mode == "concatenate":
            n_in = n_in + n_cond
        inferred_mult = tuple(max(1, int(w // (base or widths[0]))) for w in widths) if widths else ()
        mult = _tuple_or(cfg.get("channel_mult"), inferred_mult)
        attn_res = _tuple_or(cfg.get("attention_resolutions"), (1,))
        xattn_res = cfg.get("cross_attention_resolutions")
        xattn_mid = bool(cfg.get("cross_attention_in_middle", False))
        if xattn_res is None and mode == "attention":
            xattn_res = attn_res
            if "cross_attention_in_middle" not in cfg:
                xattn_mid = True
        return EfficientUNetND(
            spatial_dims=int(cfg.get("spatial_dims", 2)), in_channels=n_in, model_channels=base,
            out_channels=cfg.get("out_channels", channels or 1),
            num_res_blocks=int(cfg.get("num_res_blocks", cfg.get("layers_per_block", 2))),
            attention_resolutions=attn_res, cross_attention_resolutions=xattn_res,
            cross_attention_dim=int(cfg.get("cross_attention_dim", n_cond)), cross_attention_in_middle=xattn_mid,
            dropout=float(cfg.get("dropout", 0.0)), channel_mult=mult or (1, 2, 3, 4),
            conv_resample=bool(cfg.get("conv_resample", True)), dim_head=int(cfg.get("dim_head", 64)),
            num_heads=int(cfg.get("num_heads", 4)), use_linear_attn=bool(cfg.get("use_linear_attn", True)),
            use_scale_shift_norm=bool(cfg.get("use_scale_shift_norm", True)),
            emb_activation_before_proj=bool(cfg.get("emb_activation_before_proj", False)),
            pool_factor=int(cfg.get("pool_factor", 1)))

    # -- UNetDiffusersND ------------------------------------------------------------------------------------
    def _build_diffusers_nd(self, cfg, conditioning=None, channels=None):
        mode = (conditioning or "").lower()
        n_in = int(cfg.get("in_channels", channels or 1))
        n_cond = int(cfg.get("conditioning_channels", channels or n_in))
        if mode == "concatenate" and not bool(cfg.get("in_channels_already_conditioned", False)):
            n_in += n_cond
        if mode == "attention":
            down = ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",)
            up = ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3
            mid = "UNetMidBlock2DCrossAttn"
        else:
            down = ("DownBlock2D",) + ("AttnDownBlock2D",) * 3
            up = ("AttnUpBlock2D",) * 3 + ("UpBlock2D",)
            mid = "UNetMidBlock2D"
        return UNetDiffusersND(
            spatial_dims=int(cfg.get("spatial_dims", 2)), sample_size=cfg.get("sample_size"), in_channels=n_in,
            out_channels=int(cfg.get("out_channels", channels or 1)),
            center_input_sample=bool(cfg.get("center_input_sample", False)),
            time_embedding_type=str(cfg.get("time_embedding_type", "positional")),
            freq_shift=int(cfg.get("freq_shift", 0)), flip_sin_to_cos=bool(cfg.get("flip_sin_to_cos", True)),
            down_block_types=cfg.get("down_block_types", down), mid_block_type=cfg.get("mid_block_type", mid),
            up_block_types=cfg.get("up_block_types", up),
            block_out_channels=_tuple_or(cfg.get("block_out_channels"), (224, 448, 672, 896)),
            layers_per_block=int(cfg.get("layers_per_block", 2)),
            downsample_padding=int(cfg.get("downsample_padding", 1)), dropout=float(cfg.get("dropout", 0.0)),
            attention_head_dim=int(cfg.get("attention_head_dim", 8)),
            norm_num_groups=int(cfg.get("norm_num_groups", 32)), norm_eps=float(cfg.get("norm_eps", 1e-5)),
            resnet_time_scale_shift=str(cfg.get("resnet_time_scale_shift", "default")),
            add_attention=bool(cfg.get("add_attention", True)),
            cross_attention_dim=int(cfg.get("cross_attention_dim", n_cond)) if mode == "attention" else None)
