from .diffusion_utils import (build_diffusion_model, decode_diffusion_batch, encode_diffusion_batch,
                              warn_attention_conditioning_shape)

__all__ = ["build_diffusion_model", "decode_diffusion_batch", "encode_diffusion_batch",
           "warn_attention_conditioning_shape"]
