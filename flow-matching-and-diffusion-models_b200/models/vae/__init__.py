from .kl import LATENT_SCALE, AutoencoderKL

__all__ = ["AutoencoderKL", "LATENT_SCALE"]
