from .decoder import Decoder

__all__ = ["Decoder"]
