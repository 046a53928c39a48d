from .attention import (ContextBlock, DiffusersAttentionND, LinearQKVAttention, QKVAttention, SpatialCrossAttention,
                        SpatialSelfAttention)
from .common import zero_module
from .legacy_unet import DownBlock2DCompat, UNetMidBlock2DCompat, UpBlock2DCompat
from .residual import (ResBlockND, build_resblock_gn_silu, build_resblock_gn_swish, build_resblock_rmsnorm_silu,
                       build_resblock_rmsnorm_swish)
from .timestep import TimestepBlock

__all__ = ["QKVAttention", "LinearQKVAttention", "ContextBlock", "SpatialSelfAttention", "SpatialCrossAttention",
           "DiffusersAttentionND", "zero_module", "DownBlock2DCompat", "UpBlock2DCompat", "UNetMidBlock2DCompat",
           "ResBlockND", "build_resblock_gn_silu", "build_resblock_gn_swish", "build_resblock_rmsnorm_silu",
           "build_resblock_rmsnorm_swish", "TimestepBlock"]
