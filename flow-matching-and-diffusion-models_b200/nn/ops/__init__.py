from .convolution import ConvND, ConvTransposeND
from .normalization import RMSNormND, fused_group_norm, fused_group_norm_table, make_group_norm
from .pooling import AvgPoolND, MaxPoolND, PoolND, UnPoolND
from .time_embedding import timestep_embedding
from .upsampling import DownsampleND, UpsampleND

__all__ = ["ConvND", "ConvTransposeND", "PoolND", "AvgPoolND", "MaxPoolND", "UnPoolND", "RMSNormND",
           "make_group_norm", "fused_group_norm", "timestep_embedding", "UpsampleND", "DownsampleND"]
