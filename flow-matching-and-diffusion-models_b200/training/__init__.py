"""Training step of the denoiser on the B200 kernels: differentiable ops, the training forward
graph, flat AdamW, bucketed gradient all-reduce and the flow-matching step."""
from . import functions, graph
from .ddp import BucketedAllReduce
from .optim import FlatBuffers, FusedAdamW
from .step import FlowMatchingTrainer, flow_matching_loss

__all__ = ["functions", "graph", "BucketedAllReduce", "FlatBuffers", "FusedAdamW", "FlowMatchingTrainer",
           "flow_matching_loss"]
