"""Training step of the denoiser on the B200 kernels: differentiable ops, the training forward
graph, flat AdamW, bucketed gradient all-reduce, the flow-matching step and the epsilon-target diffusion step."""
from . import functions, graph
from .ddp import BucketedAllReduce
from .optim import FlatBuffers, FusedAdamW
from .step import DiffusionTrainer, FlowMatchingTrainer, diffusion_loss, flow_matching_loss

__all__ = ["functions", "graph", "BucketedAllReduce", "FlatBuffers", "FusedAdamW", "FlowMatchingTrainer",
           "DiffusionTrainer", "flow_matching_loss", "diffusion_loss"]
