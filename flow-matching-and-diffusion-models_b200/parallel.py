"""Multi-GPU sampling: the batch is sharded across ranks, one process per GPU, no collective inside the loop.

Every sample's N-step trajectory is independent (GroupNorm and attention are per-sample), so rank r owns the
contiguous slice [r*B/G, (r+1)*B/G) of the global batch and replicated weights; the only exchange is ONE all-gather
of the final fp32 samples over NCCL/NVLink.  The reference samples single-process
(`src/pipelines/samplers/diffusion_like.py` has no `dist` usage); this module is the north star's sharding."""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .pipelines.utils import sample_with_scheduler


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the default process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, world, local_rank


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice of `total` items for `rank` (first `total % world` ranks get one extra)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_samples(local: torch.Tensor, total: int, rank: int, world: int) -> torch.Tensor:
    """All-gather the per-rank sample slices back into the global batch (ragged slices padded to the largest)."""
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    padded = local
    if local.shape[0] < longest:
        pad = torch.zeros((longest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    padded = padded.contiguous()
    out = torch.empty((world * longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded) if hasattr(dist, "all_gather_into_tensor") and local.is_cuda else \
        dist.all_gather(list(out.chunk(world, 0)), padded)
    pieces = [out[r * longest: r * longest + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(pieces, 0)


def sample_sharded(model, scheduler, num_inference_steps: int, global_noise: torch.Tensor,
                   global_cond: Optional[torch.Tensor], device, *, conditioning_mode: Optional[str] = "concatenate",
                   rank: Optional[int] = None, world: Optional[int] = None, gather: bool = True, **kw) -> torch.Tensor:
    """Sample the global batch described by `global_noise` (+ `global_cond`): each rank runs its slice through
    `sample_with_scheduler`, then the slices are all-gathered.  Results do not depend on the number of ranks."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    total = global_noise.shape[0]
    lo, hi = shard_bounds(total, rank, world)
    noise = global_noise[lo:hi]
    cond = None if global_cond is None else global_cond[lo:hi]
    if hi > lo:
        local = sample_with_scheduler(model, scheduler, num_inference_steps, tuple(noise.shape), device,
                                      conditioning_mode=conditioning_mode if cond is not None else None,
                                      conditioning_batch=cond, init_sample=noise, **kw)
    else:
        local = torch.empty((0,) + tuple(global_noise.shape[1:]), dtype=torch.float32, device=device)
    return gather_samples(local, total, rank, world) if gather else local
