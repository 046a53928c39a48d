"""N>1 host logic on CPU: world_size-2 gloo run of the sharded sampler (dummy denoiser + oracle scheduler, since
the product kernels have no CPU path) must reproduce the single-process result sample for sample."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fmdm_b200.parallel import gather_samples, sample_sharded, shard_bounds
from oracle.sampling import make_scheduler


def test_shard_bounds():
    assert [shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_bounds(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(128, 7, 8) == (112, 128)
    with pytest.raises(ValueError):
        shard_bounds(4, 4, 4)


class _Denoiser(torch.nn.Module):
    """cheap deterministic stand-in: v = a * x + b * cond (per-sample independent, like the real denoiser)"""

    def forward(self, inp, t):
        x, c = inp[:, :1], inp[:, 1:]
        return 0.3 * x - 0.5 * c + (t.view(-1, 1, 1, 1) / 1000.0) * 0.1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(total, 1, 8, 8, generator=g)
    cond = torch.rand(total, 1, 8, 8, generator=g)
    out = sample_sharded(_Denoiser(), make_scheduler("flowmatch"), 10, noise, cond, torch.device("cpu"),
                         use_cuda_graph=False)
    lo, hi = shard_bounds(total, rank, world)
    ret[rank] = (out, (lo, hi))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [6, 5])
def test_sharded_sampling_matches_single_process(total):
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(total, 1, 8, 8, generator=g)
    cond = torch.rand(total, 1, 8, 8, generator=g)
    single = sample_sharded(_Denoiser(), make_scheduler("flowmatch"), 10, noise, cond, torch.device("cpu"),
                            rank=0, world=1, use_cuda_graph=False)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), total, ret), nprocs=2, join=True)
    for rank in range(2):
        out, _ = ret[rank]
        assert out.shape == single.shape
        assert torch.equal(out, single)


def test_gather_single_rank_is_identity():
    x = torch.arange(6.0).view(3, 2)
    assert gather_samples(x, 3, 0, 1) is x
