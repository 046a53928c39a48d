"""Data-parallel gradient reduction of the training step on CPU (world_size 2, gloo): the bucketed, hook-driven
all-reduce over the flat gradient buffer must equal the sum of the per-rank gradients, whatever the bucket size, and
parameters that receive no gradient must still be reduced (finish() flushes incomplete buckets)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fmdm_b200.training.ddp import BucketedAllReduce
from fmdm_b200.training.optim import FlatBuffers


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(12, 33), torch.nn.SiLU(), torch.nn.Linear(33, 17), torch.nn.SiLU(),
                              torch.nn.Linear(17, 5))
    net.unused = torch.nn.Parameter(torch.ones(7))  # never touched by the loss
    return net


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(6, 12, generator=g), torch.randn(6, 5, generator=g)


def _local_grads(rank):
    net = _model()
    x, y = _data(rank)
    torch.nn.functional.mse_loss(net(x), y).backward()
    return [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in net.parameters()]


def _worker(rank, world, port, bucket_bytes, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _model()
    flat = FlatBuffers(net.parameters())
    red = BucketedAllReduce(flat, bucket_bytes=bucket_bytes)
    x, y = _data(rank)
    outs = []
    for _ in range(2):  # two cycles: the hook bookkeeping re-arms correctly
        flat.grad.zero_()
        red.arm()
        torch.nn.functional.mse_loss(net(x), y).backward()
        red.finish()
        outs.append([p.grad.clone() for p in net.parameters()])
    ret[rank] = (outs, len(red.buckets), [tuple(p.shape) for p in net.parameters()])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [64, 1024, 1 << 20])
def test_bucketed_allreduce_sums_rank_gradients(bucket_bytes):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), bucket_bytes, ret), nprocs=2, join=True)
    want = [a + b for a, b in zip(_local_grads(0), _local_grads(1))]
    for rank in (0, 1):
        outs, nb, _ = ret[rank]
        assert nb >= 1
        for cycle in outs:
            for got, ref in zip(cycle, want):
                assert torch.allclose(got, ref, rtol=1e-6, atol=1e-7)
    if bucket_bytes == 64:
        assert ret[0][1] > 3  # small buckets really split the buffer


def test_flat_buffers_reseat_parameters_and_gradients():
    net = _model()
    before = [p.detach().clone() for p in net.parameters()]
    flat = FlatBuffers(net.parameters())
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p.detach(), b)
        o, n = flat.slice_of(p)
        assert p.data_ptr() == flat.data.data_ptr() + 4 * o and o % 4 == 0
        assert p.grad.data_ptr() == flat.grad.data_ptr() + 4 * o
    x, y = _data(0)
    torch.nn.functional.mse_loss(net(x), y).backward()
    ref = _local_grads(0)
    for p, r in zip(net.parameters(), ref):
        assert torch.allclose(p.grad, r)
    # a foreign zero_grad(set_to_none=True) is repaired
    for p in net.parameters():
        p.grad = None
    flat.ensure_grad_views()
    assert all(p.grad is not None and p.grad.data_ptr() == flat.grad.data_ptr() + 4 * flat.slice_of(p)[0]
               for p in net.parameters())


def test_data_parallel_mean_equals_single_process_gradient_of_the_concatenated_batch():
    """"8-GPU grads == 1-GPU grads on the concatenated batch".  Every rank back-propagates the MEAN loss of
    its shard; the all-reduce sums the shard gradients and the optimiser folds in 1/world (`FusedAdamW.grad_scale`), which
    is the gradient of the mean loss over the concatenated batch when the shards have equal size."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), 256, ret), nprocs=2, join=True)
    reduced = ret[0][0][0]                                  # rank 0, first cycle: summed shard gradients
    net = _model()
    (x0, y0), (x1, y1) = _data(0), _data(1)
    torch.nn.functional.mse_loss(net(torch.cat([x0, x1])), torch.cat([y0, y1])).backward()
    for got, p in zip(reduced, net.parameters()):
        want = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(got * 0.5, want, rtol=1e-5, atol=1e-7)


def _staged_worker(rank, world, port, bucket_bytes, ret):
    """The CUDA-graph path's reduction schedule (`FlowMatchingTrainer._step`): backward stage 1 -> all-reduce of the
    ranges whose gradients are complete -> backward stage 2 (`BackwardCut.finish`) -> all-reduce of the rest."""
    from fmdm_b200.training.graph import BackwardCut

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _model()
    flat = FlatBuffers(net.parameters())
    red = BucketedAllReduce(flat, bucket_bytes=bucket_bytes)
    x, y = _data(rank)
    outs, ranges = [], None
    for cycle in range(2):
        flat.grad.zero_()
        cut = BackwardCut()
        h = net[1](net[0](x))                              # early part: stage 2 of the backward
        h2 = cut.cross(h)
        assert cut.cross(h) is h2                          # one leaf per crossing tensor
        loss = torch.nn.functional.mse_loss(net[4](net[3](net[2](h2))), y)
        if ranges is None:
            red.record(True)
        loss.backward()                                    # stage 1: stops at the leaf
        if ranges is None:
            first = red.ranges_of(red.record(False))
            ranges = (first, red.complement(first))
        assert net[0].weight.grad.abs().sum() == 0         # nothing has reached the early layers yet
        red.launch(ranges[0])
        cut.finish()
        red.launch(ranges[1])
        red.wait()
        outs.append([p.grad.clone() for p in net.parameters()])
    ret[rank] = (outs, ranges)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [64, 1 << 20])
def test_two_stage_backward_reduction_equals_plain_sum(bucket_bytes):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_staged_worker, args=(2, _free_port(), bucket_bytes, ret), nprocs=2, join=True)
    want = [a + b for a, b in zip(_local_grads(0), _local_grads(1))]
    for rank in (0, 1):
        outs, (first, rest) = ret[rank]
        for cyc in outs:
            for got, ref in zip(cyc, want):
                assert torch.allclose(got, ref, rtol=1e-6, atol=1e-7)
        # the two range sets tile the flat buffer exactly once; stage 1 holds the late layers (flat order is reversed)
        spans = sorted(first + rest)
        assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        late = 33 * 17 + 17 + 17 * 5 + 5
        assert sum(e - b for b, e in first) >= late and sum(e - b for b, e in rest) >= 12 * 33 + 33 + 7
        if bucket_bytes == 64:
            assert max(e - b for b, e in first) <= 16
