"""BASELINE config 5: LDCT 256x256 flow-matching training step (fwd + bwd + AdamW, gradient all-reduce when launched
under torchrun), batch 16 per GPU, synthetic data.  Prints one JSON line (samples/s over all ranks)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LDCT_UNET, ClockSampler  # noqa: E402


COMPVIS = {"in_channels": 1, "out_channels": 1, "num_res_blocks": 2, "channel_mult": [1, 1, 2, 2, 4, 4],
           "model_channels": 128, "attention_resolutions": [], "block_out_channels": [128, 128, 256, 256, 512, 512]}


def model_cfg():
    """MODEL=compvis selects EfficientUNetND (configs/LDCT/LDCT_flow_matching_compvis.json); default UNetDiffusersND."""
    return COMPVIS if os.environ.get("MODEL", "").lower() in ("compvis", "efficient_nd") else LDCT_UNET


def eager_reference(B, hw, steps, warmup, dev):
    """The same step on the same GPU with stock PyTorch kernels (cuDNN / cuBLAS / SDPA under bf16 autocast, fused
    torch.optim.AdamW), through the oracle's functional restatement of the reference denoiser: context for the
    number above, not a product path."""
    import torch.nn.functional as TF

    from fmdm_b200.models.generators import DiffusionUNetFactory
    from oracle import denoiser as OD

    torch.manual_seed(0)
    model = DiffusionUNetFactory().build(model_cfg(), "concatenate", 1)
    params = {k: torch.nn.Parameter(v.detach().clone().to(dev).to(memory_format=torch.channels_last)
                                    if v.dim() == 4 else v.detach().clone().to(dev))
              for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(params.values(), lr=1e-4, fused=True)
    g = torch.Generator(device=dev).manual_seed(1)
    clean = torch.rand(B, 1, hw, hw, device=dev, generator=g)
    ldct = torch.rand(B, 1, hw, hw, device=dev, generator=g)

    def step():
        opt.zero_grad(set_to_none=True)
        noise = torch.randn_like(clean)
        t = torch.rand(B, device=dev)
        timesteps = (t * 999).long()
        x_t = (1.0 - t[:, None, None, None]) * clean + t[:, None, None, None] * noise
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = OD.denoiser_forward(params, model_cfg(), x_t, timesteps, conditioning="concatenate", channels=1,
                                       context=ldct)
            loss = TF.mse_loss(pred.float(), noise - clean)
        loss.backward()
        opt.step()
        return loss.detach()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"ms_per_step": round(ms, 2), "samples_per_s": round(B / (ms / 1e3), 2),
            "what": "torch eager bf16 autocast (cuDNN/cuBLAS/SDPA) + fused torch AdamW, same GPU, same step"}


def cpu_reference(hw):
    """The reference's training-step arithmetic on the box's host cores (oracle/training.py: the functional restatement
    of the reference denoiser under torch autograd + the AdamW update), B=1, one step timed after one warm-up: the
    reported CPU baseline of SURVEY.md 8d for config 5 (not a target)."""
    import time

    from fmdm_b200.models.generators import DiffusionUNetFactory
    from oracle import training as OT

    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in DiffusionUNetFactory().build(model_cfg(), "concatenate", 1).state_dict().items()}
    g = torch.Generator().manual_seed(2)
    batch = (torch.rand(1, 1, hw, hw, generator=g), torch.rand(1, 1, hw, hw, generator=g),
             torch.randn(1, 1, hw, hw, generator=g), torch.rand(1, generator=g))
    OT.train_steps(sd, model_cfg(), [batch], lr=1e-4, weight_decay=0.0)
    t0 = time.perf_counter()
    OT.train_steps(sd, model_cfg(), [batch], lr=1e-4, weight_decay=0.0)
    dt = time.perf_counter() - t0
    return {"value": round(1.0 / dt, 4), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"B=1, one fwd+bwd+AdamW step at {hw}x{hw}, fp32 torch CPU oracle port, all host threads"}


def main():
    import torch.distributed as dist

    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.training import FlowMatchingTrainer

    steps = int(os.environ.get("STEPS", 5))
    warmup = int(os.environ.get("WARMUP", 3))
    B = int(os.environ.get("BATCH", 16))
    hw = int(os.environ.get("HW", 256))
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    model = DiffusionUNetFactory().build(model_cfg(), "concatenate", 1).to(dev).train()
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    model_name = type(model).__name__
    kw = {}
    if os.environ.get("BACKWARD_CUT"):       # diagnostics: "none" = reduce after the replay (round-1 schedule), or an int
        kw["backward_cut"] = None if os.environ["BACKWARD_CUT"] == "none" else int(os.environ["BACKWARD_CUT"])
    if os.environ.get("NO_SIDE_STREAMS"):
        kw["side_streams"] = False
    tr = FlowMatchingTrainer(model, lr=1e-4, cuda_graph=not os.environ.get("NO_GRAPH"), **kw)
    if os.environ.get("SKIP_ALLREDUCE"):     # diagnostics: the multi-rank step without its collective (rank spread only)
        tr.reducer.launch = lambda ranges: None
        tr.reducer.reduce_all = lambda: None
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    clean = torch.rand(B, 1, hw, hw, device=dev, generator=g)
    ldct = torch.rand(B, 1, hw, hw, device=dev, generator=g)
    losses = []
    for _ in range(warmup):
        losses.append(tr.step(clean, ldct))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        losses.append(tr.step(clean, ldct))
    enqueue_ms = (time.perf_counter() - t0) * 1e3 / steps  # host time to issue one step (no sync inside)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler is not None else None
    # end to end: every step takes its batch from pinned host memory and hands the loss back to the host
    h_clean, h_ldct = clean.cpu().pin_memory(), ldct.cpu().pin_memory()
    d_clean, d_ldct = torch.empty_like(clean), torch.empty_like(ldct)
    h_loss = torch.empty((), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(steps):
        d_clean.copy_(h_clean, non_blocking=True)
        d_ldct.copy_(h_ldct, non_blocking=True)
        h_loss.copy_(tr.step(d_clean, d_ldct), non_blocking=True)
    f1.record()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([f0.elapsed_time(f1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    eager = None
    if rank == 0 and world == 1 and os.environ.get("EAGER_BASELINE"):
        del tr, model
        torch.cuda.empty_cache()
        eager = eager_reference(B, hw, steps, warmup, dev)
    cpu = cpu_reference(hw) if rank == 0 and world == 1 and os.environ.get("CPU_BASELINE") else None
    if rank == 0:
        ls = [float(x) for x in losses]
        print(json.dumps({"metric": "training_samples_per_s", "value": round(B * world / (ms.item() / 1e3), 2),
                          "unit": "samples/s", "n_gpus": world, "ms_per_step": round(ms.item(), 2), "steps": steps,
                          "warmup": warmup, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": f"LDCT {hw}x{hw} flow-matching training step, batch {B}/GPU",
                                     "model": model_name,
                                     "optimizer": "AdamW (flat, fused)", "cuda_graph": not os.environ.get("NO_GRAPH"), "loss_first": ls[0], "loss_last": ls[-1],
                                     "grad_reduce": getattr(tr, "reduce_mode", None),
                                     "diagnostic_env": {k: os.environ[k] for k in ("BACKWARD_CUT", "NO_SIDE_STREAMS",
                                                                                  "SKIP_ALLREDUCE", "NCCL_MAX_CTAS")
                                                        if k in os.environ}},
                          "host_enqueue_ms_per_step": round(enqueue_ms, 2),
                          "e2e": {"value": round(B * world / (e2e_ms.item() / 1e3), 2), "unit": "samples/s",
                                  "h2d_bytes_per_step": int(clean.numel() + ldct.numel()) * 4, "d2h_bytes_per_step": 4},
                          "clocks": clocks, "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1),
                          "gpu_eager_reference": eager, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
