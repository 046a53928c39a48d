"""smoke(): one small invocation of the hot path on cuda:0 (UNetDiffusersND + flow-match Euler, graph-replayed),
checked against the oracle (fp32 restatement of the reference)."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
       "block_out_channels": [64, 128, 128], "down_block_types": ["DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
       "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D"]}


def run_smoke():
    from fmdm_b200 import ops
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler
    from oracle import denoiser as OD
    from oracle.sampling import make_scheduler, sample_loop

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = DiffusionUNetFactory().build(CFG, "concatenate", 1)
    sd = {k: v.to(dev) for k, v in model.state_dict().items()}
    model = model.to(dev).eval()
    g = torch.Generator().manual_seed(42)
    noise = torch.randn(2, 1, 32, 32, generator=g).to(dev)
    cond = torch.rand(2, 1, 32, 32, generator=g).to(dev)
    sched, _ = build_scheduler({"name": "flow_match_euler", "params": {}}, {})
    before = ops.launch_count()
    with torch.no_grad():
        out = sample_with_scheduler(model, sched, 8, tuple(noise.shape), dev, conditioning_mode="concatenate",
                                    conditioning_batch=cond, init_sample=noise).clamp(0, 1)
    torch.cuda.synchronize()
    launched = ops.launch_count() - before

    def oracle_model(inp, t):
        return OD.unet_diffusers_nd_forward(sd, CFG, inp[:, :1], t, conditioning="concatenate", channels=1,
                                            context=inp[:, 1:])

    class _GpuSched:  # oracle scheduler tables live on the CPU; step on CPU tensors
        def __init__(self, s):
            self.s = s
            self.set_timesteps = s.set_timesteps

        @property
        def timesteps(self):
            return self.s.timesteps

        def step(self, pred, t, x):
            r = self.s.step(pred.cpu(), t, x.cpu())
            r.prev_sample = r.prev_sample.to(dev)
            return r

    with torch.no_grad():
        ref = sample_loop(oracle_model, _GpuSched(make_scheduler("flowmatch", 1000)), 8, noise, cond).clamp(0, 1)
    mse = float(((out - ref) ** 2).mean())
    psnr = 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)
    assert launched > 0, "no fmdm_b200 kernels were launched"
    assert psnr >= 40.0, f"smoke: PSNR vs oracle {psnr:.1f} dB < 40 dB"
    # one denoiser forward on rows >= 65 px: the rolling-row tcgen05 conv with the fused GroupNorm operand transform
    g2 = torch.Generator().manual_seed(43)
    x = torch.randn(1, 1, 96, 160, generator=g2).to(dev)
    c = torch.rand(1, 1, 96, 160, generator=g2).to(dev)
    t = torch.full((1,), 321.5, device=dev)
    with torch.no_grad():
        mine = model(x, t, context=c)
        orc = OD.unet_diffusers_nd_forward(sd, CFG, x, t, conditioning="concatenate", channels=1, context=c)
    rel = float((mine - orc).norm() / orc.norm())
    assert rel < 1.2e-2, f"smoke: 96x160 forward rel-L2 {rel:.4f} vs oracle"
    # one training step: loss and gradients of the hand-written backward against the oracle's autograd
    from fmdm_b200.training import flow_matching_loss
    from oracle import training as OT

    model.train()
    clean = torch.rand(2, 1, 32, 32, generator=g2).to(dev)
    nz = torch.randn(2, 1, 32, 32, generator=g2).to(dev)
    tt = torch.rand(2, generator=g2).to(dev)
    loss = flow_matching_loss(model, clean, cond, noise=nz, t=tt)
    loss.backward()
    ref_loss, ref_grads = OT.loss_and_grads(sd, CFG, clean, cond, nz, tt)
    got = torch.cat([p.grad.float().reshape(-1) for _, p in model.named_parameters()])
    want = torch.cat([ref_grads[k].reshape(-1) for k, _ in model.named_parameters()])
    grel = float((got - want).norm() / want.norm())
    loss = loss.detach()
    assert abs(float(loss) - float(ref_loss)) <= 2e-2 * abs(float(ref_loss)), "smoke: training loss vs oracle"
    assert grel < 5e-2, f"smoke: training gradient rel-L2 {grel:.4f} vs oracle"
    model.eval()
    print(f"smoke ok: training step loss {float(loss):.4f} (oracle {float(ref_loss):.4f}), gradient rel-L2 {grel:.2e}")
    print(f"smoke ok: 8-step flow-matching sample, PSNR vs oracle {psnr:.1f} dB, {launched} kernel launches (eager part); "
          f"96x160 forward rel-L2 {rel:.2e}")


if __name__ == "__main__":
    run_smoke()
