"""Throughput of the other BASELINE.json configs through the same public API (context figures; the headline metric is
bench.py's).  One JSON line per config: samples/s of complete graph-replayed sampling runs, inputs resident on the GPU."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LDCT_UNET  # noqa: E402
from fmdm_b200.models.generators import DiffusionUNetFactory  # noqa: E402
from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler  # noqa: E402

DEV = torch.device("cuda")
MNIST = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
         "block_out_channels": [64, 128, 128], "down_block_types": ["DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
         "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D"]}
COMPVIS = {"in_channels": 1, "out_channels": 1, "num_res_blocks": 2, "channel_mult": [1, 1, 2, 2, 4, 4],
           "model_channels": 128, "attention_resolutions": [], "block_out_channels": [128, 128, 256, 256, 512, 512]}


def run(name, cfg, cond, B, hw, sched_name, steps, reps=3, flop_per_sample_fwd=None):
    torch.manual_seed(0)
    m = DiffusionUNetFactory().build(cfg, cond, 1).to(DEV).eval()
    for p in m.parameters():  # EfficientUNetND zero-initialises 74 tensors
        if float(p.abs().sum()) == 0:
            torch.nn.init.normal_(p, 0, 0.02)
    params = {"beta_start": 1e-4, "beta_end": 0.02} if sched_name != "flow_match_euler" else {}
    sch, _ = build_scheduler({"name": sched_name, "params": params}, {})
    x = torch.randn(B, 1, hw, hw, device=DEV)
    c = torch.rand(B, 1, hw, hw, device=DEV) if cond else None
    kw = dict(conditioning_mode=cond, conditioning_batch=c, init_sample=x)
    with torch.no_grad():
        for _ in range(2):
            sample_with_scheduler(m, sch, steps, tuple(x.shape), DEV, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            sample_with_scheduler(m, sch, steps, tuple(x.shape), DEV, **kw)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
    line = {"config": name, "batch": B, "hw": hw, "scheduler": sched_name, "steps": steps, "s_per_run": round(dt, 4),
            "samples_per_s": round(B / dt, 2)}
    if flop_per_sample_fwd:
        line["model_tflops"] = round(B / dt * flop_per_sample_fwd * steps / 1e12, 1)
    print(json.dumps(line), flush=True)


def run_attention_configs(B=16, hw=256, steps=50, reps=2):
    """the reference's `configs/LDCT/PixelAttention/*` (conditioning: "attention", latent context 4x32x32)
    through `sample_with_scheduler`: EfficientUNetND with linear self-/cross-attention at 16x16 and 8x8 and a softmax
    cross-attention middle block; flow-match Euler and the default `ddpm` ancestral sampler (50-step strided)."""
    import os

    cfg_path = "/root/reference/configs/LDCT/PixelAttention/LDCT_flow_matching_attention_compvis.json"
    if os.path.exists(cfg_path):
        cfg = json.load(open(cfg_path))["model"]["unet"]
    else:  # the same dictionary, for the GPU box where the reference tree does not exist
        cfg = {"unet_impl": "efficient_nd", "sample_size": 256, "in_channels": 1, "out_channels": 1,
               "layers_per_block": 2, "block_out_channels": [128, 128, 256, 256, 512, 512],
               "attention_resolutions": [16, 32], "cross_attention_resolutions": [16], "cross_attention_dim": 4,
               "cross_attention_in_middle": True, "emb_activation_before_proj": False}
    torch.manual_seed(0)
    m = DiffusionUNetFactory().build(cfg, "attention", 1).to(DEV).eval()
    for p in m.parameters():
        if float(p.abs().sum()) == 0:
            torch.nn.init.normal_(p, 0, 0.02)
    x = torch.randn(B, 1, hw, hw, device=DEV)
    ctx = torch.randn(B, 4, 32, 32, device=DEV)
    for sched_name in ("flow_match_euler", "ddpm"):
        params = {"beta_start": 1e-4, "beta_end": 0.02} if sched_name != "flow_match_euler" else {}
        sch, _ = build_scheduler({"name": sched_name, "params": params}, {})
        kw = dict(conditioning_mode="attention", conditioning_batch=ctx, latent_norm="standardize", init_sample=x)
        with torch.no_grad():
            for _ in range(2):
                out = sample_with_scheduler(m, sch, steps, tuple(x.shape), DEV, **kw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = sample_with_scheduler(m, sch, steps, tuple(x.shape), DEV, **kw)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
        assert torch.isfinite(out).all()
        print(json.dumps({"config": "PixelAttention compvis (conditioning: attention, context 4x32x32), LDCT 256x256",
                          "model": type(m).__name__, "batch": B, "scheduler": sched_name, "steps": steps,
                          "s_per_run": round(dt, 4), "samples_per_s": round(B / dt, 2)}), flush=True)


def run_latent_config(B=128, lat=64, steps=50, reps=2):
    """configs[2]: latent flow matching on AutoencoderKL f=8 latents (64x64x4, concat conditioning latents -> 8 input
    channels), 50 Euler steps, then KL decode to 512x512."""
    from fmdm_b200.models.vae import AutoencoderKL

    torch.manual_seed(0)
    cfg = dict(LDCT_UNET, in_channels=4, out_channels=4)
    unet = DiffusionUNetFactory().build(cfg, "concatenate", 4).to(DEV).eval()
    vae = AutoencoderKL(in_channels=1, out_channels=1, resolution=256, down_channels=(128, 256, 512, 512),
                        num_res_blocks=2, z_channels=4, embed_dim=4, attn_heads=4, attn_dim_head=64)
    for p in vae.parameters():
        if float(p.abs().sum()) == 0:
            torch.nn.init.normal_(p, 0, 0.02)
    vae = vae.to(DEV).eval()
    sch, _ = build_scheduler({"name": "flow_match_euler", "params": {}}, {})
    x = torch.randn(B, 4, lat, lat, device=DEV)
    c = torch.randn(B, 4, lat, lat, device=DEV)
    kw = dict(conditioning_mode="concatenate", conditioning_batch=c, init_sample=x)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps, out

    with torch.no_grad():
        t_unet, z = timed(lambda: sample_with_scheduler(unet, sch, steps, tuple(x.shape), DEV, **kw))
        t_dec, img = timed(lambda: vae.raw_output_to_image(vae.decode(z, denorm=True)))
    assert img.shape == (B, 1, 8 * lat, 8 * lat) and torch.isfinite(img).all()
    flop = B * (3.1091e10 * steps + 2.4918e12)
    print(json.dumps({"config": "configs[2] latent flow matching 64x64x4 (50 Euler) + AutoencoderKL decode to 512x512",
                      "batch": B, "s_unet_50_steps": round(t_unet, 4), "s_kl_decode": round(t_dec, 4),
                      "samples_per_s": round(B / (t_unet + t_dec), 2),
                      "model_tflops": round(flop / (t_unet + t_dec) / 1e12, 1)}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "latent":
        run_latent_config()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "attention":
        run_attention_configs()
        sys.exit(0)
    run("configs[0] MNIST 28x28 uncond flow-matching, 50 Euler", MNIST, None, 64, 28, "flow_match_euler", 50,
        flop_per_sample_fwd=2.3969e9)
    run("configs[3] LDCT 256x256 DDIM 50", LDCT_UNET, "concatenate", 16, 256, "ddim", 50, flop_per_sample_fwd=4.9657e11)
    run("configs[3] LDCT 256x256 DPM-Solver++ 20", LDCT_UNET, "concatenate", 16, 256, "dpm_multistep", 20,
        flop_per_sample_fwd=4.9657e11)
    run("configs[3] LDCT 512x512 DDIM 50", LDCT_UNET, "concatenate", 16, 512, "ddim", 50, reps=2,
        flop_per_sample_fwd=1.9944e12)
    run("configs[3] LDCT 512x512 DPM-Solver++ 20", LDCT_UNET, "concatenate", 16, 512, "dpm_multistep", 20, reps=2,
        flop_per_sample_fwd=1.9944e12)
    run("EfficientUNetND (compvis) LDCT 512x512 flow-matching, 50 Euler", COMPVIS, "concatenate", 16, 512,
        "flow_match_euler", 50, reps=2, flop_per_sample_fwd=1.9726e12)
