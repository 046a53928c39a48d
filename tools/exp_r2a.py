"""Round-2 GPU experiments (diagnostics, not product code): numbers the new parity tests are calibrated on."""
import math
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import denoiser as OD  # noqa: E402
from oracle import training as OT  # noqa: E402
from oracle.sampling import make_scheduler  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"

MNIST_UNET = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
              "block_out_channels": [64, 128, 128],
              "down_block_types": ["DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
              "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D"]}
LDCT = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 2,
        "block_out_channels": [128, 128, 256, 256, 512, 512],
        "down_block_types": ["DownBlock2D"] * 4 + ["AttnDownBlock2D", "DownBlock2D"],
        "up_block_types": ["UpBlock2D", "AttnUpBlock2D"] + ["UpBlock2D"] * 4}
COMPVIS_LDCT = {"in_channels": 1, "out_channels": 1, "num_res_blocks": 2, "channel_mult": [1, 1, 2, 2, 4, 4],
                "model_channels": 128, "attention_resolutions": [], "block_out_channels": [128, 128, 256, 256, 512, 512]}
CA_DIFFUSERS = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 1,
                "block_out_channels": [64, 128], "cross_attention_dim": 4,
                "down_block_types": ["DownBlock2D", "CrossAttnDownBlock2D"], "mid_block_type": "UNetMidBlock2DCrossAttn",
                "up_block_types": ["CrossAttnUpBlock2D", "UpBlock2D"]}
CA_EFFICIENT = {"unet_impl": "efficient_nd", "in_channels": 1, "out_channels": 1, "num_res_blocks": 1,
                "channel_mult": [1, 2], "model_channels": 64, "block_out_channels": [64, 128],
                "attention_resolutions": [2], "cross_attention_resolutions": [2], "cross_attention_in_middle": True,
                "cross_attention_dim": 4, "use_linear_attn": False}
CA_EFFICIENT_LINEAR = {"unet_impl": "efficient_nd", "in_channels": 1, "out_channels": 1, "num_res_blocks": 1,
                       "channel_mult": [1, 2], "model_channels": 64, "block_out_channels": [64, 128],
                       "attention_resolutions": [1, 2], "cross_attention_resolutions": [2],
                       "cross_attention_in_middle": True, "cross_attention_dim": 4}


def rel_l2(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def psnr(a, b, peak=1.0):
    mse = float(((a.float() - b.float()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * math.log10(peak * peak / mse)


def build(cfg, conditioning, seed=1):
    from fmdm_b200.models.generators import DiffusionUNetFactory

    model = DiffusionUNetFactory().build(cfg, conditioning, 1)
    sd = OD.reinit_state_dict(model.state_dict(), seed)
    model.load_state_dict(sd)
    return model.to(DEV).eval(), {k: v.to(DEV) for k, v in sd.items()}


def section(fn):
    print(f"\n===== {fn.__name__} =====", flush=True)
    t0 = time.time()
    try:
        fn()
    except Exception:
        traceback.print_exc()
    torch.cuda.synchronize()
    print(f"[{fn.__name__}: {time.time() - t0:.1f} s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB]", flush=True)
    torch.cuda.empty_cache()


def exp_split_weights():
    for name, cfg, cond, hw, B in (("mnist28_uncond", MNIST_UNET, None, 28, 4), ("mnist32_concat", MNIST_UNET, "concatenate", 32, 3)):
        model, sd = build(cfg, cond)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(B, 1, hw, hw, generator=g).to(DEV)
        c = torch.rand(B, 1, hw, hw, generator=g).to(DEV) if cond else None
        for split in (False, True):
            model.set_weight_split(split)
            errs = []
            for tval in (999.0, 500.5, 1.0):
                t = torch.full((B,), tval, device=DEV)
                ref = OD.denoiser_forward(sd, cfg, x, t, conditioning=cond, channels=1, context=c)
                with torch.no_grad():
                    out = model(x, t, context=c)
                errs.append(rel_l2(out, ref))
            print(name, "split" if split else "plain", ["%.3e" % e for e in errs], flush=True)
    # sampling-loop teacher-forced worst (the test_sampling_loop_parity configuration, seed 2)
    model, sd = build(MNIST_UNET, "concatenate", seed=2)
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(4, 1, 32, 32, generator=g).to(DEV)
    cond = torch.rand(4, 1, 32, 32, generator=g).to(DEV)
    for split in (False, True):
        model.set_weight_split(split)
        for sched, steps in (("flowmatch", 50), ("ddim", 50), ("dpmsolver++", 20)):
            orc = make_scheduler(sched, 1000, {"beta_start": 1e-4, "beta_end": 0.02})
            orc.set_timesteps(steps)
            x = noise.cpu()
            worst = 0.0
            for t in orc.timesteps:
                tt = t.expand(4).to(DEV)
                ref_pred = OD.denoiser_forward(sd, MNIST_UNET, x.to(DEV), tt.float(), conditioning="concatenate", channels=1, context=cond)
                with torch.no_grad():
                    my = model(x.to(DEV), tt, context=cond)
                worst = max(worst, rel_l2(my, ref_pred))
                x = orc.step(ref_pred.cpu(), t, x).prev_sample
            print("teacher-forced", sched, "split" if split else "plain", "worst %.3e" % worst, flush=True)
    from fmdm_b200.models.generators import DiffusionUNetFactory
    for name, cfg in (("ca_diffusers_nd", CA_DIFFUSERS), ("ca_efficient_nd", CA_EFFICIENT), ("ca_efficient_nd_linear", CA_EFFICIENT_LINEAR)):
        model = DiffusionUNetFactory().build(cfg, "attention", 1)
        sd = OD.reinit_state_dict(model.state_dict(), 13)
        model.load_state_dict(sd)
        model = model.to(DEV).eval()
        sdd = {k: v.to(DEV) for k, v in sd.items()}
        gold = torch.load(os.path.join(ROOT, "tests", "golden", f"denoiser_{name}.pt"), weights_only=False)
        for split in (False, True):
            model.set_weight_split(split)
            g = torch.Generator().manual_seed(31)
            errs = []
            for hw, ctx_shape in ((32, (2, 4, 8, 8)), (64, (2, 4, 16, 16)), (32, (2, 4, 50)), (32, (2, 50, 4))):
                x = torch.randn(2, 1, hw, hw, generator=g).to(DEV)
                ctx = torch.randn(ctx_shape, generator=g).to(DEV)
                t = torch.tensor([812.0, 33.0], device=DEV)
                ref = OD.denoiser_forward(sdd, cfg, x, t, conditioning="attention", channels=1, context_ca=ctx)
                with torch.no_grad():
                    out = model(x, t, context_ca=ctx)
                errs.append(rel_l2(out, ref))
                ctx2 = ctx * 0.5
                ref2 = OD.denoiser_forward(sdd, cfg, x, t, conditioning="attention", channels=1, context_ca=ctx2)
                with torch.no_grad():
                    out2 = model(x, t, context_ca=ctx2)
                errs.append(rel_l2(out2, ref2))
            with torch.no_grad():
                outg = model(gold["x"].to(DEV), gold["t"].to(DEV), context_ca=gold["context_ca"].to(DEV))
            errs.append(rel_l2(outg.cpu(), gold["out"]))
            print(name, "split" if split else "plain", ["%.3e" % e for e in errs], flush=True)


def exp_compvis_large():
    model, sd = build(COMPVIS_LDCT, "concatenate")
    print("compvis weight_split:", model.weight_split)
    g = torch.Generator().manual_seed(7)
    for hw, B in ((128, 2), (256, 1), (512, 1)):
        x = torch.randn(B, 1, hw, hw, generator=g).to(DEV)
        c = torch.rand(B, 1, hw, hw, generator=g).to(DEV)
        for tval in (999.0, 500.5, 1.0):
            t = torch.full((B,), tval, device=DEV)
            ref = OD.denoiser_forward(sd, COMPVIS_LDCT, x, t, conditioning="concatenate", channels=1, context=c)
            with torch.no_grad():
                out = model(x, t, context=c)
            print("compvis", hw, tval, "%.3e" % rel_l2(out, ref), flush=True)


def exp_b16_512():
    from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler

    model, sd = build(LDCT, "concatenate", seed=6)
    g = torch.Generator().manual_seed(31)
    B, hw = 16, 512
    x = torch.randn(B, 1, hw, hw, generator=g).to(DEV)
    c = torch.rand(B, 1, hw, hw, generator=g).to(DEV)
    for tval in (999.0, 500.5, 1.0):
        t = torch.full((B,), tval, device=DEV)
        torch.cuda.synchronize(); t0 = time.time()
        with torch.no_grad():
            ref = OD.denoiser_forward(sd, LDCT, x, t, conditioning="concatenate", channels=1, context=c)
        torch.cuda.synchronize(); t1 = time.time()
        with torch.no_grad():
            out = model(x, t, context=c)
        per = [rel_l2(out[i], ref[i]) for i in range(B)]
        print("B16@512 t=%g rel_l2 %.3e (per-sample max %.3e) oracle fwd %.2f s" % (tval, rel_l2(out, ref), max(per), t1 - t0), flush=True)
    sched, _ = build_scheduler({"name": "flow_match_euler", "params": {}}, {})
    with torch.no_grad():
        out = sample_with_scheduler(model, sched, 50, x.shape, torch.device(DEV), conditioning_mode="concatenate",
                                    conditioning_batch=c, init_sample=x)
    orc = make_scheduler("flowmatch", 1000, {})
    orc.set_timesteps(50)
    xo = x.clone()
    t0 = time.time()
    with torch.no_grad():
        for t in orc.timesteps:
            pred = OD.denoiser_forward(sd, LDCT, xo, t.expand(B).to(DEV).float(), conditioning="concatenate", channels=1, context=c)
            xo = orc.step(pred, t.to(DEV), xo).prev_sample
    torch.cuda.synchronize()
    print("oracle 50 steps B=16: %.1f s" % (time.time() - t0))
    ps = [psnr(out[i].clamp(0, 1), xo[i].clamp(0, 1)) for i in range(B)]
    print("B16@512 50-step PSNR all %.2f dB, per-sample min %.2f max %.2f" % (psnr(out.clamp(0, 1), xo.clamp(0, 1)), min(ps), max(ps)), flush=True)


def synthetic_pair(b, hw, gen):
    """clean: smooth random field in [0,1]; ldct: clean + 0.05 noise clamped (SURVEY 8d synthetic conditioning)."""
    low = torch.rand(b, 1, hw // 16, hw // 16, generator=gen, device=DEV)
    clean = torch.nn.functional.interpolate(low, size=(hw, hw), mode="bicubic", align_corners=False).clamp_(0, 1)
    ldct = (clean + 0.05 * torch.randn(b, 1, hw, hw, generator=gen, device=DEV)).clamp_(0, 1)
    return clean, ldct


def exp_eps_fixture():
    from fmdm_b200.pipelines.utils import build_scheduler, resolve_scheduler_override, sample_with_scheduler
    from fmdm_b200.training import DiffusionTrainer

    hw = int(os.environ.get("EXP_HW", "256"))
    lr_peak = float(os.environ.get("EXP_LR", "1e-4"))
    warm = int(os.environ.get("EXP_WARMUP", "100"))
    if os.environ.get("EXP_INIT", "default") == "default":
        from fmdm_b200.models.generators import DiffusionUNetFactory
        torch.manual_seed(0)
        model = DiffusionUNetFactory().build(LDCT, "concatenate", 1).to(DEV)
    else:
        model, _ = build(LDCT, "concatenate", seed=5)
    ddpm, _ = build_scheduler({"name": "ddpm", "params": {"beta_start": 1e-4, "beta_end": 0.02}}, {})
    tr = DiffusionTrainer(model, ddpm, lr=lr_peak, weight_decay=0.0)
    print("init", os.environ.get("EXP_INIT", "default"), "lr", lr_peak, "warmup", warm, flush=True)
    gen = torch.Generator(device=DEV).manual_seed(123)
    done = 0
    geval = torch.Generator(device=DEV).manual_seed(77)
    clean_e, ldct_e = synthetic_pair(2, hw, geval)
    noise_e = torch.randn(2, 1, hw, hw, generator=geval, device=DEV)
    for target in [int(v) for v in os.environ.get("EXP_STEPS", "200,500,1000").split(",")]:
        t0 = time.time()
        losses = []
        while done < target:
            clean, ldct = synthetic_pair(16, hw, gen)
            tr.optimizer.param_groups[0]["lr"] = lr_peak * min(1.0, (done + 1) / warm)
            losses.append(tr.step(clean, ldct))
            done += 1
        torch.cuda.synchronize()
        print(f"trained to {done} steps in {time.time() - t0:.1f} s, loss first {float(losses[0]):.4f} last {float(sum(losses[-10:]) / 10):.4f} "
              f"trace {[round(float(l), 3) for l in losses[::max(1, len(losses) // 10)]]}", flush=True)
        model.eval()
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        for sname, steps in (("ddim", 50), ("dpmsolver++", 20)):
            ov = resolve_scheduler_override(sname)
            params = {"beta_start": 1e-4, "beta_end": 0.02}
            params.update(ov.get("params", {}))
            mine, _ = build_scheduler({"name": ov["name"], "params": params}, {})
            with torch.no_grad():
                out = sample_with_scheduler(model, mine, steps, noise_e.shape, torch.device(DEV), conditioning_mode="concatenate",
                                            conditioning_batch=ldct_e, init_sample=noise_e)
            orc = make_scheduler(sname, 1000, {"beta_start": 1e-4, "beta_end": 0.02})
            orc.set_timesteps(steps)
            x = noise_e.clone()
            worst = 0.0
            with torch.no_grad():
                for t in orc.timesteps:
                    tt = t.expand(2).to(DEV).float()
                    pred = OD.denoiser_forward(sd, LDCT, x, tt, conditioning="concatenate", channels=1, context=ldct_e)
                    mp = model(x, tt, context=ldct_e)
                    worst = max(worst, rel_l2(mp, pred))
                    x = orc.step(pred.cpu(), t, x.cpu()).prev_sample.to(DEV)
            print(f"  {sname}-{steps}: PSNR(b200 vs oracle) {psnr(out.clamp(0, 1), x.clamp(0, 1)):.2f} dB; oracle-vs-clean PSNR "
                  f"{psnr(x.clamp(0, 1), clean_e):.2f} dB; teacher-forced worst rel-L2 {worst:.3e}; out range [{float(out.min()):.2f},{float(out.max()):.2f}]",
                  flush=True)
        model.train()


def exp_grad_worst():
    from fmdm_b200.models.generators import DiffusionUNetFactory
    from fmdm_b200.training import diffusion_loss, flow_matching_loss

    SMALL = {"unet_impl": "diffusers_nd", "in_channels": 1, "out_channels": 1, "layers_per_block": 1,
             "block_out_channels": [64, 128, 128],
             "down_block_types": ["DownBlock2D", "AttnDownBlock2D", "DownBlock2D"],
             "up_block_types": ["UpBlock2D", "AttnUpBlock2D", "UpBlock2D"]}
    COMPVIS = {"in_channels": 1, "out_channels": 1, "num_res_blocks": 2, "channel_mult": [1, 1, 2, 2],
               "model_channels": 64, "attention_resolutions": [], "block_out_channels": [64, 64, 128, 128]}
    COMPVIS_ATTN = {"in_channels": 1, "out_channels": 1, "num_res_blocks": 1, "channel_mult": [1, 2],
                    "model_channels": 64, "attention_resolutions": [2], "use_linear_attn": False,
                    "block_out_channels": [64, 128]}
    for name, cfg, hw, b in (("small32", SMALL, 32, 4), ("ldct64", LDCT, 64, 2), ("ldct128", LDCT, 128, 1),
                             ("compvis32", COMPVIS, 32, 2), ("compvis_attn32", COMPVIS_ATTN, 32, 2)):
        model = DiffusionUNetFactory().build(cfg, "concatenate", 1)
        sd = OD.reinit_state_dict(model.state_dict(), 1)
        model.load_state_dict(sd)
        model = model.to(DEV).train()
        sdd = {k: v.to(DEV) for k, v in sd.items()}
        g = torch.Generator().manual_seed(11)
        clean = torch.rand(b, 1, hw, hw, generator=g).to(DEV)
        ldct = torch.rand(b, 1, hw, hw, generator=g).to(DEV)
        noise = torch.randn(b, 1, hw, hw, generator=g).to(DEV)
        t = torch.rand(b, generator=g).to(DEV)
        for kind in ("fm", "eps"):
            for p in model.parameters():
                p.grad = None
            if kind == "fm":
                loss = flow_matching_loss(model, clean, ldct, noise=noise, t=t)
                ref_loss, ref_grads = OT.loss_and_grads(sdd, cfg, clean, ldct, noise, t, 1000)
            else:
                ac = torch.cumprod(1 - torch.linspace(1e-4, 0.02, 1000), 0).to(DEV)
                ts = (t * 999).long()
                loss = diffusion_loss(model, clean, ldct, ac ** 0.5, (1 - ac) ** 0.5, noise=noise, timesteps=ts)
                ref_loss, ref_grads = OT.loss_and_grads(sdd, cfg, clean, ldct, noise, ts, 1000, alphas_cumprod=ac)
            loss.backward()
            rows = []
            fa, fb = [], []
            for k, p in model.named_parameters():
                gr, r = p.grad.float(), ref_grads[k]
                fa.append(gr.reshape(-1)); fb.append(r.reshape(-1))
                rows.append((rel_l2(gr, r), k, float(r.norm()), r.numel()))
            total_norm = float(torch.cat(fb).norm())
            rows.sort(reverse=True)
            print(name, kind, "loss %.5f ref %.5f total rel %.3e" % (float(loss), float(ref_loss), rel_l2(torch.cat(fa), torch.cat(fb))))
            for e, k, n, ne in rows[:6]:
                print("    %.3e  %-60s |g|=%.3e (%.2e of total) numel %d" % (e, k, n, n / total_norm, ne))
            # error of each parameter's gradient relative to the TOTAL gradient norm share
            worst_scaled = max((float((p.grad.float() - ref_grads[k]).norm()) / total_norm, k) for k, p in model.named_parameters())
            print("    worst |dg|/|g_total| = %.3e (%s)" % worst_scaled, flush=True)


def exp_attention_micro():
    from fmdm_b200 import ops

    b, heads, t, hd = 16, 64, 1024, 8
    c = heads * hd
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(b, t, 3 * c, generator=g).to(DEV).to(torch.bfloat16)
    out = torch.empty(b, t, c, device=DEV, dtype=torch.bfloat16)
    flat = qkv.view(-1)
    def run():
        ops.attention(flat, flat[c:], flat[2 * c:], out.view(-1), batch=b, heads=heads, tq=t, tk=t, head_dim=hd,
                      q_strides=(t * 3 * c, hd, 3 * c), kv_strides=(t * 3 * c, hd, 3 * c), o_strides=(t * c, hd, c))
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        run()
    e1.record()
    torch.cuda.synchronize()
    q, k, v = [x.reshape(b, t, heads, hd).transpose(1, 2).float() for x in qkv.split(c, dim=-1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, t, c)
    print("attention hd8 T=1024 B=16 heads=64: %.4f ms/launch, bf16p=%s, rel_l2 vs fp32 SDPA %.3e" % (
        e0.elapsed_time(e1) / 200, os.environ.get("FMDM_ATTENTION_BF16P") is not None, rel_l2(out, ref)), flush=True)


def exp_mnist_steps():
    """Two eager sampling steps of the MNIST config (for an ncu launch list) and the graph-replayed run time."""
    from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler

    model, _ = build(MNIST_UNET, None)
    sched, _ = build_scheduler({"name": "flow_match_euler", "params": {}}, {})
    x = torch.randn(64, 1, 28, 28, device=DEV)
    with torch.no_grad():
        for graph in (False, True, True):
            torch.cuda.synchronize(); t0 = time.time()
            sample_with_scheduler(model, sched, 50, tuple(x.shape), torch.device(DEV), init_sample=x, last_n_steps=2 if not graph else None,
                                  use_cuda_graph=graph)
            torch.cuda.synchronize()
            print("mnist", "graph 50 steps" if graph else "eager 2 steps", "%.2f ms" % ((time.time() - t0) * 1e3), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["split", "compvis", "grad", "eps", "b16"]
    table = {"split": exp_split_weights, "compvis": exp_compvis_large, "grad": exp_grad_worst, "eps": exp_eps_fixture,
             "b16": exp_b16_512, "att": exp_attention_micro, "mnist": exp_mnist_steps}
    for w in which:
        section(table[w])
