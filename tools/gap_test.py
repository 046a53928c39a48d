import sys, time, torch
sys.path.insert(0, "/root/repo")
from bench import LDCT_UNET, synthetic_inputs
from fmdm_b200 import ops
from fmdm_b200.models.generators import DiffusionUNetFactory
from fmdm_b200.pipelines.utils import build_scheduler, sample_with_scheduler
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = DiffusionUNetFactory().build(LDCT_UNET, "concatenate", 1).to(dev).eval()
noise, cond = synthetic_inputs(16, 42, dev)
sch, _ = build_scheduler({"name": "flow_match_euler", "params": {}}, {})
def run(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sample_with_scheduler(model, sch, 50, tuple(noise.shape), dev, conditioning_mode="concatenate", conditioning_batch=cond, init_sample=noise, last_n_steps=steps)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
with torch.no_grad():
    run(50); run(2)
    # cool GPU: single steps separated by sleeps
    cold = []
    for _ in range(6):
        time.sleep(1.0)
        cold.append(run(1))
    hot = run(50)
    print("ms per step, 1-step runs after 1 s idle:", [round(c, 2) for c in cold])
    print("ms per step, 50-step run:", round(hot, 2))
