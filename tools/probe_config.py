"""One of the smaller BASELINE configs through the public sampling API, alone (for ncu launch lists and same-box A/B):
    python tools/probe_config.py mnist [steps] [reps]      configs[0], MNIST 28x28, batch 64
    python tools/probe_config.py latent [steps] [reps]     configs[2], latent 64x64x4 denoiser only, batch 128
Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "mnist"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    peaks = {"tf_sustained": 1375.4}
    if which == "mnist":
        rec = bench.extra_sampling_config(dev, peaks, name="configs[0] MNIST 28x28, batch 64", cfg=None,
                                          manifest="mnist_diffusers_nd_uncond", cond=None, B=64, hw=28,
                                          sched="flowmatch", steps=steps, flop_fwd=bench.FLOP_MNIST28_FWD, cpu=None,
                                          reps=reps)
    else:
        rec = bench.extra_latent_config(dev, peaks, cpu=False, steps=steps, reps=reps)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
