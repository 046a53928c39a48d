"""ORACLE tooling: golden fixtures of the training step, produced by RUNNING THE REFERENCE'S OWN model class (imported
read-only from /root/reference/src) through torch autograd and torch.optim.AdamW on CPU fp32, with the loop body of
`src/pipelines/train/flow_matching_lib.py:150-176` (that module itself imports `diffusers`, which this image does not
have, so its ten lines of tensor arithmetic are restated here around the reference model).

    python oracle/make_golden_train.py

tests/golden/train_step_<name>.pt: {"cfg", "seed", "clean", "ldct", "noise", "t", "lr", "weight_decay", "losses"
(one per step), "grad_norms" {key: L2 norm of the step-1 gradient}, "grad_probe" {key: first 16 values of a few
gradients}, "param_checksums" {key: (sum, abs-sum) after the last step}}."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from models.generators.diffusionfactory import DiffusionUNetFactory  # noqa: E402  (reference)

from oracle.denoiser import reinit_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {
    "ldct_diffusers_nd": dict(cfg_path="LDCT/LDCT_flow_matching_diffusers_nd.json", hw=32, B=2, steps=2),
    "mnist_diffusers_nd": dict(cfg_path="MNIST/mnist_flow_matching_diffusers_nd.json", hw=16, B=3, steps=3),
    "ldct_compvis": dict(cfg_path="LDCT/LDCT_flow_matching_compvis.json", hw=32, B=2, steps=2),  # EfficientUNetND
}


def main():
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        cfg = json.load(open(os.path.join("/root/reference/configs", c["cfg_path"])))["model"]["unet"]
        torch.manual_seed(0)
        model = DiffusionUNetFactory().build(cfg, "concatenate", 1).train()
        seed = 17
        model.load_state_dict(reinit_state_dict(model.state_dict(), seed))
        lr, wd = 1e-3, 0.01
        opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd)
        g = torch.Generator().manual_seed(321)
        B, hw = c["B"], c["hw"]
        clean = torch.rand(B, 1, hw, hw, generator=g)
        ldct = torch.rand(B, 1, hw, hw, generator=g)
        noise = torch.randn(B, 1, hw, hw, generator=g)
        t = torch.rand(B, generator=g)
        losses, grad_norms, probe = [], {}, {}
        for step in range(c["steps"]):
            opt.zero_grad(set_to_none=True)
            timesteps = (t * (1000 - 1)).long()                                        # flow_matching_lib.py:152
            x_t = (1.0 - t[:, None, None, None]) * clean + t[:, None, None, None] * noise  # :153
            pred = model(torch.cat([x_t, ldct], dim=1), timesteps)                     # :156, :162
            loss = F.mse_loss(pred, noise - clean)                                     # :163-164
            loss.backward()                                                            # :169
            if step == 0:
                for k, p in model.named_parameters():
                    grad_norms[k] = float(p.grad.norm())
                keys = [k for k, _ in model.named_parameters()]
                for k in keys[:3] + keys[len(keys) // 2:len(keys) // 2 + 3] + keys[-3:]:
                    probe[k] = dict(model.named_parameters())[k].grad.reshape(-1)[:16].clone()
            opt.step()                                                                 # :176
            losses.append(float(loss))
        sums = {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in model.state_dict().items()
                if v.is_floating_point()}
        torch.save({"cfg": cfg, "seed": seed, "clean": clean, "ldct": ldct, "noise": noise, "t": t, "lr": lr,
                    "weight_decay": wd, "losses": losses, "grad_norms": grad_norms, "grad_probe": probe,
                    "param_checksums": sums}, os.path.join(GOLD, f"train_step_{name}.pt"))
        print(name, losses, len(grad_norms))


if __name__ == "__main__":
    main()
