"""Pins the training-step oracle (oracle/training.py) to the golden fixtures generated from the reference's own model
class + torch autograd + torch.optim.AdamW (oracle/make_golden_train.py): losses per step, step-1 gradient norms of
every parameter, probed gradient values, and the parameter checksums after the last optimiser step."""
import glob
import json
import os

import pytest
import torch

from oracle import denoiser as OD
from oracle import training as OT

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[len("train_step_"):-3] for p in glob.glob(os.path.join(GOLD, "train_step_*.pt")))


def _state(name, seed):
    with open(os.path.join(GOLD, f"state_keys_{name}.json")) as f:
        meta = json.load(f)
    return OD.reinit_state_dict({k: torch.zeros(shape) for k, shape in meta["keys"]}, seed)


@pytest.mark.parametrize("name", CASES)
def test_training_oracle_matches_reference_golden(name):
    gold = torch.load(os.path.join(GOLD, f"train_step_{name}.pt"), weights_only=False)
    sd = _state(name, gold["seed"])
    batch = (gold["clean"], gold["ldct"], gold["noise"], gold["t"])
    loss, grads = OT.loss_and_grads(sd, gold["cfg"], *batch)
    assert abs(float(loss) - gold["losses"][0]) <= 1e-5 * abs(gold["losses"][0])
    assert set(gold["grad_norms"]) <= set(grads)
    for k, n in gold["grad_norms"].items():
        assert abs(float(grads[k].norm()) - n) <= 1e-4 * max(n, 1e-6) + 1e-7, k
    for k, v in gold["grad_probe"].items():
        assert torch.allclose(grads[k].reshape(-1)[:16], v, rtol=1e-3, atol=1e-6), k
    losses, final = OT.train_steps(sd, gold["cfg"], [batch] * len(gold["losses"]), lr=gold["lr"],
                                   weight_decay=gold["weight_decay"])
    for a, b in zip(losses, gold["losses"]):
        assert abs(a - b) <= 2e-4 * abs(b), (losses, gold["losses"])
    for k, (s, a) in gold["param_checksums"].items():
        assert abs(float(final[k].double().abs().sum()) - a) <= 1e-4 * a + 1e-6, k


def test_oracle_adamw_matches_torch_optim():
    """The oracle's AdamW restatement (oracle/training.py::adamw_step) against torch.optim.AdamW on CPU, several steps,
    weight decay on: pins the optimiser arithmetic independently of the model fixtures."""
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(257, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 8):
        grad = torch.randn(257, generator=g)
        ref.grad = grad.clone()
        opt.step()
        p, m, v = OT.adamw_step(p, grad, m, v, step, lr=3e-3, weight_decay=0.05)
        assert torch.allclose(p, ref.detach(), rtol=2e-6, atol=1e-7), step
