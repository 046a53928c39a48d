"""AdamW over ONE flat fp32 parameter buffer: a single kernel launch per optimiser step.

Drop-in for the `torch.optim.AdamW(model.parameters(), lr, weight_decay)` of `flow_matching_lib.py:74`: same
hyper-parameters, same update arithmetic (decoupled weight decay, bias correction, eps outside the square root), same
`step()/zero_grad()/state_dict()` surface.  The parameters are re-seated as views of one contiguous buffer and their
`.grad`s as views of a second one, so autograd accumulates gradients straight into the flat buffer, the data-parallel
all-reduce works on contiguous slices of it (`training.ddp`) and the update is `fm_adamw_f32` over the whole thing."""
from __future__ import annotations

from typing import Iterable

import torch

from .. import _lib
from ..ops import _stream


class FlatBuffers:
    """Re-seats `params` (fp32, one device) onto a flat buffer; `.grad`s live in `self.grad`.

    Call after the model is on its final device (a later `.to()` / `load_state_dict(assign=True)` would detach the
    parameters from the buffer; plain `load_state_dict` copies in place and is fine).
    The flat order is the REVERSE of the given order: backward produces gradients roughly last-layer-first, so
    buckets of consecutive flat ranges complete in order (see `training.ddp.BucketedAllReduce`)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], order=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatBuffers: no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatBuffers: parameters must be fp32 on one device")
        if order is None:
            order = list(reversed(self.params))
        else:   # a caller-chosen layout (`training.graph.flat_param_order`): must be a permutation of the parameters
            order = [p for p in order if p.requires_grad]
            if len(order) != len(self.params) or {id(p) for p in order} != {id(p) for p in self.params}:
                raise ValueError("FlatBuffers: `order` must list every trainable parameter exactly once")
        self.offsets = {}
        off = 0
        for p in order:
            self.offsets[id(p)] = off
            off += (p.numel() + 3) // 4 * 4  # keep every slice 16-byte aligned
        self.numel = off
        self.data = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p in order:
                o, n = self.offsets[id(p)], p.numel()
                self.data[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.data[o:o + n].view(p.shape)
                p.grad = self.grad[o:o + n].view(p.shape)

    def slice_of(self, p) -> tuple:
        return self.offsets[id(p)], p.numel()

    def ensure_grad_views(self) -> None:
        """Re-attach `.grad` views (after a `zero_grad(set_to_none=True)` from foreign code)."""
        for p in self.params:
            o, n = self.slice_of(p)
            g = p.grad
            if g is None or g.data_ptr() != self.grad.data_ptr() + 4 * o:
                view = self.grad[o:o + n].view(p.shape)
                if g is not None:
                    view.copy_(g)
                p.grad = view


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 flat_order=None):
        params = list(params)
        if any(isinstance(p, dict) for p in params):
            raise ValueError("FusedAdamW: a single parameter group (flow_matching_lib.py:74 passes model.parameters())")
        if any(p.device.type != "cuda" for p in params):
            raise RuntimeError("fmdm_b200.training.FusedAdamW: parameters must live on a CUDA device (the optimiser is "
                               "one sm_100a kernel; there is no CPU implementation)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.flat = FlatBuffers(self.param_groups[0]["params"], order=flat_order)
        dev = self.flat.data.device
        self.exp_avg = torch.zeros_like(self.flat.data)
        self.exp_avg_sq = torch.zeros_like(self.flat.data)
        self.step_count = 0
        self.grad_scale = 1.0  # folded into the kernel's gradient read (1/world_size of the data-parallel mean)
        assert dev.type == "cuda", "FusedAdamW needs CUDA parameters (no CPU implementation)"

    def zero_grad(self, set_to_none: bool = True) -> None:
        # the gradients are views of one buffer: clear it in one launch and keep the views
        self.flat.ensure_grad_views()
        _lib.check(_lib.lib().fm_memset_f32(self.flat.grad.data_ptr(), self.flat.numel, _stream()), "memset")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.flat.ensure_grad_views()
        g = self.param_groups[0]
        self.step_count += 1
        b1, b2 = g["betas"]
        _lib.check(
            _lib.lib().fm_adamw_f32(self.flat.data.data_ptr(), self.flat.grad.data_ptr(), self.exp_avg.data_ptr(),
                                    self.exp_avg_sq.data_ptr(), self.flat.numel, float(g["lr"]), float(b1), float(b2),
                                    float(g["eps"]), float(g["weight_decay"]), self.step_count,
                                    float(self.grad_scale), _stream()),
            "adamw",
        )
        # the inference-side packed-weight caches key on Parameter._version, which the flat update bypasses
        torch._C._increment_version(self.flat.params)
        return loss

    # checkpoint surface: the torch.optim layout, so the 'optimizer' entry of the reference's {flow,diff}_{best,last}.pt
    # (`flow_matching_lib.py:197-211`, written by torch.optim.AdamW) resumes here and checkpoints written here resume
    # in the reference.  state[i] are per-parameter COPIES of the flat moment buffers, i = position in param_groups.
    def state_dict(self):
        params = self.param_groups[0]["params"]
        state = {}
        if self.step_count > 0:
            for i, p in enumerate(params):
                if id(p) not in self.flat.offsets:
                    continue
                o, n = self.flat.slice_of(p)
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(params)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, state_dict):
        if not isinstance(state_dict, dict) or "state" not in state_dict or "param_groups" not in state_dict:
            raise ValueError("FusedAdamW.load_state_dict expects the torch.optim layout {'state', 'param_groups'} "
                             f"(got keys {sorted(state_dict) if isinstance(state_dict, dict) else type(state_dict)})")
        groups = state_dict["param_groups"]
        params = self.param_groups[0]["params"]
        if len(groups) != 1 or len(groups[0].get("params", [])) != len(params):
            raise ValueError("FusedAdamW.load_state_dict: expected one parameter group with "
                             f"{len(params)} parameters, got {[len(g.get('params', [])) for g in groups]}")
        ids = list(groups[0]["params"])
        steps = set()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, p in enumerate(params):
            st = state_dict["state"].get(ids[i])
            if st is None or id(p) not in self.flat.offsets:
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"FusedAdamW.load_state_dict: state {ids[i]} has shape {tuple(st['exp_avg'].shape)}, "
                                 f"parameter has {tuple(p.shape)}")
            o, n = self.flat.slice_of(p)
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FusedAdamW.load_state_dict: per-parameter step counts differ ({sorted(steps)}); the flat "
                             "update keeps one step counter")
        self.step_count = steps.pop() if steps else 0
        self.param_groups[0].update({k: v for k, v in groups[0].items() if k != "params"})
