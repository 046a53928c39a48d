"""BaseUNetND — the forward contract of the denoisers (`src/models/unet/base.py:10-53`)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import torch
import torch.nn as nn


class BaseUNetND(nn.Module, ABC):
    """forward(x, t, context=None, context_ca=None) -> prediction with the shape/dtype of the reference (fp32 NCHW)."""

    def _normalize_timesteps(self, t, x: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(t):
            t = torch.tensor([t], device=x.device, dtype=torch.long)  # python scalars become int64, as upstream
        if t.ndim == 0:
            t = t[None].to(x.device)
        return t.expand(x.shape[0]).to(x.device)

    def _prepare_input(self, x, context, context_ca):
        return x

    # narrow denoisers (MNIST, the PixelAttention configs) sit AT the north star's 1e-2 per-step tolerance with plain
    # bf16 weights: half of their error is the weight rounding, and with <= 128 channels there is little averaging.
    # They are launch-bound, so their GEMMs read split-bf16 weights (w_hi + w_lo, `ops.PackedConvWeight`) for free.
    SPLIT_WEIGHT_MAX_CHANNELS = 128

    def set_weight_split(self, flag: bool) -> None:
        self.weight_split = bool(flag)
        for m in self.modules():
            if m is not self and hasattr(type(m), "weight_split"):
                m.weight_split = bool(flag)

    def _pack_temb(self, emb: torch.Tensor):
        """Project `emb` for every ResBlock of the model with one batched fp32 kernel (see nn.blocks.TembPack)."""
        from ... import ops
        from ..._runtime import ParamCache
        from ...nn.blocks.residual import ResBlockND, TembPack

        if not hasattr(self, "_temb_blocks"):
            self._temb_blocks = [m for m in self.modules() if isinstance(m, ResBlockND) and m.uses_embedding]
            self._temb_params = [p for m in self._temb_blocks
                                 for p in (m.emb_layers.weight, m.emb_layers.bias, m.conv1.conv.bias)]
            self._temb_cache = ParamCache()
        if not self._temb_blocks:
            return emb
        plan = self._temb_cache.get("plan", self._temb_params, lambda: TembPack.plan(self._temb_blocks))
        if plan is None:
            return emb
        weight, bias, offsets, silu = plan
        if emb.stride(0) == 0:  # one timestep for the whole batch (sampling loop): project one row, broadcast it
            proj = ops.linear_f32(emb[:1], weight, bias, silu_in=silu).expand(emb.shape[0], -1)
        else:
            proj = ops.linear_f32(emb, weight, bias, silu_in=silu)
        return TembPack(emb, proj, offsets)

    def _wants_autograd(self, x: torch.Tensor, t_table) -> bool:
        """Autograd is recording, the module is in training mode and has trainable parameters: run the
        differentiable graph (`training.graph`).  Sampling runs under `torch.no_grad()` / `.eval()` and is unaffected."""
        if not (torch.is_grad_enabled() and self.training and x.is_cuda and t_table is None):
            return False
        from ...training import graph

        return graph.supported(self) and any(p.requires_grad for p in self.parameters())

    @abstractmethod
    def _build_time_embedding(self, t: Optional[torch.Tensor], x: torch.Tensor, *, t_table=None,
                              step_dev=None) -> torch.Tensor:
        raise NotImplementedError

    @abstractmethod
    def _run_network(self, x, emb: torch.Tensor, context_ca: Optional[torch.Tensor]) -> torch.Tensor:
        raise NotImplementedError

    def _postprocess_output(self, y: torch.Tensor) -> torch.Tensor:
        return y

    def forward(self, x: torch.Tensor, t, context: Optional[torch.Tensor] = None,
                context_ca: Optional[torch.Tensor] = None, *, t_table: Optional[torch.Tensor] = None,
                step_dev: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        """`t_table`/`step_dev` (fp32 device table + int32 device cursor) replace `t` under CUDA-graph replay:
        every sample of the batch then uses the timestep `t_table[*step_dev]`."""
        if x.is_cuda and x.device.index != torch.cuda.current_device():
            # the kernels launch on the current device's stream: follow the input tensor (run_model --device cuda:1)
            with torch.cuda.device(x.device):
                return self.forward(x, t, context, context_ca, t_table=t_table, step_dev=step_dev, **kwargs)
        if self._wants_autograd(x, t_table):
            # training step (`flow_matching_lib.py:158-169`): same modules and parameters, differentiable kernels
            from ...training import graph

            return self._postprocess_output(graph.unet_forward(self, x, t, context, context_ca))
        if t_table is not None:
            # the sampling loop feeds one timestep to the whole batch (`pipelines/utils.py:206-209` expands a scalar):
            # the time embedding and every projection of it are computed for ONE row and broadcast (stride 0)
            emb = self._build_time_embedding(None, x[:1], t_table=t_table, step_dev=step_dev).expand(x.shape[0], -1)
        else:
            t = self._normalize_timesteps(t, x)
            emb = self._build_time_embedding(t, x)
        y = self._run_network(self._prepare_input(x, context, context_ca), emb, context_ca)
        return self._postprocess_output(y)
