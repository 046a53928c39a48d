"""Marker base class for blocks whose forward takes (x, emb) (`src/nn/blocks/timestep.py:13-23`)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Optional

import torch
import torch.nn as nn


class TimestepBlock(nn.Module, ABC):
    @abstractmethod
    def forward(self, x: torch.Tensor, emb: Optional[torch.Tensor]) -> torch.Tensor:  # pragma: no cover
        raise NotImplementedError
