"""Host-side runtime helpers shared by the module mirror: packed-weight caches and the out-of-scope policy."""
from __future__ import annotations

from typing import Callable, Dict, NoReturn, Tuple

import torch


class OutOfScopeError(NotImplementedError):
    """Raised when a module variant outside the accelerated hot path is executed."""


def out_of_scope(what: str) -> NoReturn:
    """Variants the north star does not name (1-D/3-D convs, RMSNorm blocks, pooling, ...) have no sm_100a kernel and
    are refused loudly: there is no eager-PyTorch, library or CPU path anywhere in this package."""
    raise OutOfScopeError(
        f"fmdm_b200: {what} is outside the B200 hot path (SURVEY.md §8, marked out of scope); this package has no "
        "eager/CPU fallback - use the reference implementation for it."
    )


class ParamCache:
    """Caches a derived device tensor (e.g. a K-major bf16 weight) keyed on the identity + version of its source
    parameters, so `load_state_dict`, `.to()` and in-place updates invalidate it."""

    def __init__(self):
        self._store: Dict[str, Tuple[tuple, object]] = {}

    @staticmethod
    def _sig(params) -> tuple:
        return tuple((p.data_ptr(), p._version, str(p.device), p.dtype) for p in params if p is not None)

    def get(self, key: str, params, build: Callable[[], object]):
        sig = self._sig(params)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("fmdm_b200: weight packing requested during CUDA-graph capture; run one warm-up "
                               "forward before capturing")
        val = build()
        self._store[key] = (sig, val)
        return val

    def clear(self):
        self._store.clear()


def f32(p):
    """fp32 contiguous view of a parameter (None passes through)."""
    if p is None:
        return None
    t = p.detach()
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t
