"""Shared block helpers (`src/nn/blocks/common.py`)."""
from __future__ import annotations


def zero_module(module):
    """Set every parameter of `module` to zero in place and hand the module back."""
    for param in module.parameters():
        param.detach().zero_()
    return module
