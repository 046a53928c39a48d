"""Mirror of the reference's `src/nn` building-block API (`src/nn/__init__.py:7-58`) for the sampling hot path."""
from .blocks import *  # noqa: F401,F403
from .blocks import __all__ as _blocks_all
from .ops import *  # noqa: F401,F403
from .ops import __all__ as _ops_all

__all__ = list(_blocks_all) + list(_ops_all)
