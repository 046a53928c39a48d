"""Import alias: `fmdm_b200` -> the `flow-matching-and-diffusion-models_b200/` package directory.

The package directory carries the reference repository's name (hyphens and all), which Python cannot import by
name; this shim makes its sub-modules importable as `fmdm_b200.<module>` (e.g. `fmdm_b200.nn`, `fmdm_b200.ops`).
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "flow-matching-and-diffusion-models_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
