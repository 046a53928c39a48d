"""The C-ABI library loads on a CPU-only machine, exports every symbol include/fmdm_b200.h declares, and every
compute entry point refuses to run without an sm_100 device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fmdm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from fmdm_b200 import _lib

    handle = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/fmdm_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared)
    assert handle.fm_version() >= 100


def test_struct_layout_matches_header():
    from fmdm_b200 import _lib

    assert ctypes.sizeof(_lib.ConvSeg) == 24
    # 4 segs (96) + 7 int32 (28, padded to 32) ... pointers 8-aligned
    assert ctypes.sizeof(_lib.ConvParams) == 96 + 32 + 8 * 3 + 8 + 8 * 3 + 8
    assert _lib.ConvParams.weight.offset == 128 and _lib.ConvParams.out.offset == 168


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from fmdm_b200 import _lib, ops

    handle = _lib.lib()
    rc = handle.fm_sched_flowmatch_f32(None, None, None, None, None, 0, 16, None)
    assert rc != 0 and b"no CPU path" in handle.fm_last_error()
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.sched_flowmatch(torch.zeros(4), torch.zeros(4), torch.zeros(1), 0)
    from fmdm_b200.models.generators import DiffusionUNetFactory

    model = DiffusionUNetFactory().build({"unet_impl": "diffusers_nd", "block_out_channels": [32, 32],
                                          "down_block_types": ["DownBlock2D"] * 2,
                                          "up_block_types": ["UpBlock2D"] * 2}, None, 1)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        model(torch.zeros(1, 1, 8, 8), torch.zeros(1))


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from fmdm_b200 import _lib

    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setenv("FMDM_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="not built"):
        _lib.lib()
