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


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof of the parameter structs as gcc sees the header == the ctypes mirror."""
    import subprocess
    from pathlib import Path

    from fmdm_b200 import _lib

    root = Path(__file__).resolve().parent.parent
    fields = {"fm_conv_seg": [f[0] for f in _lib.ConvSeg._fields_], "fm_conv_params": [f[0] for f in _lib.ConvParams._fields_],
              "fm_pack_entry": [f[0] for f in _lib.PackEntry._fields_]}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "fmdm_b200.h"', 'int main(void) {']
    for st, names in fields.items():
        src.append(f'  printf("{st} %zu\\n", sizeof({st}));')
        for n in names:
            src.append(f'  printf("{st}.{n} %zu\\n", offsetof({st}, {n}));')
    src += ['  return 0;', '}']
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(root / "include"), str(cfile), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    mirror = {"fm_conv_seg": _lib.ConvSeg, "fm_conv_params": _lib.ConvParams, "fm_pack_entry": _lib.PackEntry}
    for st, cls in mirror.items():
        assert int(got[st]) == ctypes.sizeof(cls), st
        for n in fields[st]:
            assert int(got[f"{st}.{n}"]) == getattr(cls, n).offset, f"{st}.{n}"


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


def test_pdl_switch_is_scoped(monkeypatch):
    """`ops.pdl` sets the programmatic-dependent-launch mode for the launches inside and restores what was there
    (host-side state only: works without a GPU); nested scopes unwind in order."""
    from fmdm_b200 import _lib, ops

    handle = _lib.lib()
    start = handle.fm_set_pdl(0)
    try:
        with ops.pdl(True):
            assert handle.fm_set_pdl(1) == 1          # on inside the scope (the call returns the previous setting)
            with ops.pdl(False):
                assert handle.fm_set_pdl(0) == 0
            assert handle.fm_set_pdl(1) == 1          # the inner scope restored "on"
        assert handle.fm_set_pdl(0) == 0              # the outer scope restored "off"
    finally:
        handle.fm_set_pdl(start)
