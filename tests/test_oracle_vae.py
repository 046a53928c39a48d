"""Pin the oracle's AutoencoderKL.decode restatement: against the golden fixtures generated from the reference's own
module (tests/golden/vae_decode_*.pt, oracle/make_golden_vae.py) and, when /root/reference is present, against the live
reference module."""
import json
import os
import sys
import warnings

import pytest
import torch

from oracle import vae_decoder as OV
from oracle.denoiser import reinit_state_dict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["ldct_kl", "small_attn"]


def _state_from_manifest(name, seed):
    man = json.load(open(os.path.join(GOLD, f"state_keys_vae_{name}.json")))
    shapes = {k: torch.empty(shape) for k, shape in man["keys"]}
    return reinit_state_dict(shapes, seed), man["cfg"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_decode_matches_reference_golden(name):
    gold = torch.load(os.path.join(GOLD, f"vae_decode_{name}.pt"), weights_only=True)
    sd, cfg = _state_from_manifest(name, gold["seed"])
    assert cfg == gold["cfg"]
    with torch.no_grad():
        raw = OV.kl_decode(sd, cfg, gold["z"], denorm=True)
    assert raw.shape == gold["raw"].shape
    assert float((raw - gold["raw"]).abs().max()) < 2e-5
    assert float((OV.raw_output_to_image(raw) - gold["image"]).abs().max()) < 2e-5


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="live reference tree not present")
def test_oracle_decode_matches_live_reference():
    sys.path.insert(0, "/root/reference/src")
    try:
        from models.vae.kl import AutoencoderKL
    finally:
        sys.path.remove("/root/reference/src")
    warnings.simplefilter("ignore")
    cfg = dict(in_channels=1, out_channels=1, resolution=64, down_channels=(32, 64, 64), num_res_blocks=1,
               attn_resolutions=(32,), z_channels=4, embed_dim=4, use_attention=True, attn_heads=2, attn_dim_head=16)
    torch.manual_seed(0)
    ref = AutoencoderKL(**cfg).eval()
    sd = reinit_state_dict(ref.state_dict(), 5)
    ref.load_state_dict(sd)
    z = torch.randn(2, 4, 16, 16)
    with torch.no_grad():
        want = ref.decode(z, denorm=False)
        got = OV.kl_decode(sd, dict(cfg), z, denorm=False)
    assert float((want - got).abs().max()) < 2e-5
