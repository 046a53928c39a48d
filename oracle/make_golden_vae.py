"""ORACLE tooling: golden fixtures for AutoencoderKL.decode, produced by RUNNING THE REFERENCE'S OWN MODULE
(/root/reference/src, CPU fp32).  Build container only:  python oracle/make_golden_vae.py

  tests/golden/vae_decode_<name>.pt : {"cfg", "seed", "z", "raw", "image"}; weights = reinit_state_dict(reference
                                      state_dict, seed) restricted to the decode path (re-derived in the tests)
  tests/golden/state_keys_vae_<name>.json : ordered [key, shape] list of the decode-path parameters
"""
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from models.vae.kl import AutoencoderKL  # noqa: E402  (reference)

from oracle.denoiser import reinit_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
KEEP = ("in_channels", "out_channels", "resolution", "base_ch", "down_channels", "num_res_blocks", "attn_resolutions",
        "z_channels", "embed_dim", "dropout", "use_attention", "attn_heads", "attn_dim_head", "spatial_dims",
        "emb_channels", "use_scale_shift_norm")

CASES = {
    # the LDCT KL config (configs/LDCT/LDCT_autoencoder_kl.json) on an 8x8 latent -> 64x64 image
    "ldct_kl": dict(cfg_path="LDCT/LDCT_autoencoder_kl.json", over={}, hw=8, B=1),
    # narrow variant with attention inside a stage too
    "small_attn": dict(cfg_path="LDCT/LDCT_autoencoder_kl.json",
                       over={"down_channels": [64, 128], "resolution": 32, "attn_resolutions": [16], "attn_heads": 2,
                             "attn_dim_head": 32}, hw=16, B=2),
}


def main():
    warnings.simplefilter("ignore")
    for name, c in CASES.items():
        cfg = json.load(open(os.path.join("/root/reference/configs", c["cfg_path"])))["model"]
        cfg.update(c["over"])
        kw = {k: cfg[k] for k in KEEP if k in cfg}
        if kw.get("down_channels") is not None:
            kw["down_channels"] = tuple(kw["down_channels"])
        kw["attn_resolutions"] = tuple(kw.get("attn_resolutions", ()))
        torch.manual_seed(0)
        model = AutoencoderKL(**kw).eval()
        seed = 17
        full = reinit_state_dict(model.state_dict(), seed)
        model.load_state_dict(full)
        keys = [[k, list(v.shape)] for k, v in model.state_dict().items()
                if k.startswith("decoder.") or k.startswith("post_quant_conv.")]
        with open(os.path.join(GOLD, f"state_keys_vae_{name}.json"), "w") as f:
            json.dump({"cfg": cfg, "keys": keys}, f)
        g = torch.Generator().manual_seed(321)
        z = torch.randn(c["B"], cfg["z_channels"], c["hw"], c["hw"], generator=g)
        with torch.no_grad():
            raw = model.decode(z, denorm=True)
            img = model.raw_output_to_image(raw, recon_type="l1")
        torch.save({"cfg": cfg, "seed": seed, "z": z, "raw": raw, "image": img},
                   os.path.join(GOLD, f"vae_decode_{name}.pt"))
        print(name, tuple(raw.shape), float(raw.abs().max()), len(keys))


if __name__ == "__main__":
    main()
