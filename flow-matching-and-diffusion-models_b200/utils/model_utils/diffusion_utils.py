"""Model construction / checkpoint loading and the per-batch decode entry of the sampling path.

Drop-in for the functions of `src/utils/model_utils/diffusion_utils.py` that sit directly above the hot loop:
`build_diffusion_model` (:88-144, incl. the legacy diffusers key remap :15-85), `encode_diffusion_batch` (:147-162),
`decode_diffusion_batch` (:165-245) and `warn_attention_conditioning_shape` (:248-272).  Same signatures, same
scheduler-override / timestep-subset / add_noise-initialisation behaviour; the model is the B200 module mirror and the
loop is `fmdm_b200.pipelines.utils.sample_with_scheduler` (CUDA-graph replayed).  Dataset helpers of the reference file
(`prepare_diffusion_visual_batch`) are out of scope.
"""
from __future__ import annotations

import logging
from typing import Dict, Optional, Tuple

import torch

from ...models.generators import DiffusionUNetFactory
from ...pipelines.utils import (build_scheduler, resolve_conditioning_mode, resolve_scheduler_override,
                                sample_with_scheduler)

# (substring in a legacy / diffusers checkpoint key, its name in this module tree); applied in order to every key
_LEGACY_KEY_RULES: Tuple[Tuple[str, str], ...] = (
    # attention projections
    (".query.", ".to_q."), (".key.", ".to_k."), (".value.", ".to_v."), (".proj_attn.", ".to_out.0."),
    # ResNet block convs / time projection / shortcut live one wrapper deeper here (ConvND holds `.conv`)
    (".conv1.weight", ".conv1.conv.weight"), (".conv1.bias", ".conv1.conv.bias"),
    (".conv2.weight", ".conv2.conv.weight"), (".conv2.bias", ".conv2.conv.bias"),
    (".time_emb_proj.weight", ".emb_layers.weight"), (".time_emb_proj.bias", ".emb_layers.bias"),
    (".conv_shortcut.weight", ".skip_connection.conv.weight"), (".conv_shortcut.bias", ".skip_connection.conv.bias"),
    # resamplers
    (".downsamplers.0.conv.weight", ".downsamplers.0.op.conv.weight"),
    (".downsamplers.0.conv.bias", ".downsamplers.0.op.conv.bias"),
    (".upsamplers.0.conv.weight", ".upsamplers.0.conv.conv.weight"),
    (".upsamplers.0.conv.bias", ".upsamplers.0.conv.conv.bias"),
)


def _remap_legacy_unet_keys(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for key, tensor in state_dict.items():
        for old, new in _LEGACY_KEY_RULES:
            key = key.replace(old, new)
        out[key] = tensor
    return out


def _load_legacy_unet_state(model: torch.nn.Module, state: Dict[str, torch.Tensor], strict_shapes: bool = True) -> None:
    """Load a renamed checkpoint: names may differ (remapped), tensor shapes may not."""
    state = _remap_legacy_unet_keys(state)
    own = model.state_dict()
    usable = {k: v for k, v in state.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
    bad_shape = [f"{k}: ckpt={tuple(v.shape)} model={tuple(own[k].shape)}"
                 for k, v in state.items() if k in own and tuple(v.shape) != tuple(own[k].shape)]
    unexpected = [k for k in state if k not in own]
    missing = [k for k in own if k not in usable]
    if strict_shapes and bad_shape:
        shown = "\n".join(bad_shape[:20])
        more = f"\n... and {len(bad_shape) - 20} more" if len(bad_shape) > 20 else ""
        raise RuntimeError("Legacy load failed due to shape mismatches:\n" + shown + more)
    model.load_state_dict(usable, strict=False)
    if strict_shapes and (missing or unexpected):
        parts = ([f"missing={len(missing)}"] if missing else []) + ([f"unexpected={len(unexpected)}"] if unexpected else [])
        raise RuntimeError("Legacy load key mismatch after conversion (" + ", ".join(parts) + "). "
                           "Architecture/config likely differs from the source checkpoint.")


def _read_checkpoint(path: str, device: torch.device) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        try:
            from safetensors.torch import load_file
        except Exception as exc:  # the dependency is optional in the reference too
            raise RuntimeError("Loading .safetensors checkpoints requires `safetensors` package.") from exc
        return load_file(path, device=str(device))
    try:
        payload = torch.load(path, map_location=device, weights_only=True)
    except TypeError:  # very old torch
        payload = torch.load(path, map_location=device)
    return payload["model"] if isinstance(payload, dict) and "model" in payload else payload


def build_diffusion_model(cfg: dict, device: torch.device, ckpt_path=None, set_eval: bool = True):
    """cfg (`train_config.json` layout) -> B200 denoiser, optionally with `{flow,diff}_{best,last}.pt` weights."""
    training_cfg = cfg["training"]
    unet_cfg = cfg["model"].get("unet", {})
    mode = resolve_conditioning_mode(training_cfg.get("conditioning") or cfg["model"].get("conditioning"))
    channels = int(training_cfg.get("channels", unet_cfg.get("out_channels", 1)))
    model = DiffusionUNetFactory().build(unet_cfg, mode, channels).to(device)
    if ckpt_path is not None:
        state = _read_checkpoint(str(ckpt_path), device)
        strict = bool(unet_cfg.get("legacy_strict_shapes", True))
        if bool(unet_cfg.get("load_legacy", False)):
            _load_legacy_unet_state(model, state, strict_shapes=strict)
        else:
            try:
                model.load_state_dict(state)
            except RuntimeError:  # external diffusers-style checkpoint: same shapes, other names
                _load_legacy_unet_state(model, state, strict_shapes=strict)
    if set_eval:
        model.eval()
    return model


def encode_diffusion_batch(scheduler, targets: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
    """Forward-noise `targets` to `timesteps` with the scheduler's `add_noise`."""
    return scheduler.add_noise(targets, torch.randn_like(targets), timesteps)


def decode_diffusion_batch(model, training_cfg: dict, model_cfg: dict, device: torch.device,
                           batch_shape: Tuple[int, ...], conditioning_batch: Optional[torch.Tensor] = None,
                           timing: Optional[dict] = None, num_inference_steps: Optional[int] = None,
                           start_step: Optional[int] = None, last_n_steps: Optional[int] = None,
                           reference_batch: Optional[torch.Tensor] = None, init_from_reference: bool = False,
                           scheduler_override: Optional[str] = None, *,
                           init_sample: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One batch of samples: scheduler from the config (or `--scheduler` override), optional partial trajectory.

    `init_sample` (keyword-only, not in the reference signature): explicit initial noise, so a caller that shards a
    global batch over GPUs gets results independent of the GPU count; ignored when `init_from_reference` applies."""
    sched_cfg = dict(model_cfg.get("scheduler", {}))
    override = resolve_scheduler_override(scheduler_override)
    if override is not None:
        sched_cfg["name"] = override["name"]
        params = dict(sched_cfg.get("params", {}))
        params.update(override.get("params", {}))
        sched_cfg["params"] = params
    scheduler, steps = build_scheduler(sched_cfg, training_cfg)
    if num_inference_steps is not None:
        steps = int(num_inference_steps)
    scheduler.set_timesteps(steps)
    selected = scheduler.timesteps
    if start_step is not None:
        selected = selected[selected <= int(start_step)]
    if last_n_steps is not None:
        selected = selected[-int(last_n_steps):]

    if init_from_reference and reference_batch is not None:
        if selected.numel() == 0:
            raise ValueError("No timesteps selected after applying start_step/last_n_steps.")
        if hasattr(scheduler, "add_noise"):
            t0 = selected[0].expand(reference_batch.size(0)).to(reference_batch.device)
            init_sample = scheduler.add_noise(reference_batch, torch.randn_like(reference_batch), t0).to(device)
        else:
            logging.warning("Requested init_from_reference but scheduler '%s' has no add_noise; falling back to random "
                            "init.", scheduler.__class__.__name__)
    mode = resolve_conditioning_mode(training_cfg.get("conditioning") or model_cfg.get("conditioning"))
    return sample_with_scheduler(model, scheduler, steps, batch_shape, device, conditioning_mode=mode,
                                 conditioning_batch=conditioning_batch, latent_norm=training_cfg.get("latent_norm"),
                                 timing=timing, start_step=start_step, last_n_steps=last_n_steps,
                                 init_sample=init_sample)


def warn_attention_conditioning_shape(conditioning_batch: Optional[torch.Tensor], model_cfg: dict) -> bool:
    """True (and a warning) when attention conditioning does not have `unet.cross_attention_dim` channels."""
    if conditioning_batch is None or conditioning_batch.dim() < 2:
        return False
    unet_cfg = model_cfg.get("unet", {}) if isinstance(model_cfg, dict) else {}
    expected = unet_cfg.get("cross_attention_dim")
    if expected is None:
        return False
    actual = int(conditioning_batch.shape[1])
    if actual != int(expected):
        logging.warning("Attention conditioning has %d channels, but model unet.cross_attention_dim is %d. This often "
                        "means the evaluation split is pointing at pixel conditioning instead of the expected latent "
                        "conditioning.", actual, int(expected))
        return True
    return False
