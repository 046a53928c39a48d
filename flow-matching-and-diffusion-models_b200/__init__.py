"""B200-native sampling hot path of tomn681/Flow-Matching-and-Diffusion-Models.

Sub-modules mirror the reference's `src/` layout for the path only:
  nn/ (ops, blocks)            <- src/nn
  models/unet, models/generators/diffusionfactory.py  <- src/models
  pipelines/utils.py           <- src/pipelines/utils.py (schedulers + sample_with_scheduler)
  utils/model_utils/diffusion_utils.py <- src/utils/model_utils/diffusion_utils.py
  run_model.py                 <- src/run_model.py (--mode sample / evaluate surface)
  csrc/                        hand-written sm_100a kernels + the C-ABI (include/fmdm_b200.h)
"""
__version__ = "0.1.0"
