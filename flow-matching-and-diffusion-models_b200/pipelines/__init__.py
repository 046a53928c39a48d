from .schedulers import (DDIMScheduler, DDPMScheduler, DPMSolverMultistepScheduler,
                         FlowMatchEulerDiscreteScheduler)
from .utils import (SCHEDULER_REGISTRY, build_scheduler, resolve_conditioning_mode, resolve_scheduler_override,
                    sample_with_scheduler)

__all__ = ["DDIMScheduler", "DDPMScheduler", "DPMSolverMultistepScheduler", "FlowMatchEulerDiscreteScheduler", "SCHEDULER_REGISTRY",
           "build_scheduler", "resolve_conditioning_mode", "resolve_scheduler_override", "sample_with_scheduler"]
