"""deco_b200 -- B200 (sm_100a) implementation of DeCo's denoise / sample / DCT-loss hot path.

Public surface mirrors the reference classes (see DESIGN.md for the file:line map):
    PixNerDiT, FlattenDiT (patch-linear baseline), EulerSampler, EulerSamplerJiT, HeunSampler, AdamLMSampler,
    LinearScheduler, GVPScheduler, ConstScheduler, REPATrainer, simple_guidance_fn, ode_step_fn, sde_*_step_fn
"""
from .denoiser import PixNerDiT  # noqa: F401
from .denoiser_baseline import FlattenDiT  # noqa: F401
from .sampling import (AdamLMSampler, BaseSampler, EulerSampler, EulerSamplerJiT, HeunSampler, ode_step_fn,  # noqa: F401
                       sde_mean_step_fn, sde_preserve_step_fn, sde_step_fn, shift_respace_fn, simple_guidance_fn)
from .scheduling import BaseScheduler, ConstScheduler, GVPScheduler, LinearScheduler  # noqa: F401
from .optim import FusedAdamWEMA  # noqa: F401
from .training import BaseTrainer, REPATrainer  # noqa: F401

__version__ = "0.1.0"
