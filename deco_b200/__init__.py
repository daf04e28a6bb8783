"""deco_b200 -- B200 (sm_100a) implementation of DeCo's denoise / sample / DCT-loss hot path.

Public surface mirrors the reference classes (see DESIGN.md for the file:line map):
    PixNerDiT, EulerSampler, HeunSampler, AdamLMSampler, LinearScheduler, REPATrainer,
    simple_guidance_fn, ode_step_fn
"""
from .denoiser import PixNerDiT  # noqa: F401
from .sampling import (AdamLMSampler, BaseSampler, EulerSampler, HeunSampler, ode_step_fn,  # noqa: F401
                       shift_respace_fn, simple_guidance_fn)
from .scheduling import BaseScheduler, LinearScheduler  # noqa: F401
from .optim import FusedAdamWEMA  # noqa: F401
from .training import BaseTrainer, REPATrainer  # noqa: F401

__version__ = "0.1.0"
