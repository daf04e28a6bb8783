"""Flow-matching schedulers with the reference's interface.

Mirrors src/diffusion/base/scheduling.py:4-32 (BaseScheduler) and src/diffusion/flow_matching/scheduling.py:6-14
(LinearScheduler: alpha = t, sigma = 1 - t, dalpha = 1, dsigma = -1, all returned as [B,1,1,1] views).
GVPScheduler (:17-28) and ConstScheduler (:30-32) serve as `w_scheduler` / scheduler of the SDE step functions; VPBeta
(ddpm tables) is not used by any flow-matching config and is out of scope.
"""
import math

import torch
from torch import Tensor


class BaseScheduler:
    def alpha(self, t) -> Tensor: ...
    def sigma(self, t) -> Tensor: ...
    def dalpha(self, t) -> Tensor: ...
    def dsigma(self, t) -> Tensor: ...

    def dalpha_over_alpha(self, t) -> Tensor:
        return self.dalpha(t) / self.alpha(t)

    def dsigma_mul_sigma(self, t) -> Tensor:
        return self.dsigma(t) * self.sigma(t)

    def drift_coefficient(self, t):
        return self.dalpha(t) / (self.alpha(t) + 1e-6)

    def diffuse_coefficient(self, t):
        alpha, sigma = self.alpha(t), self.sigma(t)
        return self.dsigma(t) * sigma - self.dalpha(t) / (alpha + 1e-6) * sigma ** 2

    def w(self, t):
        return self.sigma(t)


class LinearScheduler(BaseScheduler):
    def alpha(self, t) -> Tensor:
        return t.view(-1, 1, 1, 1)

    def sigma(self, t) -> Tensor:
        return (1 - t).view(-1, 1, 1, 1)

    def dalpha(self, t) -> Tensor:
        return torch.full_like(t, 1.0).view(-1, 1, 1, 1)

    def dsigma(self, t) -> Tensor:
        return torch.full_like(t, -1.0).view(-1, 1, 1, 1)


class GVPScheduler(BaseScheduler):
    """flow_matching/scheduling.py:17-28."""

    def alpha(self, t) -> Tensor:
        return torch.cos(t * (math.pi / 2)).view(-1, 1, 1, 1)

    def sigma(self, t) -> Tensor:
        return torch.sin(t * (math.pi / 2)).view(-1, 1, 1, 1)

    def dalpha(self, t) -> Tensor:
        return -torch.sin(t * (math.pi / 2)).view(-1, 1, 1, 1)

    def dsigma(self, t) -> Tensor:
        return torch.cos(t * (math.pi / 2)).view(-1, 1, 1, 1)

    def w(self, t):
        return torch.sin(t) ** 2


class ConstScheduler(BaseScheduler):
    """flow_matching/scheduling.py:30-32."""

    def w(self, t):
        return torch.ones(1, 1, 1, 1).to(t.device, t.dtype)
