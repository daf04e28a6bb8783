"""Input / output contract on either side of the sampling path.

Mirrors (reference paths):
  src/data/dataset/randn.py:38-90      RandomNDataset / ClassLabelRandomNDataset (per-sample seeded CPU noise)
  src/models/conditioner/class_label.py:4-13   LabelConditioner (labels -> int64, null label = num_classes)
  src/models/autoencoder/pixel.py:4-12, base.py:32-34   PixelAE (identity scale/shift) and fp2uint8
  src/lightning_data.py:142-144        DistributedSampler(shuffle=False): rank r takes indices r, r+W, ...
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops


class ClassLabelRandomNDataset:
    def __init__(self, latent_shape=(4, 64, 64), num_classes=1000, conditions=None, seeds=None,
                 max_num_instances=50000, num_samples_per_instance=-1):
        if conditions is None:
            conditions = list(range(num_classes))
        elif isinstance(conditions, int):
            conditions = list(range(conditions))
        self.conditions = list(conditions)
        self.num_conditons = len(self.conditions)
        self.seeds = seeds
        if num_samples_per_instance > 0:
            max_num_instances = num_samples_per_instance * self.num_conditons
        if seeds is not None:
            self.num_seeds = len(seeds)
        else:
            self.num_seeds = (max_num_instances + self.num_conditons - 1) // self.num_conditons
        self.max_num_instances = self.num_seeds * self.num_conditons
        self.latent_shape = tuple(latent_shape)

    def __len__(self):
        return self.max_num_instances

    def item(self, idx: int, seed: Optional[int] = None):
        condition = self.conditions[idx // self.num_seeds]
        if self.seeds is not None:
            seed = self.seeds[idx % self.num_seeds]
        elif seed is None:
            seed = idx % self.num_seeds   # deterministic stand-in for the reference's random.randint
        generator = torch.Generator().manual_seed(seed)
        latent = torch.randn(self.latent_shape, generator=generator, dtype=torch.float32)
        return latent, condition, dict(filename=f"{condition}_{seed}", seed=seed, condition=condition)

    __getitem__ = item


def rank_indices(n: int, rank: int, world_size: int) -> List[int]:
    """Index shard of DistributedSampler(shuffle=False, drop_last=False): pad by wrapping, then stride by world."""
    idx = list(range(n))
    total = (n + world_size - 1) // world_size * world_size
    idx += idx[: total - n]
    return idx[rank:total:world_size]


def seeded_noise(seeds: Sequence[int], shape: Tuple[int, ...], pin: bool = True) -> torch.Tensor:
    """One CPU generator per sample (randn.py:74-75) into a pinned host batch."""
    out = torch.empty((len(seeds),) + tuple(shape), dtype=torch.float32, pin_memory=pin and torch.cuda.is_available())
    for i, s in enumerate(seeds):
        out[i] = torch.randn(shape, generator=torch.Generator().manual_seed(int(s)), dtype=torch.float32)
    return out


class LabelConditioner(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.null_condition = num_classes

    @torch.no_grad()
    def __call__(self, y, metadata: dict = {}, device="cuda"):
        condition = torch.as_tensor(y).long().to(device)
        uncondition = torch.full((len(y),), self.null_condition, dtype=torch.long, device=device)
        return condition, uncondition


class PixelAE(nn.Module):
    def __init__(self, scale=1.0, shift=0.0):
        super().__init__()
        self.scale, self.shift = scale, shift

    def encode(self, x):
        return x / self.scale + self.shift

    def decode(self, x):
        return (x - self.shift) * self.scale


def fp2uint8(x: torch.Tensor) -> torch.Tensor:
    return ops.fp2uint8(x)
