"""Small host utilities: deterministic non-degenerate random initialisation and a GEMM timing probe."""
from __future__ import annotations

import math
from typing import List, Tuple

import torch


@torch.no_grad()
def randomize_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Seeded re-randomisation of EVERY parameter, on the parameter's own device.

    The reference's default init zeroes the decoder's output layer and adaLN layers
    (src/models/transformer/dit_c2i_DeCo.py:386-393), which makes the network output identically zero; benchmarks
    and smoke runs therefore use 'random-init weights' in this sense: fan-in scaled normals for matrices, small
    normals for biases, 1 + N(0, 0.1) for norm scales."""
    for idx, (name, p) in enumerate(sorted(module.named_parameters())):
        g = torch.Generator(device=p.device).manual_seed(seed * 100003 + idx)
        if p.dim() == 2 and "embedding_table" not in name:
            std = 1.0 / math.sqrt(p.shape[1])
            if "adaLN_modulation" in name:
                std *= 0.5
            p.copy_(torch.randn(p.shape, generator=g, device=p.device) * std)
        elif "embedding_table" in name:
            p.copy_(torch.randn(p.shape, generator=g, device=p.device) * 0.5)
        elif name.endswith(("norm.weight", "norm1.weight", "norm2.weight", "in_ln.weight")):
            p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=p.device))
        else:
            p.copy_(0.05 * torch.randn(p.shape, generator=g, device=p.device))
    return module


class GemmProbe:
    """CUDA-event pairs around every tcgen05 GEMM launch (same stream), for the live roofline figure in bench.py."""

    def __init__(self):
        self.records: List[Tuple[torch.cuda.Event, torch.cuda.Event, float]] = []

    def before(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def after(self, start, flops: float):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.records.append((start, e, flops))

    def summary(self):
        ms = sum(s.elapsed_time(e) for s, e, _ in self.records)
        fl = sum(f for _, _, f in self.records)
        return dict(launches=len(self.records), total_ms=ms, total_flops=fl,
                    avg_ms=ms / max(1, len(self.records)), tflops=(fl / (ms * 1e-3) / 1e12) if ms > 0 else 0.0)
