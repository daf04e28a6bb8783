"""Timing of the per-head RMSNorm + RoPE kernels of the training step at the configs[3] shape (8192 tokens, 16 heads of 72):
forward qknorm_rope (q and k in one launch) and the backward headnorm_rope_bwd (one launch per segment).  20 launches per
CUDA graph, rotating buffers > L2.  python scripts/headnorm_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf16 = torch.bfloat16
M, heads, d, L = 8192, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 72, 256
H = heads * d
NB = 8
raw = [torch.randn(M, 3 * H, device=dev).to(bf16) for _ in range(NB)]
out = [torch.empty(M, 3 * H, device=dev, dtype=bf16) for _ in range(NB)]
gg = [torch.randn(M, 3 * H, device=dev).to(bf16) for _ in range(NB)]
qw, kw = torch.rand(d, device=dev) + 0.5, torch.rand(d, device=dev) + 0.5
ang = torch.rand(L, d // 2, device=dev) * 6.28
pos = torch.stack([ang.cos(), ang.sin()], -1).contiguous()
dq = torch.zeros(d, device=dev)


def timeit(fn, iters=24):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / iters * 1e3


f = timeit(lambda i: ops.qknorm_rope_to(raw[i % NB], out[i % NB], qw, kw, pos, heads, d, L))
b = timeit(lambda i: ops.headnorm_rope_bwd_(gg[i % NB], raw[i % NB], 0, qw, pos, dq, heads, d, L))
print(f"head_dim {d}: qknorm_rope forward (q + k, {4 * M * H * 2 / 1e6:.0f} MB) {f:6.1f} us = {4 * M * H * 2 / f / 1e3:5.0f} GB/s | "
      f"headnorm_rope_bwd (one segment, {3 * M * H * 2 / 1e6:.0f} MB) {b:6.1f} us = {3 * M * H * 2 / b / 1e3:5.0f} GB/s")
