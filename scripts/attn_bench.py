"""Attention micro-benchmark on the bench shapes (run on the GPU box): python scripts/attn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import ops  # noqa: E402
from deco_b200.denoiser import rope_cos_sin  # noqa: E402

dev = torch.device("cuda:0")
for name, B, heads, d, hw, Lt in [("XL/16 256px B'=512", 512, 16, 72, (16, 16), 0), ("XL/16 512px B'=128", 128, 16, 72, (32, 32), 0),
                                  ("L/16 256px B'=512", 512, 16, 64, (16, 16), 0), ("XXL t2i 512px B'=64", 64, 24, 64, (32, 32), 128)]:
    L, H = hw[0] * hw[1], heads * d
    qkv = torch.randn((B * L, 3 * H), device=dev, dtype=torch.bfloat16)
    kv2 = torch.randn((B * Lt, 2 * H), device=dev, dtype=torch.bfloat16) if Lt else None
    qw = torch.ones(d, device=dev)
    kw = torch.ones(d, device=dev)
    rope = rope_cos_sin(d, *hw).to(dev)
    out = torch.empty((B * L, H), device=dev, dtype=torch.bfloat16)
    for fused in (False,):
        def run():
            ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d, out=out,
                          k2=None if kv2 is None else kv2[:, :H], v2=None if kv2 is None else kv2[:, H:])
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 4.0 * B * heads * L * (L + Lt) * d
        print(f"{name:24s} fused_norm_rope={int(fused)}  {ms:7.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s (algorithmic, d unpadded)", flush=True)
