"""Attention backward at the training shape (32 images, 16 heads x 72, L = 256) and at 512 px (L = 1024): tcgen05 kernels
(csrc/attention_bwd_tc.cu) vs the mma.sync pair (csrc/attention_bwd.cu).  python scripts/attn_bwd_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _time_kernel  # noqa: E402
from deco_b200 import ops  # noqa: E402

dev = torch.device("cuda")
for name, B, heads, d, L in [("XL/16 256px, 32 images", 32, 16, 72, 256), ("XL/16 512px, 8 images", 8, 16, 72, 1024),
                             ("L/16 256px, 32 images", 32, 16, 64, 256)]:
    H, M = heads * d, B * L
    qkv = torch.randn(M, 3 * H, device=dev).to(torch.bfloat16)
    do = torch.randn(M, H, device=dev).to(torch.bfloat16)
    o, lse = ops.attention_lse(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d)
    dqkv = torch.zeros_like(qkv)
    res = {}
    for mode in ("legacy", "tc"):
        ops.ATTN_BWD = mode
        res[mode] = _time_kernel(lambda i: ops.attention_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, do, dqkv[:, :H],
                                                             dqkv[:, H:2 * H], dqkv[:, 2 * H:], B, heads, d, lse=lse), 10, torch)
    fwd = _time_kernel(lambda i: ops.attention_lse(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d), 10, torch)
    fl = 10.0 * B * heads * L * L * d     # 5 products of 2 L^2 d flops
    print(f"{name:26s} forward {fwd * 1e3:7.1f} us | backward mma.sync {res['legacy'] * 1e3:7.1f} us, tcgen05 {res['tc'] * 1e3:7.1f} us "
          f"({fl / res['tc'] / 1e9:6.1f} TFLOP/s algorithmic), {res['legacy'] / res['tc']:.2f}x")
