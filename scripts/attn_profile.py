"""One fused and one plain attention launch on the bench shape, for ncu: python scripts/attn_profile.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import ops  # noqa: E402
from deco_b200.denoiser import rope_cos_sin  # noqa: E402

dev = torch.device("cuda:0")
B, heads, d, hw = 512, 16, 72, (16, 16)
L, H = hw[0] * hw[1], heads * d
qkv = torch.randn((B * L, 3 * H), device=dev, dtype=torch.bfloat16)
qw = torch.ones(d, device=dev)
rope = rope_cos_sin(d, *hw).to(dev)
out = torch.empty((B * L, H), device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d, out=out)
torch.cuda.synchronize()
