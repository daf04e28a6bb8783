import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import ops
dev = torch.device("cuda")
B, heads, d, L = 32, 16, 72, 256
H, M = heads * d, B * L
qkv = torch.randn(M, 3 * H, device=dev).to(torch.bfloat16)
do = torch.randn(M, H, device=dev).to(torch.bfloat16)
o, lse = ops.attention_lse(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d)
dqkv = torch.zeros_like(qkv)
for _ in range(3):
    ops.attention_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, do, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], B, heads, d, lse=lse)
torch.cuda.synchronize()
