"""Attention at the bench shape (XL/16 256 px, 512 CFG rows, 16 heads x 72) with the dense head pitch (72) and the padded one
(80): time per launch; under ncu the DRAM bytes.  DECO_ATTN_L2PROMO={0,64,128,256} selects the TMA L2 promotion."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, heads, d, L = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 16, 72, 256
g = torch.Generator(device="cuda").manual_seed(0)
base = torch.randn((B * L, 3, heads, d), device=dev, generator=g).to(torch.bfloat16)
outs = {}
for pitch in (72, 80):
    buf = torch.zeros((B * L, 3, heads, pitch), device=dev, dtype=torch.bfloat16)
    buf[..., :d] = base
    qkv = buf.view(B * L, 3 * heads * pitch)
    Hp = heads * pitch
    out = torch.empty((B * L, heads * d), device=dev, dtype=torch.bfloat16)

    def run():
        ops.attention(qkv[:, :Hp], qkv[:, Hp:2 * Hp], qkv[:, 2 * Hp:], B, heads, d, out=out, head_pitch=0 if pitch == d else pitch)
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
    outs[pitch] = out.clone()
    print(f"promo={os.environ.get('DECO_ATTN_L2PROMO', '128'):>3s} head pitch {pitch}: {ms:7.4f} ms  "
          f"{4.0 * B * heads * L * L * d / ms / 1e9:7.1f} TFLOP/s", flush=True)
print("outputs equal:", bool(torch.equal(outs[72], outs[80])))
