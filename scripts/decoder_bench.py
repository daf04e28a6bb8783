"""Pixel decoder A/B at the bench shape: legacy mma.sync kernel vs the tcgen05 kernel (plain and with the fused sampler step),
each with its cond_embed GEMM.  python scripts/decoder_bench.py [rows] [res]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import XL, _time_kernel  # noqa: E402
from deco_b200 import PixNerDiT, ops  # noqa: E402
from deco_b200.utils import randomize_  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
res = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda")
with torch.device("meta"):
    net = PixNerDiT(**XL)
net = randomize_(net.to_empty(device=dev), seed=0).eval()
P = net.prepare(dev)
L = (res // 16) ** 2
M = rows * L
s = torch.randn(M, 1152, device=dev).to(torch.bfloat16)
x = torch.randn(rows, 3, res, res, device=dev)
xh = x[: rows // 2].contiguous()
ycond = ops.gemm(s, P["wcond"], P["bcond"], ops.EPI_BIAS)
ysilu = ops.gemm(s, P["wcond"], P["bcond"], ops.EPI_BIAS_SILU)
it = 5 if rows >= 256 else 20
t = {}
t["cond_embed GEMM (bias)"] = _time_kernel(lambda i: ops.gemm(s, P["wcond"], P["bcond"], ops.EPI_BIAS, out=ycond), it, torch)
t["cond_embed GEMM (bias + SiLU)"] = _time_kernel(lambda i: ops.gemm(s, P["wcond"], P["bcond"], ops.EPI_BIAS_SILU, out=ysilu), it, torch)
t["pixel_decoder_kernel (mma.sync)"] = _time_kernel(lambda i: ops.pixel_decoder(x, ycond, P["blob"], P["postab"], 16, 32, 3), it, torch)
t["pixel_decoder_tc_kernel (plain)"] = _time_kernel(lambda i: ops.pixel_decoder_tc(x, ysilu, P["blob_tc"], 16, 32, 3), it, torch)
xo = torch.empty_like(xh)
t["pixel_decoder_tc_kernel (fused CFG step)"] = _time_kernel(
    lambda i: ops.pixel_decoder_tc_step(xh, ysilu, P["blob_tc"], 16, 32, 3, g=3.2, dt=0.01, x_out=xo), it, torch)
npx = rows * res * res
for k, v in t.items():
    extra = ""
    if "decoder" in k:
        extra = f"  {npx / v / 1e6:8.1f} Mpixel/ms  {37.2e3 * npx / (v * 1e-3) / 1e12:7.1f} TFLOP/s (37.2 kFLOP/pixel)"
    print(f"{k:45s} {v:8.3f} ms{extra}")
print(f"rows {rows}, {res}px, tokens {M}")
