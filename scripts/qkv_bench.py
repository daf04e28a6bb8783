"""Micro-benchmark of the FE_NORM_QKV GEMM (csrc/gemm_fused.cu) at the XL/16 shape: q/k-norm + axial RoPE in the epilogue,
head pitch 80.  DECO_QKV_EPI_SETS=1 / 2 selects one or two epilogue warp sets (run the script once per setting: the choice is
read once per process).  python scripts/qkv_bench.py [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf = torch.bfloat16
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L, H, heads, d = 256, 1152, 16, 72
M = rows * L
NB = 2
a = [torch.randn(M, H, device=dev).to(bf) for _ in range(NB)]
w = (torch.randn(3 * H, H, device=dev) * H ** -0.5).to(bf)
out = [torch.empty(M, 3 * heads * 80, device=dev, dtype=bf) for _ in range(NB)]
ssq = torch.rand(6, M, device=dev) * H / 6
shw = torch.randn(rows, 3 * H, device=dev)
qn, kn = torch.rand(d, device=dev) + 0.5, torch.rand(d, device=dev) + 0.5
ang = torch.rand(L, d // 2, device=dev) * 6.28
rope = torch.stack([ang.cos(), ang.sin()], -1).contiguous()


def run(i):
    ops.gemm_norm_qkv(a[i % NB], w, out[i % NB], L, heads, d, seg_w=(qn, kn, None), rope_mask=3, rope=rope,
                      rope_tokens_per_row=16, ssq=ssq, norm_hidden=H, shw=shw, out_head_pitch=80)


for i in range(3):
    run(i)
torch.cuda.synchronize()
iters = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(iters):
    run(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"FE_NORM_QKV rows={rows} (M={M}) sets={os.environ.get('DECO_QKV_EPI_SETS', '2')}: {ms * 1e3:8.1f} us  {2.0 * M * 3 * H * H / ms / 1e9:7.1f} TFLOP/s")
