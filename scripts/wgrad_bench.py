"""Timing of the weight-gradient (TN, MN-major operands) GEMMs at the configs[3] shapes: dW [N_w, K_w] = dY^T [N_w, T] . X [T, K_w],
T = 8192 tokens.  16 launches per CUDA graph, rotating buffers.  python scripts/wgrad_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf16 = torch.bfloat16
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
NB = 4


def timeit(fn, iters=16):
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


tot = 0.0
for name, n, k in (("w2   [1152 x 3072]", 1152, 3072), ("w13  [6144 x 1152]", 6144, 1152), ("proj [1152 x 1152]", 1152, 1152),
                   ("qkv  [3456 x 1152]", 3456, 1152), ("cond [8192 x 1152]", 8192, 1152)):
    dy = [torch.randn(T, n, device=dev).to(bf16) for _ in range(NB)]
    x = [torch.randn(T, k, device=dev).to(bf16) for _ in range(NB)]
    out = [torch.empty(n, k, device=dev) for _ in range(NB)]
    us = timeit(lambda i: ops.gemm_tn(dy[i % NB], x[i % NB], out=out[i % NB]))
    fl = 2.0 * T * n * k
    print(f"wgrad {name}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
    if not name.startswith("cond"):
        tot += us
print(f"four block wgrads: {tot:.1f} us")
