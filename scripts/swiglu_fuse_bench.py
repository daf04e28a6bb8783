"""A/B of the SwiGLU training epilogues at the configs[3] shapes (8192 tokens): GEMM + stand-alone SwiGLU kernel vs the
fused GEMM.  20 launches captured in a CUDA graph, rotating buffers > L2.  python scripts/swiglu_fuse_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf16 = torch.bfloat16
M, H, Fp = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 1152, 3072
NB = 8


def rnd(*s, sc=1.0):
    return (torch.randn(*s, device=dev) * sc).to(bf16)


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


h2 = [rnd(M, H) for _ in range(NB)]
w13 = [rnd(2 * Fp, H, sc=H ** -0.5) for _ in range(NB)]
y13 = [torch.empty(M, 2 * Fp, device=dev, dtype=bf16) for _ in range(NB)]
u = [torch.empty(M, Fp, device=dev, dtype=bf16) for _ in range(NB)]
da = [rnd(M, H) for _ in range(NB)]
w2T = [rnd(Fp, H, sc=H ** -0.5) for _ in range(NB)]
du = [torch.empty(M, Fp, device=dev, dtype=bf16) for _ in range(NB)]
dy = [torch.empty(M, 2 * Fp, device=dev, dtype=bf16) for _ in range(NB)]
for i in range(NB):
    ops.gemm(h2[i], w13[i], None, ops.EPI_BIAS, out=y13[i])

print(f"M = {M} tokens, hidden {H}, FFN {Fp}")
a = timeit(lambda i: ops.gemm(h2[i % NB], w13[i % NB], None, ops.EPI_BIAS, out=y13[i % NB]))
b = timeit(lambda i: ops.swiglu_fwd(y13[i % NB], out=u[i % NB]))
c = timeit(lambda i: ops.gemm(h2[i % NB], w13[i % NB], None, ops.EPI_SWIGLU_DUAL, out=u[i % NB], aux=y13[i % NB]))
d = timeit(lambda i: ops.gemm(h2[i % NB], w13[i % NB], None, ops.EPI_SWIGLU, out=u[i % NB]))
print(f"forward : GEMM -> y13 {a:7.1f} us + swiglu_fwd {b:6.1f} us = {a + b:7.1f} us | fused dual output {c:7.1f} us | (u only {d:7.1f} us)")
a = timeit(lambda i: ops.gemm(da[i % NB], w2T[i % NB], None, ops.EPI_BIAS, out=du[i % NB]))
b = timeit(lambda i: ops.swiglu_bwd(y13[i % NB], du[i % NB]))
c = timeit(lambda i: ops.gemm(da[i % NB], w2T[i % NB], None, ops.EPI_SWIGLU_BWD, out=dy[i % NB], aux=y13[i % NB]))
print(f"backward: GEMM -> du  {a:7.1f} us + swiglu_bwd {b:6.1f} us = {a + b:7.1f} us | fused dy13 epilogue {c:7.1f} us")
fl = 2.0 * M * 2 * Fp * H
print(f"(w13 GEMM = {fl / 1e9:.1f} GFLOP, du GEMM = {fl / 2e9:.1f} GFLOP)")
