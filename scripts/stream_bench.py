"""Micro-benchmark of the FE_STREAM epilogue: the proj-shaped GEMM at shrinking K isolates the residual streaming rate."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deco_b200 import ops

dev = torch.device("cuda")
M, N, L = 131072, 1152, 256
bf = torch.bfloat16
resid = torch.randn(M, N, device=dev)
gate = torch.randn(M // L, N, device=dev).to(bf)
nsc = torch.randn(M // L, N, device=dev).to(bf)
nw = torch.ones(N, device=dev)
bias = torch.zeros(N, device=dev)
xg = torch.empty(M, N, device=dev, dtype=bf)


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for K in (64, 256, 576, 1152, 3072):
    a = torch.randn(M, K, device=dev).to(bf)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(bf)
    ssq = torch.empty(ops.gemm_stream_parts(N, K), M, device=dev)
    full = t(lambda: ops.gemm_stream(a, w, bias, resid, resid=resid, gate=gate, rows_per_image=L, next_w=nw, next_scale=nsc, xg=xg, ssq=ssq))
    noxg = t(lambda: ops.gemm_stream(a, w, bias, resid, resid=resid, gate=gate, rows_per_image=L, ssq=ssq))
    nores = t(lambda: ops.gemm_stream(a, w, bias, resid, rows_per_image=L, ssq=ssq))
    byt = M * N * (4 + 4 + 2) + M * K * 2
    print(f"K={K:5d}  full {full*1e3:7.1f} us ({byt/full/1e6:6.0f} GB/s)   no-xg {noxg*1e3:7.1f} us   no-resid-no-xg {nores*1e3:7.1f} us   "
          f"MMA-only bound {2*M*N*K/1.5e15*1e6:6.1f} us")
