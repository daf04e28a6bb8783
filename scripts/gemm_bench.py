"""Micro-benchmark of the tcgen05 GEMM on the DeCo-XL shapes, for every (cta_group, epilogue style, tile_n) variant.
CUDA events, 3 warm-up + 10 timed launches per case, buffers > L2.  python scripts/gemm_bench.py [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import _lib, ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
H, F, L = 1152, 3072, 256
dev = torch.device("cuda:0")
bf16 = torch.bfloat16
lib = _lib.load()


def rnd(*s, dt=bf16, sc=1.0):
    return (torch.randn(*s, device=dev) * sc).to(dt)


cases = {
    "qkv      N=3456 K=1152 bias->bf16": dict(N=3 * H, K=H, epi=ops.EPI_BIAS),
    "proj     N=1152 K=1152 gate+res  ": dict(N=H, K=H, epi=ops.EPI_GATE_RESIDUAL),
    "w13      N=6144 K=1152 swiglu    ": dict(N=2 * F, K=H, epi=ops.EPI_SWIGLU),
    "w2       N=1152 K=3072 gate+res  ": dict(N=H, K=F, epi=ops.EPI_GATE_RESIDUAL),
    "cond     N=8192 K=1152 bias->bf16": dict(N=8192, K=H, epi=ops.EPI_BIAS),
}
variants = [(1, 0), (1, 1), (2, 0), (2, 1)]
for name, c in cases.items():
    N, K, epi = c["N"], c["K"], c["epi"]
    a, w = rnd(M, K), rnd(N, K, sc=K ** -0.5)
    bias = rnd(N, dt=torch.float32) if epi != ops.EPI_SWIGLU else None
    resid = rnd(M, N, dt=torch.float32) if epi == ops.EPI_GATE_RESIDUAL else None
    gate = rnd(M // L, N) if epi == ops.EPI_GATE_RESIDUAL else None
    out = torch.empty((M, N // 2 if epi == ops.EPI_SWIGLU else N), device=dev,
                      dtype=torch.float32 if epi == ops.EPI_GATE_RESIDUAL else bf16)
    tiles = [t for t in (128, 192, 256) if N % t == 0]
    for cg, st in variants:
        for tn in tiles:
            lib.deco_gemm_set_tuning(cg, st)
            try:
                for _ in range(3):
                    ops.gemm(a, w, bias, epi, out=out, resid=resid, gate=gate, rows_per_gate=L, tile_n=tn)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    ops.gemm(a, w, bias, epi, out=out, resid=resid, gate=gate, rows_per_gate=L, tile_n=tn)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(f"{name} cta_group={cg} staged={st} BN={tn}: {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"{name} cta_group={cg} staged={st} BN={tn}: FAILED {ex}", flush=True)
                raise
lib.deco_gemm_set_tuning(-1, -1)
