"""Where the pipelined attention-backward kernels (csrc/attention_bwd_tc.cu) spend their clocks: per-role totals and the
timeline of CTA 0.  scripts/abt_trace.sh builds deco_b200/_C/libdeco_trace.so with -DABT_TRACE (here, nvcc cross-compiles);
on the GPU: DECO_B200_LIB=$PWD/deco_b200/_C/libdeco_trace.so python scripts/abt_trace.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deco_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda")
B, heads, d, L = 32, 16, 72, 256
H, M = heads * d, B * L
qkv = torch.randn(M, 3 * H, device=dev).to(torch.bfloat16)
do = torch.randn(M, H, device=dev).to(torch.bfloat16)
o, lse = ops.attention_lse(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d)
dqkv = torch.zeros_like(qkv)
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * 32)()
run = lambda: ops.attention_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, do, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:],
                                B, heads, d, lse=lse)
for _ in range(3):
    run()
torch.cuda.synchronize()
lib.deco_abt_trace_read(buf, 1)
N = 10
for _ in range(N):
    run()
torch.cuda.synchronize()
lib.deco_abt_trace_read(buf, 0)
names = ["S issuer: wait math_done(st-2)", "G issuer: wait math_done(st)", "G issuer: issue + commits", "S issuer: wait loads",
         "S issuer: issue + commits", "math: wait s_full", "math: row math (+ wait g_done)", "math: wait acc_full", "math: whole loop",
         "math: accumulator read-out", "S issuer: wait row tiles (part of wait loads)", "producer: wait row_empty",
         "producer: wait ld_empty", "producer: TMA issue (column block)", "producer: whole loop"]
ctas = 148
for p in range(2):
    print(f"pass {p} ({'dQ' if p == 0 else 'dK/dV'}): clocks per CTA per launch")
    for i, nm in enumerate(names):
        print(f"   {nm:46s} {buf[p * 16 + i] / N / ctas:10.0f}")

# timeline of CTA 0 (last launch): clocks relative to the first recorded event
tl = (ctypes.c_longlong * (2 * 6 * 64 * 4))()
lib.deco_abt_timeline_read(tl)
import numpy as np
T = np.array(tl, dtype=np.int64).reshape(2, 6, 64, 4)
for p in range(2):
    t0 = T[p][T[p] > 0].min()
    R = np.where(T[p] > 0, T[p] - t0, -1)
    print(f"pass {p}: step | S-issuer: enter, loads ready, math_done(st-2) ready, issued | G-issuer: enter, math_done ready, issued | "
          f"math: enter, s_full ready, done | producer: enter, stage free, issued")
    for st in range(20):
        print(f"  {st:2d} | {R[0][st][0]:6d} {R[0][st][1]:6d} {R[0][st][2]:6d} {R[0][st][3]:6d} | {R[2][st][0]:6d} {R[2][st][1]:6d} {R[2][st][3]:6d} | "
              f"{R[3][st][0]:6d} {R[3][st][1]:6d} {R[3][st][2]:6d} | {R[4][st][0]:6d} {R[4][st][1]:6d} {R[4][st][2]:6d}")
    print("  read-out of item n: enter (accumulators complete), TMEM loaded, scratch written + pair barrier, done")
    for n in range(5):
        print(f"  {n:2d} | {R[5][n][0]:6d} {R[5][n][1]:6d} {R[5][n][2]:6d} {R[5][n][3]:6d}")
