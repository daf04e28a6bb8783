"""Where the pipelined attention-backward kernels spend their clocks (build csrc/attention_bwd_tc.cu with -DABT_TRACE into
a scratch library first: see the command in the comment below).  python scripts/abt_trace.py
  nvcc ... -DABT_TRACE  (scripts/abt_trace.sh builds deco_b200/_C/libdeco_trace.so and runs this with DECO_B200_LIB set)"""
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
