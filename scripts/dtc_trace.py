"""Clock-stamp trace of csrc/decoder_tc.cu (CTA 0, slot 0): build the instrumented library (-DDECO_DTC_TRACE), run one launch
at the bench shape and print the per-stage latencies.  DECO_B200_LIB=build/libdeco_b200_trace.so python scripts/dtc_trace.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import XL  # noqa: E402
from deco_b200 import PixNerDiT, _lib, ops  # noqa: E402
from deco_b200.utils import randomize_  # noqa: E402

rows, res = 64, 256
dev = torch.device("cuda")
with torch.device("meta"):
    net = PixNerDiT(**XL)
net = randomize_(net.to_empty(device=dev), seed=0).eval()
P = net.prepare(dev)
M = rows * (res // 16) ** 2
s = torch.randn(M, 1152, device=dev).to(torch.bfloat16)
x = torch.randn(rows, 3, res, res, device=dev)
ysilu = ops.gemm(s, P["wcond"], P["bcond"], ops.EPI_BIAS_SILU)
for _ in range(2):
    ops.pixel_decoder_tc(x, ysilu, P["blob_tc"], 16, 32, 3)
torch.cuda.synchronize()
buf = np.zeros(2 * 4096, dtype=np.int64)
lib = _lib.load()
lib.deco_dtc_trace_copy.argtypes = [ctypes.c_void_p]
rc = lib.deco_dtc_trace_copy(buf.ctypes.data)
assert rc == 0, rc
for region, name in ((0, "epilogue thread 0 of slot 0"), (1, "issuer of slot 0")):
    r = buf[region * 4096:(region + 1) * 4096].reshape(-1, 2)
    r = r[r[:, 0] > 0]
    t0 = r[0, 0]
    print(f"--- {name}: {len(r)} stamps")
    prev = t0
    for i, (t, c) in enumerate(r[:120]):
        print(f"{int(c):4d}  t={int(t - t0):8d}  +{int(t - prev):6d}")
        prev = t
