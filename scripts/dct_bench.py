"""DCT + FM loss kernel at the BASELINE configs[3] size (32 images) and at 256 images: HBM fraction, fwd and fwd+bwd."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _time_kernel, measured_peaks  # noqa: E402
from deco_b200 import ops  # noqa: E402
from deco_b200.training import build_freq_weight  # noqa: E402

dev = torch.device("cuda")
peak = measured_peaks()["hbm_gbs"]
fw = build_freq_weight(85).reshape(3, 8, 8).contiguous().to(dev)
res = 256
for Bt, nset, iters in ((32, 8, 24), (64, 4, 12), (256, 2, 6)):
    outs = [torch.randn(Bt, 3, res, res, device=dev) for _ in range(nset)]
    vts = [torch.randn(Bt, 3, res, res, device=dev) for _ in range(nset)]
    for name, grad, nb in (("fwd+bwd", True, 12.0), ("fwd", False, 8.0)):
        ms = _time_kernel(lambda i: ops.dct_fm_loss(outs[i % nset], vts[i % nset], fw, 1.0, want_loss=True, want_grad=grad), iters, torch)
        gbs = nb * Bt * 3 * res * res / (ms * 1e-3) / 1e9
        print(f"{Bt:4d} images {name:8s} {ms * 1e3:8.2f} us  {gbs:7.1f} GB/s  {gbs / peak:5.3f} of the measured HBM peak")
