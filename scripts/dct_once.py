"""Launch the fused DCT/FM loss kernel a few times at the BASELINE configs[3] shape (32 x 3 x 256 x 256, fp32), for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deco_b200 import LinearScheduler, REPATrainer

dev = torch.device("cuda")
tr = REPATrainer(scheduler=LinearScheduler()).to(dev)
for _ in range(4):
    out = torch.randn(32, 3, 256, 256, device=dev, requires_grad=True)
    v_t = torch.randn(32, 3, 256, 256, device=dev)
    tr.loss(out, v_t)["loss"].backward()
torch.cuda.synchronize()
print("ok")
