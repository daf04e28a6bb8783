"""One training step (forward + hand-written backward) of the text-to-image denoiser at the configs_t2i/sft_res512.yaml
architecture (DeCo-XXL: H 1536, 24 heads, 4 text + 16 joint blocks), 512 px, synthetic text-encoder states.
python scripts/t2i_train_time.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_t2i_module  # noqa: E402
from oracle import deco_oracle as O  # noqa: E402  (seeded weights only)

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = O.CFG_XXL_T2I
m, _ = build_t2i_module(cfg, dev)
m.train()
x = torch.tanh(torch.randn(B, 3, 512, 512, device=dev))
t = torch.rand(B, device=dev)
y = torch.randn(B, cfg.txt_max_length, cfg.txt_embed_dim, device=dev)
w = torch.randn(B, 3, 512, 512, device=dev)


def step():
    for p in m.parameters():
        p.grad = None
    (m(x, t, y) * w).sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
params = sum(p.numel() for p in m.parameters())
# forward FLOPs per image: image blocks + text blocks + joint keys, x3 for forward + backward
print(f"t2i XXL 512 px, batch {B}, {params / 1e6:.0f} M parameters: {ms:.1f} ms per forward + backward step = {B / ms * 1e3:.1f} images/s "
      f"(eager launches); peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
