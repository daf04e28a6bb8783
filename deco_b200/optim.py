"""Fused AdamW + EMA step (csrc/optimizer.cu): one launch over every parameter tensor.

Mirrors `torch.optim.AdamW` as configured by the reference (`configs_c2i/DeCo_XL.yaml:89-93`) and the `SimpleEMA` callback
(`src/callbacks/simple_ema.py:27-39`: `ema = decay * ema + (1 - decay) * param` after every optimizer step).  fp32
parameters / gradients / moments, CUDA only (no fallback)."""
from __future__ import annotations

import math
from typing import Iterable, Optional

import torch

from . import _lib
from ._lib import call, ptr


class FusedAdamWEMA:
    def __init__(self, params: Iterable[torch.nn.Parameter], ema_params: Optional[Iterable[torch.Tensor]] = None,
                 lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 ema_decay: float = 0.9999):
        self.params = [p for p in params if p.requires_grad]
        self.ema = list(ema_params) if ema_params is not None else None
        if not self.params:
            raise ValueError("no trainable parameters")
        if self.ema is not None and len(self.ema) != len(self.params):
            raise ValueError("ema_params must pair up with the trainable parameters")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdamWEMA needs contiguous fp32 CUDA parameters (no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        self.step_count = 0
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self._tables = None
        self._grad_ptrs = None
        self._chunks = None
        self._upload_done = None      # CUDA event after the last pointer-table upload (guards the pinned staging buffer)

    def _build_tables(self):
        """Device tables of tensor pointers / CTA chunks.  Rebuilt when a gradient buffer moved (autograd hands out new
        .grad tensors every step); the upload goes through a persistent pinned buffer so that it never synchronises."""
        dev = self.params[0].device
        if self._chunks is None:
            chunk = _lib.load().deco_opt_chunk_elems()
            chunks = []
            for i, p in enumerate(self.params):
                e = self.ema[i] if self.ema is not None else None
                if e is not None and (e.shape != p.shape or e.dtype != torch.float32 or not e.is_contiguous() or e.device != dev):
                    raise RuntimeError("EMA tensors must match their parameters (fp32, contiguous, same device)")
                chunks += [[i, c] for c in range((p.numel() + chunk - 1) // chunk)]
            self._chunks = torch.tensor(chunks, dtype=torch.int32).to(dev)
            self._rows_host = torch.empty((len(self.params), 6), dtype=torch.int64).pin_memory()
            self._rows_dev = torch.empty((len(self.params), 6), dtype=torch.int64, device=dev)
        rows = [[p.data_ptr(), p.grad.data_ptr(), self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr(),
                 self.ema[i].data_ptr() if self.ema is not None else 0, p.numel()] for i, p in enumerate(self.params)]
        # The eager loop may run more than one step ahead of the GPU: the previous asynchronous upload must have READ the
        # pinned buffer before the host rewrites it (otherwise step N's kernel would dereference step N+1's gradient
        # pointers).  Normally the event has long completed and the wait is free.  The device table itself is ordered
        # by the stream (the previous step's kernel precedes this copy).
        if self._upload_done is not None:
            self._upload_done.synchronize()
        self._rows_host.copy_(torch.tensor(rows, dtype=torch.int64))
        self._rows_dev.copy_(self._rows_host, non_blocking=True)
        self._upload_done = torch.cuda.Event()
        self._upload_done.record(torch.cuda.current_stream(dev))
        self._grad_ptrs = [p.grad.data_ptr() for p in self.params]
        self._tables = (self._rows_dev, self._chunks)

    @torch.no_grad()
    def step(self):
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("FusedAdamWEMA.step(): a parameter has no gradient")
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                p.grad = p.grad.float().contiguous()
        if self._tables is None or self._grad_ptrs != [p.grad.data_ptr() for p in self.params]:
            self._build_tables()          # gradient buffers moved (first step, or .grad was re-created)
        self.step_count += 1
        b1, b2 = self.betas
        tens, chunks = self._tables
        call("deco_adamw_ema_step", ptr(tens), ptr(chunks), chunks.shape[0], float(self.lr), float(b1), float(b2),
             float(self.eps), float(self.weight_decay), 1.0 - math.pow(b1, self.step_count),
             1.0 - math.pow(b2, self.step_count), float(self.ema_decay),
             torch.cuda.current_stream(self.params[0].device).cuda_stream)
        # the kernel wrote the parameters behind autograd's back: bump their version counters (no kernel) so that caches
        # keyed on them -- PixNerDiT.prepare()'s bf16 / packed weights -- are rebuilt
        torch.autograd.graph.increment_version(self.params)
        if self.ema is not None:
            torch.autograd.graph.increment_version(self.ema)

    # ------------------------------------------------------------------ checkpointing (torch.optim.AdamW layout)
    def state_dict(self) -> dict:
        """Same layout as `torch.optim.AdamW.state_dict()` (what Lightning writes into `optimizer_states[0]`): `state` =
        {index: {step, exp_avg, exp_avg_sq}} over the trainable parameters in order, one `param_groups` entry."""
        state = {i: dict(step=torch.tensor(float(self.step_count)), exp_avg=self.exp_avg[i], exp_avg_sq=self.exp_avg_sq[i])
                 for i in range(len(self.params))}
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay, amsgrad=False,
                     maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                     decoupled_weight_decay=True,
                     params=list(range(len(self.params))))
        return dict(state=state, param_groups=[group])

    @torch.no_grad()
    def load_state_dict(self, sd: dict) -> None:
        """Accepts this class's `state_dict()` or one written by `torch.optim.AdamW` over the same parameter list."""
        groups = sd["param_groups"]
        order = [i for g in groups for i in g["params"]]
        if len(order) != len(self.params):
            raise ValueError(f"optimizer state has {len(order)} parameters, this optimizer {len(self.params)}")
        g0 = groups[0]
        self.lr, self.betas, self.eps = g0["lr"], tuple(g0["betas"]), g0["eps"]
        self.weight_decay = g0["weight_decay"]
        steps = set()
        for slot, idx in enumerate(order):
            st = sd["state"].get(idx)
            if st is None:                      # a parameter that never received a gradient
                self.exp_avg[slot].zero_()
                self.exp_avg_sq[slot].zero_()
                continue
            if st["exp_avg"].shape != self.params[slot].shape:
                raise ValueError(f"optimizer state {idx}: shape {tuple(st['exp_avg'].shape)} != {tuple(self.params[slot].shape)}")
            self.exp_avg[slot].copy_(st["exp_avg"])
            self.exp_avg_sq[slot].copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): the fused kernel keeps one")
        self.step_count = steps.pop() if steps else 0

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
