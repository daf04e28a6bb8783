"""Data-parallel sampling: batch sharded over ranks, no collective inside the loop, ONE all-gather of the final
uint8 images (reference: src/callbacks/save_images.py:56 `pl_module.all_gather(samples)`; shard rule
src/lightning_data.py:142-144)."""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None):
    """One process per GPU (torchrun env).  Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            # asynchronous collectives here are always waited for (OverlappedGradientAverager.finish): keep their tensors
            # alive in the Work object instead of record_stream, which would park 2.7 GB of freed gradient blocks per
            # training step in the caching allocator until a later event query
            os.environ.setdefault("TORCH_NCCL_AVOID_RECORD_STREAMS", "1")
        dist.init_process_group(backend=backend)
    return rank, world, local


def all_gather_images(local_u8: torch.Tensor, world_size: int, total: Optional[int] = None) -> torch.Tensor:
    """Gather [B_local, C, H, W] uint8 shards and undo the rank-strided sharding: sample i of rank r is global
    index r + i * world_size."""
    if world_size == 1:
        return local_u8
    local_u8 = local_u8.contiguous()
    gathered = torch.empty((world_size,) + tuple(local_u8.shape), dtype=local_u8.dtype, device=local_u8.device)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(gathered, local_u8)      # one collective straight into the [world, B_local, ...] buffer
    else:
        dist.all_gather(list(gathered.unbind(0)), local_u8)
    out = gathered.transpose(0, 1).reshape((-1,) + tuple(local_u8.shape[1:]))   # global order r + i*world
    return out if total is None else out[:total]


def all_reduce_gradients(params, world_size: int) -> None:
    """Average the .grad tensors of `params` over the ranks in ONE collective (what DDP does for the reference's
    training step, src/lightning_model.py under `strategy: ddp`): flatten, all-reduce, scatter back in place."""
    if world_size == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat)
        flat.div_(world_size)
    torch._foreach_copy_(grads, torch._utils._unflatten_dense_tensors(flat, grads))


class OverlappedGradientAverager:
    """Average gradients over the ranks WHILE the backward runs: installed as `deco_b200.autograd.GRAD_READY_HOOK`, it
    receives each group of final gradient tensors (one DiT block at a time, then the tail) and starts an asynchronous
    all-reduce of every distinct underlying buffer on NCCL's own stream; `finish()` makes the current stream wait for all
    of them.  Same result as `all_reduce_gradients` (DDP's average, src/lightning_model.py under `strategy: ddp`) without
    the flatten / unflatten copies and with the transfer hidden behind the remaining dgrad / wgrad GEMMs.  Views are
    reduced through their base tensor once (e.g. the per-block slices of the batched adaLN gradient)."""

    def __init__(self, world_size: int):
        self.world = world_size
        self._seen = set()
        self._work = []
        self._avg = dist.is_initialized() and dist.get_backend() == "nccl"

    def __call__(self, tensors) -> None:
        if self.world == 1:
            return
        for t in tensors:
            base = t._base if t._base is not None else t
            key = (base.data_ptr(), base.numel())
            if key in self._seen or base.numel() == 0:
                continue
            if not base.is_contiguous():
                raise RuntimeError("OverlappedGradientAverager needs contiguous gradient buffers")
            self._seen.add(key)
            if self._avg:
                self._work.append((dist.all_reduce(base, op=dist.ReduceOp.AVG, async_op=True), None))
            else:
                self._work.append((dist.all_reduce(base, async_op=True), base))

    def finish(self) -> None:
        for work, base in self._work:
            work.wait()
            if base is not None:
                base.div_(self.world)
        self._work.clear()
        self._seen.clear()


class overlap_gradient_average:
    """`with overlap_gradient_average(world): loss.backward()` -- installs the averager for the deco_b200 denoiser's backward
    and waits for the collectives on exit; afterwards every `param.grad` holds the rank average."""

    def __init__(self, world_size: int, reserve_sms: Optional[int] = None):
        self.avg = OverlappedGradientAverager(world_size)
        # SMs the persistent GEMMs leave to NCCL's kernels while the context is active (include/deco_b200.h,
        # deco_gemm_reserve_sms); pair it with NCCL_MAX_CTAS so NCCL stays inside the reservation
        if reserve_sms is None:
            reserve_sms = int(os.environ.get("DECO_B200_RESERVE_SMS", "0"))
        self.reserve = reserve_sms if world_size > 1 else 0

    def __enter__(self):
        from . import _lib, autograd
        self._prev = autograd.GRAD_READY_HOOK
        autograd.GRAD_READY_HOOK = self.avg
        if self.reserve:
            _lib.load().deco_gemm_reserve_sms(self.reserve)
        return self.avg

    def __exit__(self, *exc):
        from . import _lib, autograd
        autograd.GRAD_READY_HOOK = self._prev
        if self.reserve:
            _lib.load().deco_gemm_reserve_sms(0)
        self.avg.finish()
        return False
