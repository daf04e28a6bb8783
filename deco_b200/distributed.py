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
