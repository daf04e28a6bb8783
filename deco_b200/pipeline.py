"""Lightning-free wiring of the sampling path, mirroring the ORIGINAL `LightningModel.predict_step`
(src/__pycache__/lightning_model.cpython-310.pyc, source lines 115-129) and `app.py:81-103`:

    xT, y, metadata = batch
    condition, uncondition = conditioner(y, metadata)
    samples = diffusion_sampler(ema_denoiser, xT, condition, uncondition)
    samples = fp2uint8(vae.decode(samples))            # PixelAE: identity
    all_samples = all_gather(samples)                  # src/callbacks/save_images.py:56

One process per GPU; rank r owns dataset indices r, r + W, ... (DistributedSampler(shuffle=False),
src/lightning_data.py:142-144); no collective inside the sampling loop.

`TrainingPipeline` is the same for the ORIGINAL `LightningModel.training_step` (pyc source lines 105-113):

    x, y, metadata = batch
    x = vae.encode(x)                                   # PixelAE: identity for scale 1, shift 0
    condition, uncondition = conditioner(y, metadata)
    loss = diffusion_trainer(denoiser, ema_denoiser, diffusion_sampler, x, condition, uncondition, metadata)
    loss["loss"].backward() -> [DDP gradient average] -> AdamW step -> SimpleEMA.ema_step

plus what Lightning does around it (configure_optimizers over the denoiser's trainable parameters, src/lightning_model.py:
163-177; `ema_denoiser = deepcopy(denoiser)` kept in fp32 and excluded from gradients, :48,:85-92; checkpoints in the
`denoiser. / ema_denoiser. / diffusion_trainer.` layout, :333-350).
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import torch

from . import config, distributed
from .data import ClassLabelRandomNDataset, rank_indices
from .io import ImageSink, ModelLoader, load_checkpoint, save_checkpoint


class SamplingPipeline:
    def __init__(self, denoiser, sampler, conditioner, vae=None, device="cuda"):
        self.device = torch.device(device)
        self.denoiser = ModelLoader().load(denoiser).to(self.device).eval()
        self.sampler = sampler
        self.conditioner = conditioner
        self.vae = vae

    @classmethod
    def from_yaml(cls, yaml_path: str, device="cuda") -> "SamplingPipeline":
        parts = config.load_model_section(yaml_path, parts=("vae", "denoiser", "conditioner", "diffusion_sampler"))
        return cls(parts["denoiser"], parts["diffusion_sampler"], parts["conditioner"], parts.get("vae"), device)

    @torch.no_grad()
    def predict_step(self, batch) -> torch.Tensor:
        """(xT [B,C,H,W] fp32, y labels, metadata) -> uint8 images [B,C,H,W] on the device."""
        xT, y, metadata = batch
        xT = xT.to(self.device, non_blocking=True)
        condition, uncondition = self.conditioner(y, metadata, device=self.device)
        x, u8 = self.sampler.sample_uint8(self.denoiser, xT, condition, uncondition)
        scale, shift = (getattr(self.vae, "scale", 1.0), getattr(self.vae, "shift", 0.0)) if self.vae is not None else (1.0, 0.0)
        if scale != 1.0 or shift != 0.0:     # non-identity PixelAE: decode, then quantise
            from . import ops
            u8 = ops.fp2uint8(self.vae.decode(x))
        return u8

    def predict(self, dataset: ClassLabelRandomNDataset, batch_size: int, rank: int = 0, world_size: int = 1,
                sink: Optional[ImageSink] = None) -> Iterator[Tuple[torch.Tensor, List[Dict]]]:
        """Iterate this rank's shard; yields (global-order uint8 batch after all-gather, metadata of the LOCAL shard).
        The sink gets this rank's images + metadata (per-image files) and the gathered batch (output.npz on rank 0),
        like SaveImagesHook.process_batch."""
        idx = rank_indices(len(dataset), rank, world_size)
        for s in range(0, len(idx), batch_size):
            items = [dataset[i] for i in idx[s:s + batch_size]]
            xT = torch.stack([it[0] for it in items]).pin_memory() if torch.cuda.is_available() else torch.stack([it[0] for it in items])
            y = [it[1] for it in items]
            md = [it[2] for it in items]
            u8 = self.predict_step((xT, y, md))
            gathered = distributed.all_gather_images(u8, world_size)
            if sink is not None:
                sink.process_batch(u8, md, gathered_u8=gathered, is_global_zero=(rank == 0))
            yield gathered, md


class TrainingPipeline:
    """One optimisation step of the class-conditional model without Lightning (see the module docstring).  AdamW and the
    EMA update are ONE fused kernel (deco_b200.optim.FusedAdamWEMA = torch.optim.AdamW + SimpleEMA.ema_step); with
    world_size > 1 gradients are averaged like DDP does (deco_b200.distributed.all_reduce_gradients)."""

    def __init__(self, denoiser, diffusion_trainer, diffusion_sampler=None, conditioner=None, vae=None, lr: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, ema_decay: float = 0.9999,
                 device="cuda", world_size: int = 1):
        import copy
        from .optim import FusedAdamWEMA
        self.device = torch.device(device)
        self.denoiser = ModelLoader().load(denoiser).to(self.device).train()
        self.ema_denoiser = copy.deepcopy(self.denoiser).to(torch.float32).eval()      # lightning_model.py:48, :193-203
        for p in self.ema_denoiser.parameters():
            p.requires_grad_(False)                                                     # no_grad(ema_denoiser), :91
        self.diffusion_trainer = diffusion_trainer.to(self.device)
        self.diffusion_sampler = diffusion_sampler
        self.conditioner = conditioner
        self.vae = vae
        self.world_size = world_size
        self.ema_decay = ema_decay
        self.params = [p for p in self.denoiser.parameters() if p.requires_grad]       # filter_nograd_tensors, :164
        pairs = dict(self.ema_denoiser.named_parameters())
        ema = [pairs[n] for n, p in self.denoiser.named_parameters() if p.requires_grad]
        self.optimizer = FusedAdamWEMA(self.params, ema, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                       ema_decay=ema_decay)
        self.global_step = 0

    @classmethod
    def from_yaml(cls, yaml_path: str, device="cuda", **kw) -> "TrainingPipeline":
        parts = config.load_model_section(yaml_path, parts=("vae", "denoiser", "conditioner", "diffusion_trainer",
                                                            "diffusion_sampler"))
        return cls(parts["denoiser"], parts["diffusion_trainer"], parts.get("diffusion_sampler"), parts.get("conditioner"),
                   parts.get("vae"), device=device, **kw)

    def training_step(self, batch) -> Dict[str, torch.Tensor]:
        """(images [B,C,H,W] in [-1, 1], labels, metadata) -> the trainer's loss dict (detached); parameters, EMA and
        optimizer state are updated in place."""
        x, y, metadata = batch
        x = x.to(self.device, non_blocking=True)
        if self.vae is not None:
            with torch.no_grad():
                x = self.vae.encode(x)
        if self.conditioner is not None:
            condition, uncondition = self.conditioner(y, metadata, device=self.device)
        else:
            condition, uncondition = y
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.diffusion_trainer(self.denoiser, self.ema_denoiser, self.diffusion_sampler, x, condition, uncondition,
                                      metadata)
        loss["loss"].backward()
        if self.world_size > 1:
            distributed.all_reduce_gradients(self.params, self.world_size)
        self.optimizer.step()
        self.global_step += 1
        return {k: v.detach() for k, v in loss.items()}

    # ---------------------------------------------------------------------------------------- checkpoints
    def state_dict(self) -> Dict[str, torch.Tensor]:
        from .io import lightning_state_dict
        return lightning_state_dict(self.denoiser, self.ema_denoiser, self.diffusion_trainer)

    def save_checkpoint(self, path: str) -> str:
        return save_checkpoint(path, self.denoiser, self.ema_denoiser, self.diffusion_trainer, self.optimizer,
                               global_step=self.global_step, ema_decay=self.ema_decay)

    def load_checkpoint(self, path_or_dict, strict: bool = True) -> None:
        ckpt = load_checkpoint(path_or_dict, self.denoiser, self.ema_denoiser, self.optimizer, strict=strict)
        self.global_step = int(ckpt.get("global_step", 0))
        cb = (ckpt.get("callbacks") or {}).get("SimpleEMA")
        if cb:
            self.ema_decay = self.optimizer.ema_decay = float(cb["decay"])
