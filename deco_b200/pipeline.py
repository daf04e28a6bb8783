"""Lightning-free wiring of the sampling path, mirroring the ORIGINAL `LightningModel.predict_step`
(src/__pycache__/lightning_model.cpython-310.pyc, source lines 115-129) and `app.py:81-103`:

    xT, y, metadata = batch
    condition, uncondition = conditioner(y, metadata)
    samples = diffusion_sampler(ema_denoiser, xT, condition, uncondition)
    samples = fp2uint8(vae.decode(samples))            # PixelAE: identity
    all_samples = all_gather(samples)                  # src/callbacks/save_images.py:56

One process per GPU; rank r owns dataset indices r, r + W, ... (DistributedSampler(shuffle=False),
src/lightning_data.py:142-144); no collective inside the sampling loop.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import torch

from . import config, distributed
from .data import ClassLabelRandomNDataset, rank_indices
from .io import ImageSink, ModelLoader


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
