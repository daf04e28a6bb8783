"""The two steps either side of the sampling path: checkpoint loading and the image sink.

Mirrors (reference paths):
  src/utils/model_loader.py:10-28          ModelLoader.load: copy `ema_denoiser.*` / `denoiser.*` entries of a Lightning
                                           checkpoint's `state_dict` into the denoiser by parameter name
  app.py:56-63                             load_model (always the EMA weights)
  src/lightning_model.py:322-368           LightningModel.state_dict / load_state_dict: checkpoint keys are
                                           `denoiser.*`, `ema_denoiser.*`, `diffusion_trainer.*` (the trainer contributes
                                           nothing, training_repa_DeCo.py:290-291); `_orig_mod.` (torch.compile) and
                                           `.module.` (DDP) fragments are stripped from incoming keys
  src/callbacks/save_images.py:31-64       SaveImagesHook: NHWC uint8 samples -> per-image files on a thread pool and/or
                                           one `output.npz` (`arr_0`) for the ADM FID suite
  src/data/dataset/randn.py:34-36          save_fn: `{target_dir}/{filename}.png`

PNG encoding is done here with zlib + struct (8-bit RGB, filter 0) so the sink has no imaging dependency.
"""
from __future__ import annotations

import logging
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

logger = logging.getLogger(__name__)


class ModelLoader:
    """src/utils/model_loader.py:10-28.  `denoiser.weight_path` / `.load_ema` select file and prefix; entries that are
    missing or mis-shaped are reported and skipped, as in the reference."""

    def load(self, denoiser, weight: Optional[Dict] = None):
        if weight is None:
            if not getattr(denoiser, "weight_path", None):
                return denoiser
            weight = torch.load(denoiser.weight_path, map_location=torch.device("cpu"))
        prefix = "ema_denoiser." if getattr(denoiser, "load_ema", False) else "denoiser."
        load_prefixed_state_dict(denoiser, weight["state_dict"], prefix)
        return denoiser


@torch.no_grad()
def load_prefixed_state_dict(module: torch.nn.Module, state_dict: Dict[str, torch.Tensor], prefix: str) -> List[str]:
    """Copy `state_dict[prefix + name]` into every entry of `module.state_dict()`; returns the names that failed."""
    failed = []
    for k, v in module.state_dict().items():
        try:
            v.copy_(state_dict[prefix + k])
        except Exception:   # noqa: BLE001  (the reference swallows every failure and logs it)
            logger.warning("Failed to copy %s to denoiser weight", prefix + k)
            failed.append(k)
    return failed


def clean_checkpoint_keys(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """src/lightning_model.py:322-331 and :352-368: checkpoints written while the denoiser was wrapped by torch.compile
    (`denoiser._orig_mod.blocks...`) or DDP (`denoiser.module.blocks...`) load into the bare modules."""
    return {k.replace("_orig_mod.", "").replace(".module.", "."): v for k, v in state_dict.items()}


def lightning_state_dict(denoiser: torch.nn.Module, ema_denoiser: torch.nn.Module,
                         diffusion_trainer: Optional[torch.nn.Module] = None) -> Dict[str, torch.Tensor]:
    """The `state_dict` entry of a reference checkpoint (src/lightning_model.py:333-350): `denoiser.` + `ema_denoiser.` +
    `diffusion_trainer.` prefixed entries, in that order."""
    out: Dict[str, torch.Tensor] = {}
    for prefix, mod in (("denoiser.", denoiser), ("ema_denoiser.", ema_denoiser)):
        for k, v in mod.state_dict().items():
            out[prefix + k] = v
    if diffusion_trainer is not None:
        sd = diffusion_trainer.state_dict()          # REPATrainer: None (its buffers are rebuilt from the config)
        for k, v in (sd or {}).items():
            out["diffusion_trainer." + k] = v
    return out


def save_checkpoint(path: str, denoiser, ema_denoiser, diffusion_trainer=None, optimizer=None, global_step: int = 0,
                    epoch: int = 0, ema_decay: Optional[float] = None, extra: Optional[Dict] = None) -> str:
    """Write a checkpoint in the layout Lightning gives the reference's LightningModel: `state_dict` (keys as above, CPU
    tensors), `optimizer_states` (torch.optim.AdamW layout), `global_step`, `epoch` and the SimpleEMA callback state
    (src/callbacks/simple_ema.py:51-55).  `ModelLoader` / `app.py:56-63` read `state_dict` only."""
    sd = {k: v.detach().to("cpu", copy=True) for k, v in lightning_state_dict(denoiser, ema_denoiser, diffusion_trainer).items()}
    ckpt = dict(state_dict=sd, global_step=int(global_step), epoch=int(epoch))
    if optimizer is not None:
        osd = optimizer.state_dict()
        osd["state"] = {i: {k: (v.detach().to("cpu", copy=True) if torch.is_tensor(v) else v) for k, v in st.items()}
                        for i, st in osd["state"].items()}
        ckpt["optimizer_states"] = [osd]
    if ema_decay is not None:
        ckpt["callbacks"] = {"SimpleEMA": dict(decay=float(ema_decay), every_n_steps=1)}
    if extra:
        ckpt.update(extra)
    tmp = path + ".tmp"
    torch.save(ckpt, tmp)
    os.replace(tmp, path)          # never leave a half-written checkpoint under the final name
    return path


def load_checkpoint(path_or_dict, denoiser, ema_denoiser=None, optimizer=None, strict: bool = True) -> Dict:
    """Inverse of save_checkpoint; also accepts the reference's own checkpoints (keys cleaned as in
    src/lightning_model.py:322-368).  Returns the checkpoint dict (for `global_step`, `epoch`, callbacks)."""
    ckpt = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu")
    sd = clean_checkpoint_keys(ckpt["state_dict"])
    for prefix, mod in (("denoiser.", denoiser), ("ema_denoiser.", ema_denoiser)):
        if mod is None:
            continue
        sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        if not sub and prefix == "ema_denoiser.":
            sub = {k[len("denoiser."):]: v for k, v in sd.items() if k.startswith("denoiser.")}   # no EMA stored yet
        mod.load_state_dict(sub, strict=strict)
    if optimizer is not None and ckpt.get("optimizer_states"):
        optimizer.load_state_dict(ckpt["optimizer_states"][0])
    return ckpt


def encode_png(img: np.ndarray) -> bytes:
    """8-bit RGB (or grey) PNG of an [H, W, C] uint8 array."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] in (1, 3)
    h, w, c = img.shape
    raw = np.empty((h, 1 + w * c), dtype=np.uint8)
    raw[:, 0] = 0                       # filter type 0 on every scanline
    raw[:, 1:] = img.reshape(h, w * c)

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    ihdr = struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 0, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(raw.tobytes(), 3)) + chunk(b"IEND", b"")


def save_png(sample: np.ndarray, metadata: Dict, target_dir: str) -> str:
    """The dataset's save_fn (randn.py:34-36)."""
    path = os.path.join(target_dir, f"{metadata['filename']}.png")
    with open(path, "wb") as f:
        f.write(encode_png(sample))
    return path


class ImageSink:
    """SaveImagesHook without Lightning: feed it the (already all-gathered, global-order) uint8 batches."""

    def __init__(self, target_dir: str, save_compressed: bool = False, max_workers: int = 8):
        self.target_dir = target_dir
        self.save_compressed = save_compressed
        os.makedirs(target_dir, exist_ok=True)
        self.pool = ThreadPoolExecutor(max_workers=max_workers)
        self.samples: List[np.ndarray] = []
        self._saved_num = 0
        self._futures = []

    def process_batch(self, images_u8: torch.Tensor, metadatas: Sequence[Dict], gathered_u8: Optional[torch.Tensor] = None,
                      is_global_zero: bool = True):
        """images_u8 [B, C, H, W] uint8: THIS rank's samples with their metadata -> per-image files (every rank writes
        its own, save_images.py:52-54: all of them, or only the first 10 when `save_compressed`).  gathered_u8: the
        all-gathered batch; kept on the global-zero rank for output.npz when `save_compressed` (:56-59)."""
        if not self.save_compressed or self._saved_num < 10:
            nhwc = images_u8.permute(0, 2, 3, 1).contiguous().cpu().numpy()
            self._saved_num += nhwc.shape[0]
            for sample, md in zip(nhwc, metadatas):
                fn = md.get("save_fn", save_png)
                self._futures.append(self.pool.submit(fn, sample, {k: v for k, v in md.items() if k != "save_fn"},
                                                      self.target_dir))
        if self.save_compressed and is_global_zero:
            g = images_u8 if gathered_u8 is None else gathered_u8
            self.samples.append(g.permute(0, 2, 3, 1).contiguous().cpu().numpy())

    def close(self) -> Optional[str]:
        for f in self._futures:
            f.result()
        self.pool.shutdown(wait=True)
        path = None
        if self.save_compressed and self.samples:
            path = os.path.join(self.target_dir, "output.npz")
            np.savez(path, arr_0=np.concatenate(self.samples))
        self.samples = []
        return path
