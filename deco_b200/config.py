"""Minimal YAML instantiator for the reference's configs (configs_c2i/*.yaml, configs_t2i/*.yaml).

The reference wires everything through LightningCLI / jsonargparse `class_path` + `init_args` (main.py:41-116; the
minimal loader is app.py:49-54).  Here the same YAML is read with PyYAML and the hot-path classes are remapped to
this package; `trainer` / `data` sections and out-of-scope classes are ignored.
"""
from __future__ import annotations

import importlib
from typing import Any, Dict

import yaml

CLASS_MAP = {
    "src.models.transformer.dit_c2i_DeCo.PixNerDiT": "deco_b200.denoiser.PixNerDiT",
    "src.models.transformer.dit_t2i_DeCo.PixNerDiT": "deco_b200.denoiser_t2i.PixNerDiT",
    "src.diffusion.flow_matching.sampling.EulerSampler": "deco_b200.sampling.EulerSampler",
    "src.diffusion.flow_matching.sampling.EulerSamplerJiT": "deco_b200.sampling.EulerSamplerJiT",
    "src.models.transformer.dit_c2i_baseline.FlattenDiT": "deco_b200.denoiser_baseline.FlattenDiT",
    "src.models.transformer.dit_c2i_pixnerd.PixNerDiT": "deco_b200.denoiser_pixnerd.PixNerDiT",
    "src.diffusion.flow_matching.sampling.HeunSampler": "deco_b200.sampling.HeunSampler",
    "src.diffusion.flow_matching.adam_sampling.AdamLMSampler": "deco_b200.sampling.AdamLMSampler",
    "src.diffusion.flow_matching.scheduling.LinearScheduler": "deco_b200.scheduling.LinearScheduler",
    "src.diffusion.flow_matching.training_repa_DeCo.REPATrainer": "deco_b200.training.REPATrainer",
    "src.diffusion.base.guidance.simple_guidance_fn": "deco_b200.sampling.simple_guidance_fn",
    "src.diffusion.flow_matching.sampling.ode_step_fn": "deco_b200.sampling.ode_step_fn",
    "src.diffusion.flow_matching.sampling.sde_mean_step_fn": "deco_b200.sampling.sde_mean_step_fn",
    "src.diffusion.flow_matching.sampling.sde_step_fn": "deco_b200.sampling.sde_step_fn",
    "src.diffusion.flow_matching.sampling.sde_preserve_step_fn": "deco_b200.sampling.sde_preserve_step_fn",
    "src.diffusion.flow_matching.scheduling.GVPScheduler": "deco_b200.scheduling.GVPScheduler",
    "src.diffusion.flow_matching.scheduling.ConstScheduler": "deco_b200.scheduling.ConstScheduler",
    "src.diffusion.flow_matching.adam_sampling.ode_step_fn": "deco_b200.sampling.ode_step_fn",
    "src.utils.model_loader.ModelLoader": "deco_b200.io.ModelLoader",
    "src.models.autoencoder.pixel.PixelAE": "deco_b200.data.PixelAE",
    "src.models.conditioner.class_label.LabelConditioner": "deco_b200.data.LabelConditioner",
    "src.data.dataset.randn.ClassLabelRandomNDataset": "deco_b200.data.ClassLabelRandomNDataset",
}
# init_args that configure out-of-scope subsystems and are dropped
DROPPED_ARGS = {"encoder"}


def resolve(path: str):
    path = CLASS_MAP.get(path, path)
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def _is_dotted_symbol(v) -> bool:
    return isinstance(v, str) and v.startswith("src.") and v in CLASS_MAP


def instantiate(spec: Any):
    """class_path/init_args dicts -> objects; mapped dotted strings -> the callable, or an instance for classes
    (the reference passes `scheduler: src...LinearScheduler` and jsonargparse instantiates it with no args)."""
    if isinstance(spec, dict) and "class_path" in spec:
        cls = resolve(spec["class_path"])
        kwargs = {k: instantiate(v) for k, v in (spec.get("init_args") or {}).items() if k not in DROPPED_ARGS}
        return cls(**kwargs)
    if _is_dotted_symbol(spec):
        obj = resolve(spec)
        return obj() if isinstance(obj, type) else obj
    if isinstance(spec, dict):
        return {k: instantiate(v) for k, v in spec.items()}
    if isinstance(spec, list):
        return [instantiate(v) for v in spec]
    return spec


def load_model_section(yaml_path: str, parts=("vae", "denoiser", "conditioner", "diffusion_trainer",
                                             "diffusion_sampler")) -> Dict[str, Any]:
    with open(yaml_path) as f:
        cfg = yaml.safe_load(f)
    model = cfg["model"]
    return {k: instantiate(model[k]) for k in parts if k in model}
