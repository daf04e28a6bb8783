import os

import numpy as np
import torch

from oracle import deco_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a, b, peak):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    mse = float(((a - b) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def cfg_from_array(a) -> O.DenoiserCfg:
    a = [int(v) for v in a]
    return O.DenoiserCfg(in_channels=a[0], num_groups=a[1], hidden_size=a[2], hidden_size_x=a[3], num_blocks=a[4],
                         num_cond_blocks=a[5], patch_size=a[6], num_classes=a[7])


def build_module(cfg: O.DenoiserCfg, device, seed=1234):
    """deco_b200.PixNerDiT holding oracle.seeded_params(cfg) (the same weights the golden fixtures were made with)."""
    from deco_b200 import PixNerDiT
    with torch.device("meta"):
        m = PixNerDiT(in_channels=cfg.in_channels, num_groups=cfg.num_groups, hidden_size=cfg.hidden_size,
                      hidden_size_x=cfg.hidden_size_x, num_blocks=cfg.num_blocks, num_cond_blocks=cfg.num_cond_blocks,
                      patch_size=cfg.patch_size, num_classes=cfg.num_classes)
    P = O.seeded_params(cfg, seed)
    m = m.to_empty(device=device)
    m.load_state_dict({k: v.to(device) for k, v in P.items()})
    return m.eval(), P


def seeded_noise(n, shape, seed0=0):
    return torch.stack([torch.randn(shape, generator=torch.Generator().manual_seed(seed0 + i), dtype=torch.float32)
                        for i in range(n)])


def toy_net(x, t, y):
    """Analytic stand-in network used for the sampler golden fixtures (tests/golden/make_golden.py)."""
    return torch.tanh(x * (0.5 + t.view(-1, 1, 1, 1))) - 0.1 * y.view(-1, 1, 1, 1).float()


def t2i_cfg_from_array(a) -> O.T2ICfg:
    a = [int(v) for v in a]
    return O.T2ICfg(in_channels=a[0], num_groups=a[1], hidden_size=a[2], decoder_hidden_size=a[3],
                    num_encoder_blocks=a[4], num_decoder_blocks=a[5], num_text_blocks=a[6], patch_size=a[7],
                    txt_embed_dim=a[8], txt_max_length=a[9])


def build_t2i_module(cfg: O.T2ICfg, device, seed=4321):
    """deco_b200 t2i PixNerDiT holding oracle.t2i_seeded_params(cfg)."""
    from deco_b200.denoiser_t2i import PixNerDiT
    with torch.device("meta"):
        m = PixNerDiT(in_channels=cfg.in_channels, num_groups=cfg.num_groups, hidden_size=cfg.hidden_size,
                      decoder_hidden_size=cfg.decoder_hidden_size, num_encoder_blocks=cfg.num_encoder_blocks,
                      num_decoder_blocks=cfg.num_decoder_blocks, num_text_blocks=cfg.num_text_blocks,
                      patch_size=cfg.patch_size, txt_embed_dim=cfg.txt_embed_dim, txt_max_length=cfg.txt_max_length)
    P = O.t2i_seeded_params(cfg, seed)
    m = m.to_empty(device=device)
    m.load_state_dict({k: v.to(device) for k, v in P.items()})
    return m.eval(), P


def toy_xnet(x, t, y):
    """Analytic x-prediction stand-in used for the EulerSamplerJiT fixtures (tests/golden/make_golden.py)."""
    return torch.tanh(0.7 * x + 0.3 * t.view(-1, 1, 1, 1)) - 0.05 * y.view(-1, 1, 1, 1).float()


def baseline_cfg_from_array(a) -> O.BaselineCfg:
    a = [int(v) for v in a]
    return O.BaselineCfg(in_channels=a[0], num_groups=a[1], hidden_size=a[2], num_blocks=a[3], patch_size=a[4],
                         num_classes=a[5])


def build_baseline_module(cfg: O.BaselineCfg, device, seed=2468):
    """deco_b200 FlattenDiT holding oracle.baseline_seeded_params(cfg)."""
    from deco_b200.denoiser_baseline import FlattenDiT
    with torch.device("meta"):
        m = FlattenDiT(in_channels=cfg.in_channels, num_groups=cfg.num_groups, hidden_size=cfg.hidden_size,
                       num_blocks=cfg.num_blocks, patch_size=cfg.patch_size, num_classes=cfg.num_classes)
    P = O.baseline_seeded_params(cfg, seed)
    m = m.to_empty(device=device)
    m.load_state_dict({k: v.to(device) for k, v in P.items()})
    return m.eval(), P


def pixnerd_cfg_from_array(a) -> O.PixNerdCfg:
    a = [int(v) for v in a]
    return O.PixNerdCfg(in_channels=a[0], num_groups=a[1], hidden_size=a[2], hidden_size_x=a[3], nerf_mlpratio=a[4],
                        num_blocks=a[5], num_cond_blocks=a[6], patch_size=a[7], num_classes=a[8])


def build_pixnerd_module(cfg: O.PixNerdCfg, device, seed=1357):
    """deco_b200 PixNerd PixNerDiT holding oracle.pixnerd_seeded_params(cfg)."""
    from deco_b200.denoiser_pixnerd import PixNerDiT
    with torch.device("meta"):
        m = PixNerDiT(in_channels=cfg.in_channels, num_groups=cfg.num_groups, hidden_size=cfg.hidden_size,
                      hidden_size_x=cfg.hidden_size_x, nerf_mlpratio=cfg.nerf_mlpratio, num_blocks=cfg.num_blocks,
                      num_cond_blocks=cfg.num_cond_blocks, patch_size=cfg.patch_size, num_classes=cfg.num_classes)
    P = O.pixnerd_seeded_params(cfg, seed)
    m = m.to_empty(device=device)
    m.load_state_dict({k: v.to(device) for k, v in P.items()})
    return m.eval(), P
