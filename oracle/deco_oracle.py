"""CPU/torch restatement of the DeCo hot path -- TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: a plain-PyTorch (fp32 by default) restatement of the
reference algorithm for the path named in BASELINE.json.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product (`deco_b200/`) never does.

Parity status: PINNED.  `oracle/validate_against_reference.py` (run in the build
container, where /root/reference exists) checks every function below against the
live reference modules on identical weights/inputs, and `tests/golden/make_golden.py`
stores reference outputs as fixtures that the CPU test-suite re-checks.

Everything is functional: parameters come in as a flat ``dict[str, Tensor]`` whose
keys are the reference module's ``state_dict`` names (the checkpoint contract,
SURVEY.md section 8a).  Using F.linear / F.layer_norm / F.scaled_dot_product_attention
means `torch.autocast(..., dtype=torch.bfloat16)` reproduces the reference's
mixed-precision rounding points exactly as well.

Reference citations are `path:line` under /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- config
@dataclass(frozen=True)
class DenoiserCfg:
    """Constructor arguments of the class-conditional denoiser
    (src/models/transformer/dit_c2i_DeCo.py:417-433)."""
    in_channels: int = 3
    num_groups: int = 16
    hidden_size: int = 1152
    hidden_size_x: int = 32
    num_blocks: int = 31
    num_cond_blocks: int = 28
    patch_size: int = 16
    num_classes: int = 1000
    max_freqs: int = 8

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_groups

    @property
    def ffn_hidden(self) -> int:
        # FlattenDiTBlock: mlp_hidden_dim=int(hidden*4.0); FeedForward: int(2*h/3)
        # (dit_c2i_DeCo.py:200-201, :108)
        return int(2 * int(self.hidden_size * 4.0) / 3)

    @property
    def num_res_blocks(self) -> int:
        return self.num_blocks - self.num_cond_blocks


CFG_XL = DenoiserCfg()                                            # configs_c2i/DeCo_XL.yaml:44-55
CFG_L = DenoiserCfg(hidden_size=1024, num_blocks=25, num_cond_blocks=22)  # configs_c2i/DeCo_large.yaml


def param_shapes(cfg: DenoiserCfg) -> Dict[str, Tuple[int, ...]]:
    """state_dict names and shapes of the reference module (SURVEY.md 8a contract)."""
    H, Hx, p, C = cfg.hidden_size, cfg.hidden_size_x, cfg.patch_size, cfg.in_channels
    d, ffn = cfg.head_dim, cfg.ffn_hidden
    s: Dict[str, Tuple[int, ...]] = {
        "x_embedder.embedder.0.weight": (Hx, C + cfg.max_freqs ** 2),
        "x_embedder.embedder.0.bias": (Hx,),
        "s_embedder.proj.weight": (H, C * p * p),
        "s_embedder.proj.bias": (H,),
        "t_embedder.mlp.0.weight": (H, 256),
        "t_embedder.mlp.0.bias": (H,),
        "t_embedder.mlp.2.weight": (H, H),
        "t_embedder.mlp.2.bias": (H,),
        "y_embedder.embedding_table.weight": (cfg.num_classes + 1, H),
    }
    for i in range(cfg.num_cond_blocks):
        b = f"blocks.{i}."
        s[b + "norm1.weight"] = (H,)
        s[b + "attn.qkv.weight"] = (3 * H, H)
        s[b + "attn.q_norm.weight"] = (d,)
        s[b + "attn.k_norm.weight"] = (d,)
        s[b + "attn.proj.weight"] = (H, H)
        s[b + "attn.proj.bias"] = (H,)
        s[b + "norm2.weight"] = (H,)
        s[b + "mlp.w1.weight"] = (ffn, H)
        s[b + "mlp.w3.weight"] = (ffn, H)
        s[b + "mlp.w2.weight"] = (H, ffn)
        s[b + "adaLN_modulation.0.weight"] = (6 * H, H)
        s[b + "adaLN_modulation.0.bias"] = (6 * H,)
    s["dec_net.cond_embed.weight"] = (p * p * Hx, H)
    s["dec_net.cond_embed.bias"] = (p * p * Hx,)
    s["dec_net.input_proj.weight"] = (Hx, Hx)
    s["dec_net.input_proj.bias"] = (Hx,)
    for j in range(cfg.num_res_blocks):
        b = f"dec_net.res_blocks.{j}."
        s[b + "in_ln.weight"] = (Hx,)
        s[b + "in_ln.bias"] = (Hx,)
        s[b + "mlp.0.weight"] = (Hx, Hx)
        s[b + "mlp.0.bias"] = (Hx,)
        s[b + "mlp.2.weight"] = (Hx, Hx)
        s[b + "mlp.2.bias"] = (Hx,)
        s[b + "adaLN_modulation.1.weight"] = (3 * Hx, Hx)
        s[b + "adaLN_modulation.1.bias"] = (3 * Hx,)
    s["dec_net.final_layer.linear.weight"] = (C, Hx)
    s["dec_net.final_layer.linear.bias"] = (C,)
    return s


def seeded_params(cfg: DenoiserCfg, seed: int = 1234, device="cpu") -> Params:
    """Deterministic, *fully non-zero* random weights keyed by parameter name.

    The reference's default init zeroes the decoder output layers
    (dit_c2i_DeCo.py:386-393) so parity on default init is vacuous.
    Each tensor is drawn from its own generator seeded with (seed, name index) so the
    same weights can be produced for the reference module, the oracle and the CUDA
    module on any box without shipping a checkpoint.
    """
    out: Params = {}
    for idx, (name, shape) in enumerate(sorted(param_shapes(cfg).items())):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if len(shape) == 2 and not name.startswith("y_embedder"):
            std = 1.0 / math.sqrt(shape[1])
            if "adaLN_modulation" in name:
                std *= 0.5
            w = torch.randn(shape, generator=g) * std
        elif name.startswith("y_embedder"):
            w = torch.randn(shape, generator=g) * 0.5
        elif name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight") \
                or name.endswith("in_ln.weight"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:  # biases
            w = 0.05 * torch.randn(shape, generator=g)
        out[name] = w.to(device)
    return out


# --------------------------------------------------------------------------- tables
def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: float = 10.0) -> torch.Tensor:
    """[cos || sin] sinusoid with max_period 10 (dit_c2i_DeCo.py:43-53)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period)
                      * torch.arange(0, half, dtype=torch.float32, device=t.device) / half)
    args = t[..., None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def rope_table_2d(head_dim: int, height: int, width: int, theta: float = 10000.0,
                  scale: float = 16.0) -> torch.Tensor:
    """2-D axial RoPE angles, returned as real angles [L, head_dim/2] (pair 2k -> x, 2k+1 -> y)
    (dit_c2i_DeCo.py:116-131: positions linspace(0,scale,W), freqs theta^(-4k/d))."""
    x_pos = torch.linspace(0, scale, width)
    y_pos = torch.linspace(0, scale, height)
    y_pos, x_pos = torch.meshgrid(y_pos, x_pos, indexing="ij")
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 4)[: head_dim // 4].float() / head_dim))
    xa = torch.outer(x_pos.reshape(-1), freqs).float()
    ya = torch.outer(y_pos.reshape(-1), freqs).float()
    return torch.stack([xa, ya], dim=-1).reshape(height * width, -1)


def apply_rope(x: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    """x: [B, N, heads, d]; rotate consecutive pairs (2j, 2j+1) by angles[n, j]
    (dit_c2i_DeCo.py:134-145, written with real arithmetic)."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    cos = torch.cos(angles)[None, :, None, :]
    sin = torch.sin(angles)[None, :, None, :]
    a, b = xf[..., 0], xf[..., 1]
    out = torch.stack([a * cos - b * sin, a * sin + b * cos], dim=-1)
    return out.flatten(3).type_as(x)


def nerf_pos_table(patch_size: int, max_freqs: int = 8) -> torch.Tensor:
    """Constant per-pixel positional table [p*p, max_freqs^2] (dit_c2i_DeCo.py:221-236)."""
    pos = torch.linspace(0, 1, patch_size)
    pos_y, pos_x = torch.meshgrid(pos, pos, indexing="ij")
    pos_x = pos_x.reshape(-1, 1, 1)
    pos_y = pos_y.reshape(-1, 1, 1)
    freqs = torch.linspace(0, max_freqs, max_freqs)
    fx, fy = freqs[None, :, None], freqs[None, None, :]
    coeffs = (1 + fx * fy) ** -1
    return (torch.cos(pos_x * fx * torch.pi) * torch.cos(pos_y * fy * torch.pi) * coeffs
            ).view(-1, max_freqs ** 2)


# --------------------------------------------------------------------------- layers
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """dit_c2i_DeCo.py:94-99 (fp32 statistics, cast back, then weight * x)."""
    dt = x.dtype
    xf = x.to(torch.float32)
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return weight * xf.to(dt)


def modulate(x, shift, scale):
    """dit_c2i_DeCo.py:11-12."""
    return x * (1 + scale) + shift


def attention(P: Params, pre: str, x: torch.Tensor, angles: torch.Tensor, heads: int,
              mask=None) -> torch.Tensor:
    """RAttention.forward (dit_c2i_DeCo.py:174-190)."""
    B, N, C = x.shape
    d = C // heads
    qkv = F.linear(x, P[pre + "qkv.weight"]).reshape(B, N, 3, heads, d).permute(2, 0, 1, 3, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = rmsnorm(q, P[pre + "q_norm.weight"])
    k = rmsnorm(k, P[pre + "k_norm.weight"])
    q, k = apply_rope(q, angles), apply_rope(k, angles)
    q, k, v = q.transpose(1, 2), k.transpose(1, 2).contiguous(), v.transpose(1, 2).contiguous()
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0)
    o = o.transpose(1, 2).reshape(B, N, C)
    return F.linear(o, P[pre + "proj.weight"], P[pre + "proj.bias"])


def feed_forward(P: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward.forward (dit_c2i_DeCo.py:112-114)."""
    return F.linear(F.silu(F.linear(x, P[pre + "w1.weight"])) * F.linear(x, P[pre + "w3.weight"]),
                    P[pre + "w2.weight"])


def dit_block(P: Params, i: int, x, c, angles, heads, mask=None):
    """FlattenDiTBlock.forward (dit_c2i_DeCo.py:206-210)."""
    b = f"blocks.{i}."
    mod = F.linear(c, P[b + "adaLN_modulation.0.weight"], P[b + "adaLN_modulation.0.bias"])
    sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=-1)
    x = x + g1 * attention(P, b + "attn.", modulate(rmsnorm(x, P[b + "norm1.weight"]), sh1, sc1),
                           angles, heads, mask)
    x = x + g2 * feed_forward(P, b + "mlp.", modulate(rmsnorm(x, P[b + "norm2.weight"]), sh2, sc2))
    return x


def pixel_decoder(P: Params, cfg: DenoiserCfg, x: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """SimpleMLPAdaLN.forward + ResBlock + decoder FinalLayer
    (dit_c2i_DeCo.py:395-415, :313-317, :329-332).  x: [BL, p*p, Hx], s: [BL, H]."""
    Hx = cfg.hidden_size_x
    x = F.linear(x, P["dec_net.input_proj.weight"], P["dec_net.input_proj.bias"])
    y = F.linear(s, P["dec_net.cond_embed.weight"], P["dec_net.cond_embed.bias"])
    y = y.reshape(y.shape[0], cfg.patch_size ** 2, -1)
    for j in range(cfg.num_res_blocks):
        b = f"dec_net.res_blocks.{j}."
        mod = F.linear(F.silu(y), P[b + "adaLN_modulation.1.weight"], P[b + "adaLN_modulation.1.bias"])
        sh, sc, g = mod.chunk(3, dim=-1)
        h = modulate(F.layer_norm(x, (Hx,), P[b + "in_ln.weight"], P[b + "in_ln.bias"], 1e-6), sh, sc)
        h = F.linear(F.silu(F.linear(h, P[b + "mlp.0.weight"], P[b + "mlp.0.bias"])),
                     P[b + "mlp.2.weight"], P[b + "mlp.2.bias"])
        x = x + g * h
    x = F.layer_norm(x, (Hx,), None, None, 1e-6)
    return F.linear(x, P["dec_net.final_layer.linear.weight"], P["dec_net.final_layer.linear.bias"])


def denoiser_forward(P: Params, cfg: DenoiserCfg, x: torch.Tensor, t: torch.Tensor, y: torch.Tensor,
                     s: Optional[torch.Tensor] = None, mask=None, return_s: bool = False):
    """PixNerDiT.forward (dit_c2i_DeCo.py:488-510).  x:[B,C,H,W] t:[B] y:[B] int64."""
    B, _, Hh, Ww = x.shape
    p, H = cfg.patch_size, cfg.hidden_size
    angles = rope_table_2d(cfg.head_dim, Hh // p, Ww // p).to(x.device)
    xp = F.unfold(x, kernel_size=p, stride=p).transpose(1, 2)                      # [B, L, C*p*p]
    tf = timestep_embedding(t.view(-1))
    te = F.linear(F.silu(F.linear(tf, P["t_embedder.mlp.0.weight"], P["t_embedder.mlp.0.bias"])),
                  P["t_embedder.mlp.2.weight"], P["t_embedder.mlp.2.bias"]).view(B, -1, H)
    ye = F.embedding(y, P["y_embedder.embedding_table.weight"]).view(B, 1, H)
    c = F.silu(te + ye)
    if s is None:
        s = F.linear(xp, P["s_embedder.proj.weight"], P["s_embedder.proj.bias"])
        for i in range(cfg.num_cond_blocks):
            s = dit_block(P, i, s, c, angles, cfg.num_groups, mask)
        s = F.silu(te + s)
    Bn, L, _ = s.shape
    px = xp.reshape(Bn * L, cfg.in_channels, p * p).transpose(1, 2)               # [BL, p*p, C]
    sf = s.reshape(Bn * L, H)
    tab = nerf_pos_table(p, cfg.max_freqs).to(device=px.device, dtype=px.dtype)
    emb_in = torch.cat([px, tab[None].expand(Bn * L, -1, -1)], dim=-1)
    px = F.linear(emb_in, P["x_embedder.embedder.0.weight"], P["x_embedder.embedder.0.bias"])
    out = pixel_decoder(P, cfg, px, sf)                                            # [BL, p*p, C]
    out = out.transpose(1, 2).reshape(Bn, L, -1)
    out = F.fold(out.transpose(1, 2).contiguous(), (Hh, Ww), kernel_size=p, stride=p)
    return (out, s) if return_s else out


# --------------------------------------------------------------------------- scheduler / sampler
def shift_respace(t, shift: float = 3.0):
    """flow_matching/sampling.py:11-12."""
    return t / (t + (1 - t) * shift)


def make_timesteps(num_steps: int, timeshift: float = 1.0, last_step: Optional[float] = None) -> torch.Tensor:
    """fp32 schedule linspace(0, 1-last, n) || 1.0, then timeshift (sampling.py:52-57)."""
    if last_step is None or num_steps == 1:
        last_step = 1.0 / num_steps
    ts = torch.linspace(0.0, 1 - last_step, num_steps)
    ts = torch.cat([ts, torch.tensor([1.0])], dim=0)
    return shift_respace(ts, timeshift)


def cfg_combine(out: torch.Tensor, g: float) -> torch.Tensor:
    """simple_guidance_fn (base/guidance.py:3-6): rows [uncond || cond]."""
    u, c = out.chunk(2, dim=0)
    return u + g * (c - u)


def euler_sample(net: Callable, noise: torch.Tensor, cond: torch.Tensor, uncond: torch.Tensor,
                 num_steps: int, guidance: float, gmin: float = 0.0, gmax: float = 1.0,
                 timeshift: float = 1.0, last_step: Optional[float] = None,
                 return_trajs: bool = False):
    """EulerSampler._impl_sampling with ode_step_fn and LinearScheduler
    (flow_matching/sampling.py:66-107).  Guidance applies iff gmin < t <= gmax (:93)."""
    steps = make_timesteps(num_steps, timeshift, last_step).to(noise.device, noise.dtype)
    B = noise.shape[0]
    cfg_c = torch.cat([uncond, cond], dim=0)
    x = noise
    xs, vs = [noise], []
    for t_cur, t_next in zip(steps[:-1], steps[1:]):
        dt = t_next - t_cur
        out = net(torch.cat([x, x], 0), t_cur.repeat(2 * B), cfg_c)
        g = guidance if (t_cur > gmin and t_cur <= gmax) else 1.0
        v = cfg_combine(out, g)
        x = x + v * dt
        xs.append(x)
        vs.append(v)
    return (x, xs, vs) if return_trajs else x


def heun_sample(net: Callable, noise, cond, uncond, num_steps: int, guidance: float,
                gmin: float = 0.0, gmax: float = 1.0, timeshift: float = 1.0,
                last_step: Optional[float] = None, exact_henu: bool = False):
    """HeunSampler._impl_sampling, ODE step (flow_matching/sampling.py:230-296).
    With exact_henu=False the corrector's velocity is reused as the next predictor."""
    steps = make_timesteps(num_steps, timeshift, last_step).to(noise.device)
    B = noise.shape[0]
    cfg_c = torch.cat([uncond, cond], dim=0)
    x = noise
    v_hat = None
    for i, (t_cur, t_next) in enumerate(zip(steps[:-1], steps[1:])):
        dt = t_next - t_cur
        g = guidance if (t_cur > gmin and t_cur <= gmax) else 1.0
        if i == 0 or exact_henu:
            v = cfg_combine(net(torch.cat([x, x], 0), t_cur.repeat(2 * B), cfg_c), g)
        else:
            v = v_hat
        x_hat = x + v * dt
        if i < num_steps - 1:
            v_hat = cfg_combine(net(torch.cat([x_hat, x_hat], 0), t_next.repeat(2 * B), cfg_c), g)
            v = (v + v_hat) / 2
            x = x + v * dt
        else:
            x = x + v * dt
    return x


def lagrange_coeffs(order: int, ts: Sequence[float], t0: float, t1: float) -> Tuple[float, ...]:
    """Normalised integrals of the Lagrange basis over [t0, t1] for the last `order` nodes of ts
    (pre_integral.py:4-125, orders 1-4; evaluated here in closed form with float64 polynomials)."""
    import numpy as np
    order = min(order, len(ts))
    nodes = [float(v) for v in ts[-order:]]
    if order == 1:
        return (1.0,)
    ints = []
    for j, tj in enumerate(nodes):
        poly = np.poly1d([1.0])
        den = 1.0
        for m, tm in enumerate(nodes):
            if m != j:
                poly = poly * np.poly1d([1.0, -tm])
                den *= (tj - tm)
        ip = poly.integ()
        ints.append((ip(t1) - ip(t0)) / den)
    tot = sum(ints)
    return tuple(v / tot for v in ints)


def adam_coeffs(num_steps: int, order: int, timeshift: float, last_step: Optional[float] = None):
    """AdamLMSampler.__init__/_reparameterize_coeffs (adam_sampling.py:60-84)."""
    if last_step is None:
        last_step = 1.0 / num_steps
    ts = torch.linspace(0.0, 1 - last_step, num_steps)
    ts = torch.cat([ts, torch.tensor([1.0])], dim=0)
    ts = shift_respace(ts, timeshift)
    deltas = ts[1:] - ts[:-1]
    coeffs = []
    for i in range(num_steps):
        o = min(order, i + 1)
        coeffs.append(lagrange_coeffs(o, [float(v) for v in ts[: i + 1]], float(ts[i]), float(ts[i + 1])))
    return ts, deltas, coeffs


def adam_sample(net: Callable, noise, cond, uncond, num_steps: int, guidance: float, order: int = 2,
                gmin: float = 0.0, gmax: float = 1.0, timeshift: float = 1.0):
    """AdamLMSampler._impl_sampling (adam_sampling.py:86-122); strict (gmin, gmax) window (:104)."""
    ts, deltas, coeffs = adam_coeffs(num_steps, order, timeshift)
    B = noise.shape[0]
    cfg_c = torch.cat([uncond, cond], dim=0)
    x = noise
    preds: List[torch.Tensor] = []
    t_cur = torch.zeros([B]).to(noise.device, noise.dtype)
    for i in range(num_steps):
        out = net(torch.cat([x, x], 0), t_cur.repeat(2), cfg_c)
        g = guidance if (t_cur[0] > gmin and t_cur[0] < gmax) else 1.0
        preds.append(cfg_combine(out, g))
        o = len(coeffs[i])
        v = torch.zeros_like(preds[-1])
        for j in range(o):
            v = v + coeffs[i][j] * preds[-o:][j]
        x = x + v * deltas[i].to(x.device)
        t_cur = t_cur + deltas[i].to(x.device)
    return x


def fp2uint8(x: torch.Tensor) -> torch.Tensor:
    """src/models/autoencoder/base.py:32-34."""
    return torch.clamp((x + 1) * 127.5 + 0.5, 0, 255).to(torch.uint8)


# --------------------------------------------------------------------------- DCT / FM loss
JPEG_LUMA = [
    [16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55],
    [14, 13, 16, 24, 40, 57, 69, 56], [14, 17, 22, 29, 51, 87, 80, 62],
    [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
    [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]]
JPEG_CHROMA = [
    [17, 18, 24, 47, 99, 99, 99, 99], [18, 21, 26, 66, 99, 99, 99, 99],
    [24, 26, 56, 99, 99, 99, 99, 99], [47, 66, 99, 99, 99, 99, 99, 99],
    [99] * 8, [99] * 8, [99] * 8, [99] * 8]


def dct_matrix(n: int = 8) -> torch.Tensor:
    """Orthonormal DCT-II matrix (training_repa_DeCo.py:95-104)."""
    i = torch.arange(n, dtype=torch.float32)
    k = i.unsqueeze(1)
    C = torch.cos(math.pi * (2 * i + 1) * k / (2.0 * n))
    a = torch.sqrt(torch.tensor(2.0) / n) * torch.ones(n)
    a[0] = math.sqrt(1.0 / n)
    return a.unsqueeze(1) * C


def freq_weight(quality: int = 85, mode: str = "inv_gamma", gamma: float = 1.0) -> torch.Tensor:
    """JPEG-table frequency weights [3,8,8] (training_repa_DeCo.py:138-195)."""
    def scale_q(base):
        q = max(1, min(100, int(quality)))
        sc = 5000 / q if q < 50 else 200 - 2 * q
        return torch.floor((torch.tensor(base, dtype=torch.float32) * sc + 50) / 100).clamp(1, 255)

    def to_w(Q):
        if mode == "inv":
            w = 1.0 / Q
        elif mode == "inv_gamma":
            w = (Q.mean() / Q) ** gamma
        else:
            raise ValueError("mode must be 'inv' or 'inv_gamma'")
        return w / w.mean()
    wy, wc = to_w(scale_q(JPEG_LUMA)), to_w(scale_q(JPEG_CHROMA))
    return torch.stack([wy, wc, wc], dim=0)


def rgb2ycbcr(x: torch.Tensor) -> torch.Tensor:
    """BT.601 full range without offsets (training_repa_DeCo.py:106-114)."""
    r, g, b = x[:, 0:1], x[:, 1:2], x[:, 2:3]
    return torch.cat([0.299 * r + 0.587 * g + 0.114 * b,
                      -0.168736 * r - 0.331264 * g + 0.5 * b,
                      0.5 * r - 0.418688 * g - 0.081312 * b], dim=1)


def block_dct(x: torch.Tensor, bs: int = 8) -> torch.Tensor:
    """8x8 block DCT -> [B,C,Bh,Bw,8,8], reflect padding for ragged sizes
    (training_repa_DeCo.py:116-136)."""
    B, C, H, W = x.shape
    ph, pw = (-H) % bs, (-W) % bs
    if ph or pw:
        x = F.pad(x, (0, pw, 0, ph), mode="reflect")
    B, C, H2, W2 = x.shape
    blocks = x.unfold(2, bs, bs).unfold(3, bs, bs).contiguous().view(-1, bs, bs)
    Cm = dct_matrix(bs).to(x.device, x.dtype)
    d = torch.matmul(torch.matmul(Cm.unsqueeze(0), blocks), Cm.t().unsqueeze(0))
    return d.view(B, C, H2 // bs, W2 // bs, bs, bs)


def dct_fm_loss(out: torch.Tensor, v_t: torch.Tensor, freq_loss_weight: float = 1.0,
                quality: int = 85, mode: str = "inv_gamma", gamma: float = 1.0) -> Dict[str, torch.Tensor]:
    """loss = mean((out-v_t)^2) + freq_loss_weight * mean(freq_w * (dct(ycbcr(out)) - dct(ycbcr(v_t)))^2)
    (training_repa_DeCo.py:273-285 and the original trainer)."""
    fm = ((out - v_t) ** 2).mean()
    w = freq_weight(quality, mode, gamma).to(out.device)[None, :, None, None]
    fr = (w * (block_dct(rgb2ycbcr(out)) - block_dct(rgb2ycbcr(v_t))) ** 2).mean()
    return dict(fm_loss=fm, fm_loss_freq=fr, loss=fm + freq_loss_weight * fr)


def time_shift(t, timeshift: float = 1.0):
    """training_repa_DeCo.py:39-40."""
    return t / (t + (1 - t) * timeshift)


def make_xt_vt(x: torch.Tensor, noise: torch.Tensor, t: torch.Tensor):
    """LinearScheduler interpolation: x_t = t*x + (1-t)*eps, v_t = x - eps
    (training_repa_DeCo.py:231-237, scheduling.py:6-14)."""
    a = t.view(-1, 1, 1, 1)
    return a * x + (1 - a) * noise, x - noise


def label_dropout(condition: torch.Tensor, uncondition: torch.Tensor, null_condition_p: float) -> torch.Tensor:
    """BaseTrainer.preproprocess (src/diffusion/base/training.py:14-20): one torch.rand(bsz) draw on the condition's device,
    rows with u < p take the null condition."""
    if null_condition_p <= 0:
        return condition
    bsz = condition.shape[0]
    mask = torch.rand((bsz), device=condition.device) < null_condition_p
    mask = mask.view(-1, *([1] * (len(condition.shape) - 1))).to(condition.dtype)
    return condition * (1 - mask) + uncondition * mask


def trainstep_inputs(x: torch.Tensor, timeshift: float = 1.0):
    """REPATrainer._impl_trainstep up to the network call (training_repa_DeCo.py:222-237), LinearScheduler
    (scheduling.py:6-14: alpha = t, sigma = 1 - t, dalpha = 1, dsigma = -1), drawing from torch's global generator of
    x's device in the reference's order: randn(B) -> rand(B) [uniform t] -> rand(B) [90/10 selector] -> randn_like(x).
    Returns (t [B], x_t, v_t)."""
    B = x.shape[0]
    nt = torch.randn((B,), device=x.device, dtype=torch.float32)
    t_lognorm = torch.sigmoid(nt)
    t_uniform = torch.rand((B,), device=x.device, dtype=torch.float32)
    base_t = torch.where(torch.rand((B,), device=x.device) <= 0.9, t_lognorm, t_uniform)
    t = time_shift(base_t, timeshift)
    noise = torch.randn_like(x)
    a = t.view(-1, 1, 1, 1)
    alpha, sigma = a, 1.0 - a
    dalpha, dsigma = torch.full_like(a, 1.0), torch.full_like(a, -1.0)
    return t, alpha * x + noise * sigma, dalpha * x + dsigma * noise


def trainstep(net: Callable, x: torch.Tensor, condition: torch.Tensor, uncondition: torch.Tensor,
              null_condition_p: float = 0.1, timeshift: float = 1.0, freq_loss_weight: float = 0.0) -> Dict[str, torch.Tensor]:
    """BaseTrainer.__call__ + REPATrainer._impl_trainstep (base/training.py:25-28, training_repa_DeCo.py:216-288).
    freq_loss_weight = 0 is the objective of the checked-in fork (loss = fm_loss.mean(), :276-288); > 0 adds the block-DCT
    term of the original trainer (the formula kept in the comment at :276-285)."""
    y = label_dropout(condition, uncondition, null_condition_p)
    t, x_t, v_t = trainstep_inputs(x, timeshift)
    out = net(x_t, t, y)
    if freq_loss_weight:
        return dct_fm_loss(out.float(), v_t, freq_loss_weight)
    fm = ((out.float() - v_t) ** 2).mean()
    return dict(fm_loss=fm, loss=fm)


# --------------------------------------------------------------------------- text-to-image denoiser (config 5)
# The original `src/models/transformer/dit_t2i_DeCo.py` survives only as CPython-3.10 bytecode in the checkout
# (SURVEY.md 8c); its encoder is the logic of `src/models/transformer/dit_t2i_pixnerd.py` (Attention :16-63,
# FlattenDiTBlock :65-81, NerfEmbedder :83-108, TextRefineAttention/Block :144-198, forward :276-297) and its
# decoder is SimpleMLPAdaLN of `dit_c2i_DeCo.py:288-415`.  tests/golden/make_golden.py composes exactly those
# importable reference classes into a module and pins `t2i_forward` against it.
@dataclass(frozen=True)
class T2ICfg:
    """Constructor arguments of the text-to-image denoiser (configs_t2i/sft_res512.yaml:45-56)."""
    in_channels: int = 3
    num_groups: int = 24
    hidden_size: int = 1536
    decoder_hidden_size: int = 32
    num_encoder_blocks: int = 16
    num_decoder_blocks: int = 3
    num_text_blocks: int = 4
    patch_size: int = 16
    txt_embed_dim: int = 2048
    txt_max_length: int = 128
    max_freqs: int = 8

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_groups

    @property
    def ffn_hidden(self) -> int:
        # FlattenDiTBlock: mlp_hidden_dim = int(hidden * 4); SwiGLU keeps it un-scaled (layers/swiglu.py:11-12)
        return int(self.hidden_size * 4)


CFG_XXL_T2I = T2ICfg()


def t2i_param_shapes(cfg: T2ICfg) -> Dict[str, Tuple[int, ...]]:
    """state_dict names/shapes of the original t2i module (members per the 3.10 bytecode: s_embedder, x_embedder,
    t_embedder, y_embedder, y_pos_embedding, blocks, dec_net, text_refine_blocks)."""
    H, Hx, p, C = cfg.hidden_size, cfg.decoder_hidden_size, cfg.patch_size, cfg.in_channels
    d, ffn = cfg.head_dim, cfg.ffn_hidden
    s: Dict[str, Tuple[int, ...]] = {
        "x_embedder.embedder.0.weight": (Hx, C + cfg.max_freqs ** 2),
        "x_embedder.embedder.0.bias": (Hx,),
        "s_embedder.proj.weight": (H, C * p * p),
        "s_embedder.proj.bias": (H,),
        "t_embedder.mlp.0.weight": (H, 256),
        "t_embedder.mlp.0.bias": (H,),
        "t_embedder.mlp.2.weight": (H, H),
        "t_embedder.mlp.2.bias": (H,),
        "y_embedder.proj.weight": (H, cfg.txt_embed_dim),
        "y_embedder.proj.bias": (H,),
        "y_embedder.norm.weight": (H,),
        "y_pos_embedding": (1, cfg.txt_max_length, H),
    }

    def block(b, joint):
        s[b + "norm1.weight"] = (H,)
        if joint:
            s[b + "attn.qkv_x.weight"] = (3 * H, H)
            s[b + "attn.kv_y.weight"] = (2 * H, H)
        else:
            s[b + "attn.qkv.weight"] = (3 * H, H)
        s[b + "attn.q_norm.weight"] = (d,)
        s[b + "attn.k_norm.weight"] = (d,)
        s[b + "attn.proj.weight"] = (H, H)
        s[b + "attn.proj.bias"] = (H,)
        s[b + "norm2.weight"] = (H,)
        s[b + "mlp.w12.weight"] = (2 * ffn, H)
        s[b + "mlp.w3.weight"] = (H, ffn)
        s[b + "adaLN_modulation.0.weight"] = (6 * H, H)
        s[b + "adaLN_modulation.0.bias"] = (6 * H,)

    for i in range(cfg.num_encoder_blocks):
        block(f"blocks.{i}.", True)
    for i in range(cfg.num_text_blocks):
        block(f"text_refine_blocks.{i}.", False)
    s["dec_net.cond_embed.weight"] = (p * p * Hx, H)
    s["dec_net.cond_embed.bias"] = (p * p * Hx,)
    s["dec_net.input_proj.weight"] = (Hx, Hx)
    s["dec_net.input_proj.bias"] = (Hx,)
    for j in range(cfg.num_decoder_blocks):
        b = f"dec_net.res_blocks.{j}."
        s[b + "in_ln.weight"] = (Hx,)
        s[b + "in_ln.bias"] = (Hx,)
        s[b + "mlp.0.weight"] = (Hx, Hx)
        s[b + "mlp.0.bias"] = (Hx,)
        s[b + "mlp.2.weight"] = (Hx, Hx)
        s[b + "mlp.2.bias"] = (Hx,)
        s[b + "adaLN_modulation.1.weight"] = (3 * Hx, Hx)
        s[b + "adaLN_modulation.1.bias"] = (3 * Hx,)
    s["dec_net.final_layer.linear.weight"] = (C, Hx)
    s["dec_net.final_layer.linear.bias"] = (C,)
    return s


def t2i_seeded_params(cfg: T2ICfg, seed: int = 4321, device="cpu") -> Params:
    """Deterministic fully non-zero weights keyed by name (same recipe as `seeded_params`)."""
    out: Params = {}
    for idx, (name, shape) in enumerate(sorted(t2i_param_shapes(cfg).items())):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if name == "y_pos_embedding":
            w = 0.5 * torch.randn(shape, generator=g)
        elif len(shape) == 2:
            std = 1.0 / math.sqrt(shape[1])
            if "adaLN_modulation" in name:
                std *= 0.5
            w = torch.randn(shape, generator=g) * std
        elif name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight") \
                or name.endswith("in_ln.weight"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            w = 0.05 * torch.randn(shape, generator=g)
        out[name] = w.to(device)
    return out


def rope_table_ex2d(head_dim: int, height: int, width: int, theta: float = 10000.0, scale: float = 1.0) -> torch.Tensor:
    """Angles of precompute_freqs_cis_ex2d (layers/rope.py:22-37): positions linspace(0, height*scale, width) for x and
    linspace(0, width*scale, height) for y (sic); pair 2k -> x, 2k+1 -> y.  Returned as real angles [L, head_dim/2]."""
    x_pos = torch.linspace(0, height * scale, width)
    y_pos = torch.linspace(0, width * scale, height)
    y_pos, x_pos = torch.meshgrid(y_pos, x_pos, indexing="ij")
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 4)[: head_dim // 4].float() / head_dim))
    xa = torch.outer(x_pos.reshape(-1), freqs).float()
    ya = torch.outer(y_pos.reshape(-1), freqs).float()
    return torch.stack([xa, ya], dim=-1).reshape(height * width, -1)


def t2i_nerf_pos_table(patch_size: int, max_freqs: int = 8) -> torch.Tensor:
    """NerfEmbedder.fetch_pos of the t2i model (dit_t2i_pixnerd.py:92-96): the complex ex2d table of dim
    2*max_freqs^2 cast to a real dtype, i.e. its real part cos(angle): [p*p, max_freqs^2]."""
    return torch.cos(rope_table_ex2d(max_freqs ** 2 * 2, patch_size, patch_size))


def _swiglu12(P: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    """layers/swiglu.py:15-17."""
    x1, x2 = F.linear(x, P[pre + "w12.weight"]).chunk(2, dim=-1)
    return F.linear(F.silu(x1) * x2, P[pre + "w3.weight"])


def _heads(x: torch.Tensor, n: int, heads: int) -> Tuple[torch.Tensor, ...]:
    B, N, C = x.shape
    return tuple(x.reshape(B, N, n, heads, C // n // heads).permute(2, 0, 3, 1, 4))   # n x [B, heads, N, d]


def _rope_bhnd(x: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    """apply_rotary_emb in the [B, heads, N, d] layout (layers/rope.py:40-51)."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    cos, sin = torch.cos(angles)[None, None], torch.sin(angles)[None, None]
    a, b = xf[..., 0], xf[..., 1]
    return torch.stack([a * cos - b * sin, a * sin + b * cos], dim=-1).flatten(3).type_as(x)


def t2i_joint_attention(P: Params, pre: str, x, y, angles, heads: int) -> torch.Tensor:
    """Attention.forward (dit_t2i_pixnerd.py:41-63): image queries over [image || text] keys; k_norm is shared by
    the image and the text keys, RoPE touches the image q/k only."""
    B, N, C = x.shape
    q, kx, vx = _heads(F.linear(x, P[pre + "qkv_x.weight"]), 3, heads)
    q = rmsnorm(q, P[pre + "q_norm.weight"])
    kx = rmsnorm(kx, P[pre + "k_norm.weight"])
    q, kx = _rope_bhnd(q, angles), _rope_bhnd(kx, angles)
    ky, vy = _heads(F.linear(y, P[pre + "kv_y.weight"]), 2, heads)
    ky = rmsnorm(ky, P[pre + "k_norm.weight"])
    k, v = torch.cat([kx, ky], dim=2), torch.cat([vx, vy], dim=2)
    o = F.scaled_dot_product_attention(q, k, v)
    return F.linear(o.transpose(1, 2).reshape(B, N, C), P[pre + "proj.weight"], P[pre + "proj.bias"])


def t2i_text_attention(P: Params, pre: str, x, heads: int) -> torch.Tensor:
    """TextRefineAttention.forward (dit_t2i_pixnerd.py:166-179): q/k-normed self-attention, no RoPE."""
    B, N, C = x.shape
    q, k, v = _heads(F.linear(x, P[pre + "qkv.weight"]), 3, heads)
    q, k = rmsnorm(q, P[pre + "q_norm.weight"]), rmsnorm(k, P[pre + "k_norm.weight"])
    o = F.scaled_dot_product_attention(q, k, v)
    return F.linear(o.transpose(1, 2).reshape(B, N, C), P[pre + "proj.weight"], P[pre + "proj.bias"])


def t2i_forward(P: Params, cfg: T2ICfg, x: torch.Tensor, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Original t2i PixNerDiT.forward: dit_t2i_pixnerd.py:276-297 up to x_embedder, then dec_net + fold as
    dit_c2i_DeCo.py:501-509.  x:[B,C,H,W], t:[B], y:[B, T, txt_embed_dim]."""
    B, _, Hh, Ww = x.shape
    p, H, heads = cfg.patch_size, cfg.hidden_size, cfg.num_groups
    xp = F.unfold(x, kernel_size=p, stride=p).transpose(1, 2)
    angles = rope_table_ex2d(cfg.head_dim, Hh // p, Ww // p).to(x.device)
    tf = timestep_embedding(t.view(-1)).to(t.dtype)
    te = F.linear(F.silu(F.linear(tf, P["t_embedder.mlp.0.weight"], P["t_embedder.mlp.0.bias"])),
                  P["t_embedder.mlp.2.weight"], P["t_embedder.mlp.2.bias"]).view(B, -1, H)
    ye = rmsnorm(F.linear(y, P["y_embedder.proj.weight"], P["y_embedder.proj.bias"]), P["y_embedder.norm.weight"])
    ye = ye.view(B, -1, H) + P["y_pos_embedding"].to(y.dtype)
    c = F.silu(te)

    def mods(b):
        return F.linear(c, P[b + "adaLN_modulation.0.weight"], P[b + "adaLN_modulation.0.bias"]).chunk(6, dim=-1)

    for i in range(cfg.num_text_blocks):
        b = f"text_refine_blocks.{i}."
        sh1, sc1, g1, sh2, sc2, g2 = mods(b)
        ye = ye + g1 * t2i_text_attention(P, b + "attn.", modulate(rmsnorm(ye, P[b + "norm1.weight"]), sh1, sc1), heads)
        ye = ye + g2 * _swiglu12(P, b + "mlp.", modulate(rmsnorm(ye, P[b + "norm2.weight"]), sh2, sc2))
    s = F.linear(xp, P["s_embedder.proj.weight"], P["s_embedder.proj.bias"])
    for i in range(cfg.num_encoder_blocks):
        b = f"blocks.{i}."
        sh1, sc1, g1, sh2, sc2, g2 = mods(b)
        s = s + g1 * t2i_joint_attention(P, b + "attn.", modulate(rmsnorm(s, P[b + "norm1.weight"]), sh1, sc1),
                                         ye, angles, heads)
        s = s + g2 * _swiglu12(P, b + "mlp.", modulate(rmsnorm(s, P[b + "norm2.weight"]), sh2, sc2))
    s = F.silu(te + s)
    L = s.shape[1]
    px = xp.reshape(B * L, cfg.in_channels, p * p).transpose(1, 2)
    tab = t2i_nerf_pos_table(p, cfg.max_freqs).to(device=px.device, dtype=px.dtype)
    px = F.linear(torch.cat([px, tab[None].expand(B * L, -1, -1)], dim=-1),
                  P["x_embedder.embedder.0.weight"], P["x_embedder.embedder.0.bias"])
    dcfg = DenoiserCfg(in_channels=cfg.in_channels, hidden_size=H, hidden_size_x=cfg.decoder_hidden_size,
                       num_blocks=cfg.num_decoder_blocks, num_cond_blocks=0, patch_size=p)
    out = pixel_decoder(P, dcfg, px, s.reshape(B * L, H))
    out = out.transpose(1, 2).reshape(B, L, -1)
    return F.fold(out.transpose(1, 2).contiguous(), (Hh, Ww), kernel_size=p, stride=p)


# --------------------------------------------------------------------------- patch-linear baseline denoiser
@dataclass(frozen=True)
class BaselineCfg:
    """Constructor arguments of FlattenDiT (src/models/transformer/dit_c2i_baseline.py:290-303;
    configs_c2i/Baseline_DiT_JiT.yaml:46-54)."""
    in_channels: int = 3
    num_groups: int = 16
    hidden_size: int = 1024
    num_blocks: int = 24
    patch_size: int = 16
    num_classes: int = 1000

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_groups

    @property
    def ffn_hidden(self) -> int:
        return int(2 * int(self.hidden_size * 4.0) / 3)      # dit_c2i_baseline.py:200, :108


def baseline_param_shapes(cfg: BaselineCfg) -> Dict[str, Tuple[int, ...]]:
    """state_dict names and shapes of FlattenDiT (dit_c2i_baseline.py:312-322)."""
    H, p, C = cfg.hidden_size, cfg.patch_size, cfg.in_channels
    d, ffn = cfg.head_dim, cfg.ffn_hidden
    s: Dict[str, Tuple[int, ...]] = {
        "x_embedder.proj.weight": (H, C * p * p),
        "x_embedder.proj.bias": (H,),
        "t_embedder.mlp.0.weight": (H, 256),
        "t_embedder.mlp.0.bias": (H,),
        "t_embedder.mlp.2.weight": (H, H),
        "t_embedder.mlp.2.bias": (H,),
        "y_embedder.embedding_table.weight": (cfg.num_classes + 1, H),
        "final_layer.linear.weight": (C * p * p, H),
        "final_layer.linear.bias": (C * p * p,),
        "final_layer.adaLN_modulation.0.weight": (2 * H, H),
        "final_layer.adaLN_modulation.0.bias": (2 * H,),
    }
    for i in range(cfg.num_blocks):
        b = f"blocks.{i}."
        s[b + "norm1.weight"] = (H,)
        s[b + "attn.qkv.weight"] = (3 * H, H)
        s[b + "attn.q_norm.weight"] = (d,)
        s[b + "attn.k_norm.weight"] = (d,)
        s[b + "attn.proj.weight"] = (H, H)
        s[b + "attn.proj.bias"] = (H,)
        s[b + "norm2.weight"] = (H,)
        s[b + "mlp.w1.weight"] = (ffn, H)
        s[b + "mlp.w3.weight"] = (ffn, H)
        s[b + "mlp.w2.weight"] = (H, ffn)
        s[b + "adaLN_modulation.0.weight"] = (6 * H, H)
        s[b + "adaLN_modulation.0.bias"] = (6 * H,)
    return s


def baseline_seeded_params(cfg: BaselineCfg, seed: int = 2468, device="cpu") -> Params:
    """Fully non-zero seeded weights (the default init zeroes the whole final layer, dit_c2i_baseline.py:351-355)."""
    out: Params = {}
    for idx, (name, shape) in enumerate(sorted(baseline_param_shapes(cfg).items())):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if name.startswith("y_embedder"):
            w = torch.randn(shape, generator=g) * 0.5
        elif len(shape) == 2:
            std = 1.0 / math.sqrt(shape[1])
            if "adaLN_modulation" in name:
                std *= 0.5
            w = torch.randn(shape, generator=g) * std
        elif name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            w = 0.05 * torch.randn(shape, generator=g)
        out[name] = w.to(device)
    return out


def baseline_forward(P: Params, cfg: BaselineCfg, x: torch.Tensor, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """FlattenDiT.forward (dit_c2i_baseline.py:357-379) with FinalLayer (:70-83); the blocks are `dit_block`
    (dit_c2i_baseline.py:194-210 is the module of dit_c2i_DeCo.py:194-210)."""
    B, _, Hh, Ww = x.shape
    p, H = cfg.patch_size, cfg.hidden_size
    angles = rope_table_2d(cfg.head_dim, Hh // p, Ww // p).to(x.device)
    h = F.unfold(x, kernel_size=p, stride=p).transpose(1, 2)
    h = F.linear(h, P["x_embedder.proj.weight"], P["x_embedder.proj.bias"])
    tf = timestep_embedding(t.view(-1))
    te = F.linear(F.silu(F.linear(tf, P["t_embedder.mlp.0.weight"], P["t_embedder.mlp.0.bias"])),
                  P["t_embedder.mlp.2.weight"], P["t_embedder.mlp.2.bias"]).view(B, -1, H)
    ye = F.embedding(y, P["y_embedder.embedding_table.weight"]).view(B, 1, H)
    c = F.silu(te + ye)
    for i in range(cfg.num_blocks):
        h = dit_block(P, i, h, c, angles, cfg.num_groups)
    shift, scale = F.linear(c, P["final_layer.adaLN_modulation.0.weight"],
                            P["final_layer.adaLN_modulation.0.bias"]).chunk(2, dim=-1)
    h = modulate(F.layer_norm(h, (H,), None, None, 1e-6), shift, scale)
    h = F.linear(h, P["final_layer.linear.weight"], P["final_layer.linear.bias"])
    return F.fold(h.transpose(1, 2).contiguous(), (Hh, Ww), kernel_size=p, stride=p)


# --------------------------------------------------------------------------- Euler variants: x-prediction and SDE steps
def linear_score_terms(t: torch.Tensor):
    """LinearScheduler (flow_matching/scheduling.py:6-14): returns (1/dalpha_over_alpha, sigma, dsigma_mul_sigma) at
    scalar t, as the sampler evaluates them (sampling.py:81-83)."""
    alpha, sigma = t, 1 - t
    return 1 / (1.0 / alpha), sigma, -1.0 * sigma


def euler_sample_ex(net: Callable, noise: torch.Tensor, cond: torch.Tensor, uncond: torch.Tensor, num_steps: int,
                    guidance: float, gmin: float = 0.0, gmax: float = 1.0, timeshift: float = 1.0,
                    last_step: Optional[float] = None, x_prediction: bool = False, step: str = "ode",
                    last: str = "ode", w_fn: Optional[Callable] = None, randn: Optional[Callable] = None):
    """EulerSampler / EulerSamplerJiT._impl_sampling with LinearScheduler and any step function
    (flow_matching/sampling.py:66-107, :144-188; step functions :14-24).  step / last in {"ode", "sde_mean", "sde",
    "sde_preserve"}; w_fn(t) = w_scheduler.w (default: sigma = 1 - t, base/scheduling.py:31-32); randn(x) supplies the
    Gaussian increments (default torch.randn_like: the reference's own call)."""
    steps = make_timesteps(num_steps, timeshift, last_step).to(noise.device, noise.dtype)
    B = noise.shape[0]
    cfg_c = torch.cat([uncond, cond], dim=0)
    randn = randn or torch.randn_like
    w_fn = w_fn or (lambda t: 1 - t)
    x = noise
    for i, (t_cur, t_next) in enumerate(zip(steps[:-1], steps[1:])):
        dt = t_next - t_cur
        cfg_x = torch.cat([x, x], 0)
        out = net(cfg_x, t_cur.repeat(2 * B), cfg_c)
        if x_prediction:
            out = (out - cfg_x) / (1.0 - t_cur).clamp_min(5e-2)                      # sampling.py:170
        g = guidance if (t_cur > gmin and t_cur <= gmax) else 1.0
        v = cfg_combine(out, g)
        kd, sigma, dms = linear_score_terms(t_cur)
        s = (kd * v - x) / (sigma ** 2 - kd * dms)                                     # sampling.py:98
        w = w_fn(t_cur)
        kind = step if i < num_steps - 1 else last
        if kind == "ode":
            x = x + v * dt
        elif kind == "sde_mean":
            x = x + v * dt + s * w * dt
        elif kind == "sde":
            x = x + v * dt + s * w * dt + torch.sqrt(2 * w * dt) * randn(x)
        elif kind == "sde_preserve":
            x = x + v * dt + 0.5 * s * w * dt + torch.sqrt(w * dt) * randn(x)
        else:
            raise ValueError(kind)
    return x


def _sde_step(kind: str, x, v, dt, s, w, randn):
    """Step functions of flow_matching/sampling.py:14-24."""
    if kind == "ode":
        return x + v * dt
    if kind == "sde_mean":
        return x + v * dt + s * w * dt
    if kind == "sde":
        return x + v * dt + s * w * dt + torch.sqrt(2 * w * dt) * randn(x)
    if kind == "sde_preserve":
        return x + v * dt + 0.5 * s * w * dt + torch.sqrt(w * dt) * randn(x)
    raise ValueError(kind)


def heun_sample_ex(net: Callable, noise, cond, uncond, num_steps: int, guidance: float, gmin: float = 0.0,
                   gmax: float = 1.0, timeshift: float = 1.0, last_step: Optional[float] = None, exact_henu: bool = False,
                   step: str = "ode", last: str = "ode", w_fn: Optional[Callable] = None, randn: Optional[Callable] = None):
    """HeunSampler._impl_sampling with LinearScheduler and any step function (flow_matching/sampling.py:230-296): the
    score s = (kd v - x) / (sigma^2 - kd dsigma sigma) is evaluated at (x, t_cur) and at (x_hat, t_next) and AVERAGED like
    the velocity (:283-291); with exact_henu=False both are re-used as the next predictor (:273-275); the last step
    re-applies last_step_fn to x (not x_hat) with the predictor's (v, s) (:293); the guidance window of both evaluations is
    tested on t_cur (:263, :280).  Every step_fn call with a noise term draws one randn_like(x) -- also the predictor of the
    last step, whose x_hat is discarded."""
    steps = make_timesteps(num_steps, timeshift, last_step).to(noise.device, noise.dtype)
    B = noise.shape[0]
    cfg_c = torch.cat([uncond, cond], dim=0)
    randn = randn or torch.randn_like
    w_fn = w_fn or (lambda t: 1 - t)
    x = noise
    v_hat = s_hat = None
    for i, (t_cur, t_next) in enumerate(zip(steps[:-1], steps[1:])):
        dt = t_next - t_cur
        kd, sigma, dms = linear_score_terms(t_cur)
        kdh, sigmah, dmsh = linear_score_terms(t_next)
        w = w_fn(t_cur)
        g = guidance if (t_cur > gmin and t_cur <= gmax) else 1.0
        if i == 0 or exact_henu:
            v = cfg_combine(net(torch.cat([x, x], 0), t_cur.repeat(2 * B), cfg_c), g)
            s = (kd * v - x) / (sigma ** 2 - kd * dms)
        else:
            v, s = v_hat, s_hat
        x_hat = _sde_step(step, x, v, dt, s, w, randn)
        if i < num_steps - 1:
            v_hat = cfg_combine(net(torch.cat([x_hat, x_hat], 0), t_next.repeat(2 * B), cfg_c), g)
            s_hat = (kdh * v_hat - x_hat) / (sigmah ** 2 - kdh * dmsh)
            v = (v + v_hat) / 2
            s = (s + s_hat) / 2
            x = _sde_step(step, x, v, dt, s, w, randn)
        else:
            x = _sde_step(last, x, v, dt, s, w, randn)
    return x


# --------------------------------------------------------------------------- PixNerd baseline (hyper-network decoder)
# Restated ahead of its CUDA path (DESIGN.md section 8, rank-4 leftovers): configs_c2i/Baseline_PixNerd.yaml.
@dataclass(frozen=True)
class PixNerdCfg:
    """Constructor arguments of dit_c2i_pixnerd.PixNerDiT (src/models/transformer/dit_c2i_pixnerd.py:288-303;
    configs_c2i/Baseline_PixNerd.yaml:44-55)."""
    in_channels: int = 3
    num_groups: int = 16
    hidden_size: int = 1024
    hidden_size_x: int = 64
    nerf_mlpratio: int = 2
    num_blocks: int = 24
    num_cond_blocks: int = 22
    patch_size: int = 16
    num_classes: int = 1000
    max_freqs: int = 8

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_groups

    @property
    def ffn_hidden(self) -> int:
        return int(2 * int(self.hidden_size * 4.0) / 3)


def pixnerd_param_shapes(cfg: PixNerdCfg) -> Dict[str, Tuple[int, ...]]:
    """state_dict names and shapes of dit_c2i_pixnerd.PixNerDiT: DiT blocks 0..num_cond_blocks-1 and NerfBlocks
    num_cond_blocks..num_blocks-1 share ONE ModuleList (dit_c2i_pixnerd.py:325-330)."""
    H, Hx, p, C = cfg.hidden_size, cfg.hidden_size_x, cfg.patch_size, cfg.in_channels
    d, ffn = cfg.head_dim, cfg.ffn_hidden
    s: Dict[str, Tuple[int, ...]] = {
        "x_embedder.embedder.0.weight": (Hx, C + cfg.max_freqs ** 2),
        "x_embedder.embedder.0.bias": (Hx,),
        "s_embedder.proj.weight": (H, C * p * p),
        "s_embedder.proj.bias": (H,),
        "t_embedder.mlp.0.weight": (H, 256),
        "t_embedder.mlp.0.bias": (H,),
        "t_embedder.mlp.2.weight": (H, H),
        "t_embedder.mlp.2.bias": (H,),
        "y_embedder.embedding_table.weight": (cfg.num_classes + 1, H),
        "final_layer.norm.weight": (Hx,),
        "final_layer.linear.weight": (C, Hx),
        "final_layer.linear.bias": (C,),
    }
    for i in range(cfg.num_cond_blocks):
        b = f"blocks.{i}."
        s[b + "norm1.weight"] = (H,)
        s[b + "attn.qkv.weight"] = (3 * H, H)
        s[b + "attn.q_norm.weight"] = (d,)
        s[b + "attn.k_norm.weight"] = (d,)
        s[b + "attn.proj.weight"] = (H, H)
        s[b + "attn.proj.bias"] = (H,)
        s[b + "norm2.weight"] = (H,)
        s[b + "mlp.w1.weight"] = (ffn, H)
        s[b + "mlp.w3.weight"] = (ffn, H)
        s[b + "mlp.w2.weight"] = (H, ffn)
        s[b + "adaLN_modulation.0.weight"] = (6 * H, H)
        s[b + "adaLN_modulation.0.bias"] = (6 * H,)
    for i in range(cfg.num_cond_blocks, cfg.num_blocks):
        b = f"blocks.{i}."
        s[b + "param_generator1.0.weight"] = (2 * Hx * Hx * cfg.nerf_mlpratio, H)
        s[b + "param_generator1.0.bias"] = (2 * Hx * Hx * cfg.nerf_mlpratio,)
        s[b + "norm.weight"] = (Hx,)
    return s


def pixnerd_seeded_params(cfg: PixNerdCfg, seed: int = 1357, device="cpu") -> Params:
    """Fully non-zero seeded weights (the default init zeroes final_layer.linear, dit_c2i_pixnerd.py:353-355)."""
    out: Params = {}
    for idx, (name, shape) in enumerate(sorted(pixnerd_param_shapes(cfg).items())):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if name.startswith("y_embedder"):
            w = torch.randn(shape, generator=g) * 0.5
        elif len(shape) == 2:
            std = 1.0 / math.sqrt(shape[1])
            if "adaLN_modulation" in name:
                std *= 0.5
            w = torch.randn(shape, generator=g) * std
        elif name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            w = 0.05 * torch.randn(shape, generator=g)
        out[name] = w.to(device)
    return out


def nerf_block(P: Params, pre: str, x: torch.Tensor, s: torch.Tensor, mlp_ratio: int) -> torch.Tensor:
    """NerfBlock.forward (dit_c2i_pixnerd.py:250-273): a per-patch MLP whose weights are generated from the patch's DiT
    condition and L2-normalised over their input dimension.  x [BL, p*p, Hx], s [BL, H]."""
    n, _, Hx = x.shape
    params = F.linear(s, P[pre + "param_generator1.0.weight"], P[pre + "param_generator1.0.bias"])
    fc1, fc2 = params.chunk(2, dim=-1)
    fc1 = F.normalize(fc1.view(n, Hx, Hx * mlp_ratio), dim=-2)
    fc2 = F.normalize(fc2.view(n, Hx * mlp_ratio, Hx), dim=-2)
    h = rmsnorm(x, P[pre + "norm.weight"])
    h = torch.bmm(F.silu(torch.bmm(h, fc1)), fc2)
    return h + x


def pixnerd_forward(P: Params, cfg: PixNerdCfg, x: torch.Tensor, t: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """dit_c2i_pixnerd.PixNerDiT.forward (dit_c2i_pixnerd.py:358-381)."""
    B, _, Hh, Ww = x.shape
    p, H = cfg.patch_size, cfg.hidden_size
    angles = rope_table_2d(cfg.head_dim, Hh // p, Ww // p).to(x.device)
    xp = F.unfold(x, kernel_size=p, stride=p).transpose(1, 2)
    tf = timestep_embedding(t.view(-1))
    te = F.linear(F.silu(F.linear(tf, P["t_embedder.mlp.0.weight"], P["t_embedder.mlp.0.bias"])),
                  P["t_embedder.mlp.2.weight"], P["t_embedder.mlp.2.bias"]).view(B, -1, H)
    ye = F.embedding(y, P["y_embedder.embedding_table.weight"]).view(B, 1, H)
    c = F.silu(te + ye)
    s = F.linear(xp, P["s_embedder.proj.weight"], P["s_embedder.proj.bias"])
    for i in range(cfg.num_cond_blocks):
        s = dit_block(P, i, s, c, angles, cfg.num_groups)
    s = F.silu(te + s)
    L = s.shape[1]
    px = xp.reshape(B * L, cfg.in_channels, p * p).transpose(1, 2)
    sf = s.reshape(B * L, H)
    tab = nerf_pos_table(p, cfg.max_freqs).to(device=px.device, dtype=px.dtype)
    h = F.linear(torch.cat([px, tab[None].expand(B * L, -1, -1)], dim=-1),
                 P["x_embedder.embedder.0.weight"], P["x_embedder.embedder.0.bias"])
    for i in range(cfg.num_cond_blocks, cfg.num_blocks):
        h = nerf_block(P, f"blocks.{i}.", h, sf, cfg.nerf_mlpratio)
    h = F.linear(rmsnorm(h, P["final_layer.norm.weight"]), P["final_layer.linear.weight"], P["final_layer.linear.bias"])
    out = h.transpose(1, 2).reshape(B, L, -1)
    return F.fold(out.transpose(1, 2).contiguous(), (Hh, Ww), kernel_size=p, stride=p)
