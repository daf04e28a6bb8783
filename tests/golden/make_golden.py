"""Generate golden fixtures by executing the UNMODIFIED reference (/root/reference) on CPU fp32,
and pin the oracle (oracle/deco_oracle.py) against it in the same run.

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  Every fixture is reproducible from seeds: weights come from
oracle.deco_oracle.seeded_params (name-keyed generators), inputs from torch.Generator seeds.
"""
import os
import sys
import types
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

# training_repa_DeCo.py:3,10 import names from timm only; stub them
timm = types.ModuleType("timm")
timm_data = types.ModuleType("timm.data")
timm_data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
timm_data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
timm.data = timm_data
sys.modules["timm"] = timm
sys.modules["timm.data"] = timm_data

from src.models.transformer.dit_c2i_DeCo import PixNerDiT as RefDiT  # noqa: E402
from src.diffusion.flow_matching.sampling import EulerSampler as RefEuler, HeunSampler as RefHeun, ode_step_fn  # noqa: E402
from src.diffusion.flow_matching.adam_sampling import AdamLMSampler as RefAdam  # noqa: E402
from src.diffusion.flow_matching.scheduling import LinearScheduler as RefSched  # noqa: E402
from src.diffusion.base.guidance import simple_guidance_fn as ref_guidance  # noqa: E402
from src.diffusion.flow_matching.training_repa_DeCo import REPATrainer as RefTrainer  # noqa: E402

from src.models.transformer.dit_c2i_baseline import FlattenDiT as RefFlattenDiT  # noqa: E402
from src.models.transformer.dit_c2i_pixnerd import PixNerDiT as RefPixNerd  # noqa: E402
from src.diffusion.flow_matching.sampling import (EulerSamplerJiT as RefEulerJiT, sde_mean_step_fn, sde_step_fn,  # noqa: E402
                                                  sde_preserve_step_fn)

from oracle import deco_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_grad_enabled(False)


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def build_ref(cfg: O.DenoiserCfg, seed=1234):
    m = RefDiT(in_channels=cfg.in_channels, num_groups=cfg.num_groups, hidden_size=cfg.hidden_size,
               hidden_size_x=cfg.hidden_size_x, num_blocks=cfg.num_blocks,
               num_cond_blocks=cfg.num_cond_blocks, patch_size=cfg.patch_size,
               num_classes=cfg.num_classes)
    P = O.seeded_params(cfg, seed)
    sd = m.state_dict()
    assert set(sd.keys()) == set(P.keys()), set(sd.keys()) ^ set(P.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(P[k].shape), k
    m.load_state_dict(P)
    return m.eval(), P


def seeded_noise(n, shape, seed0=0):
    # src/data/dataset/randn.py:74-75: one CPU generator per sample
    return torch.stack([torch.randn(shape, generator=torch.Generator().manual_seed(seed0 + i),
                                    dtype=torch.float32) for i in range(n)])


def golden_forward(name, cfg, B, res, seed):
    m, P = build_ref(cfg)
    x = seeded_noise(B, (cfg.in_channels, res, res), seed)
    t = torch.linspace(0.05, 0.95, B)
    y = torch.tensor([(7 * i + 3) % (cfg.num_classes + 1) for i in range(B)])
    y[-1] = cfg.num_classes  # null label row
    ref = m(x, t, y)
    ora = O.denoiser_forward(P, cfg, x, t, y)
    e = rel_l2(ora, ref)
    print(f"[{name}] oracle vs reference forward rel-L2 = {e:.3e}")
    assert e < 2e-6, e
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ref_bf = m(x, t, y).float()
    print(f"[{name}] reference bf16-autocast vs fp32 rel-L2 = {rel_l2(ref_bf, ref):.3e} (noise floor)")
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), x=x.numpy(), t=t.numpy(), y=y.numpy(),
                        out=ref.numpy(), cfg=np.array([cfg.in_channels, cfg.num_groups, cfg.hidden_size,
                                                       cfg.hidden_size_x, cfg.num_blocks, cfg.num_cond_blocks,
                                                       cfg.patch_size, cfg.num_classes]),
                        bf16_floor=np.float64(rel_l2(ref_bf, ref)))


def golden_samplers():
    """Sampler control flow pinned with an analytic 'network' so the fixture is tiny."""
    def toy_net(x, t, y):
        # depends on x, t and the label so CFG order [uncond || cond] matters
        return torch.tanh(x * (0.5 + t.view(-1, 1, 1, 1))) - 0.1 * y.view(-1, 1, 1, 1).float()
    noise = seeded_noise(3, (3, 8, 8), 100)
    cond = torch.tensor([1, 2, 3])
    unc = torch.tensor([10, 10, 10])
    out = {}
    sch = RefSched()
    for n, g, lo, hi, shift in [(10, 3.2, 0.1, 1.0, 1.0), (7, 2.0, 0.0, 0.6, 3.0)]:
        s = RefEuler(scheduler=sch, w_scheduler=sch, guidance_fn=ref_guidance, num_steps=n, guidance=g,
                     guidance_interval_min=lo, guidance_interval_max=hi, timeshift=shift, step_fn=ode_step_fn)
        r = s(toy_net, noise, cond, unc)
        o = O.euler_sample(toy_net, noise, cond, unc, n, g, lo, hi, shift)
        assert torch.equal(O.make_timesteps(n, shift), s.timesteps)
        assert rel_l2(o, r) < 1e-6, rel_l2(o, r)
        out[f"euler_{n}"] = r.numpy()
        out[f"euler_{n}_ts"] = s.timesteps.numpy()
        h = RefHeun(scheduler=sch, w_scheduler=sch, guidance_fn=ref_guidance, num_steps=n, guidance=g,
                    guidance_interval_min=lo, guidance_interval_max=hi, timeshift=shift, step_fn=ode_step_fn)
        r = h(toy_net, noise, cond, unc)
        o = O.heun_sample(toy_net, noise, cond, unc, n, g, lo, hi, shift)
        assert rel_l2(o, r) < 1e-6, rel_l2(o, r)
        out[f"heun_{n}"] = r.numpy()
        h = RefHeun(scheduler=sch, w_scheduler=sch, guidance_fn=ref_guidance, num_steps=n, guidance=g,
                    guidance_interval_min=lo, guidance_interval_max=hi, timeshift=shift, step_fn=ode_step_fn,
                    exact_henu=True)
        r = h(toy_net, noise, cond, unc)
        o = O.heun_sample(toy_net, noise, cond, unc, n, g, lo, hi, shift, exact_henu=True)
        assert rel_l2(o, r) < 1e-6, rel_l2(o, r)
        out[f"heun_exact_{n}"] = r.numpy()
    for n, order, shift, g in [(25, 2, 3.0, 4.0), (8, 3, 1.0, 2.0), (6, 4, 2.0, 1.5)]:
        a = RefAdam(order=order, timeshift=shift, scheduler=sch, guidance_fn=ref_guidance, num_steps=n,
                    guidance=g, guidance_interval_min=0.0, guidance_interval_max=1.0)
        ts, deltas, coeffs = O.adam_coeffs(n, order, shift)
        assert torch.equal(ts, a.timesteps)
        for i in range(n):
            rc = [float(c) for c in a.solver_coeffs[i]]
            assert np.allclose(rc, coeffs[i], rtol=2e-4, atol=2e-5), (i, rc, coeffs[i])
        r = a(toy_net, noise, cond, unc)
        o = O.adam_sample(toy_net, noise, cond, unc, n, g, order, 0.0, 1.0, shift)
        assert rel_l2(o, r) < 1e-4, rel_l2(o, r)
        out[f"adam_{n}_{order}"] = r.numpy()
        out[f"adam_{n}_{order}_coeffs"] = np.array(
            [[float(c) for c in a.solver_coeffs[i]] + [0.0] * (4 - len(a.solver_coeffs[i])) for i in range(n)])
        out[f"adam_{n}_{order}_ts"] = a.timesteps.numpy()
    out["noise"] = noise.numpy()
    np.savez_compressed(os.path.join(OUT, "samplers_toy.npz"), **out)
    print("[samplers] euler/heun/adam oracle == reference")


def golden_dct():
    tr = RefTrainer(scheduler=RefSched(), encoder=torch.nn.Identity(), freq_loss_weight=1, freq_quality=85)
    dct = RefTrainer._dct._torchdynamo_orig_callable if hasattr(RefTrainer._dct, "_torchdynamo_orig_callable") \
        else RefTrainer._dct
    assert torch.equal(tr.dct_mat, O.dct_matrix(8))
    assert torch.allclose(tr.freq_w[0, :, 0, 0], O.freq_weight(85), rtol=0, atol=0)
    res = {}
    for name, shape, seed in [("256", (2, 3, 256, 256), 7), ("ragged", (3, 3, 36, 44), 8), ("one", (1, 3, 8, 8), 9)]:
        g = torch.Generator().manual_seed(seed)
        out = torch.randn(shape, generator=g)
        v = torch.randn(shape, generator=g)
        with torch.enable_grad():
            o = out.clone().requires_grad_(True)
            fm = ((o - v) ** 2).mean()
            fr = (tr.freq_w * (dct(tr, tr._rgb2ycbcr(o)) - dct(tr, tr._rgb2ycbcr(v))) ** 2).mean()
            loss = fm + tr.freq_loss_weight * fr
            loss.backward()
        with torch.enable_grad():
            o2 = out.clone().requires_grad_(True)
            d = O.dct_fm_loss(o2, v)
            d["loss"].backward()
        assert abs(float(d["loss"]) - float(loss)) < 1e-6 * abs(float(loss))
        assert rel_l2(o2.grad, o.grad) < 1e-6
        print(f"[dct {name}] fm={float(fm):.6f} freq={float(fr):.6f} loss={float(loss):.6f} "
              f"|grad|={float(o.grad.norm()):.6e}")
        res[f"{name}_seed"] = np.int64(seed)
        res[f"{name}_shape"] = np.array(shape)
        res[f"{name}_fm"] = np.float64(fm)
        res[f"{name}_freq"] = np.float64(fr)
        res[f"{name}_loss"] = np.float64(loss)
        res[f"{name}_grad_norm"] = np.float64(o.grad.norm())
        if name != "256":
            res[f"{name}_grad"] = o.grad.numpy()
        else:
            res[f"{name}_grad_sub"] = o.grad[:, :, ::8, ::8].numpy()
    res["freq_w"] = tr.freq_w[0, :, 0, 0].numpy()
    res["dct_mat"] = tr.dct_mat.numpy()
    np.savez_compressed(os.path.join(OUT, "dct_loss.npz"), **res)


def golden_cfg1():
    """BASELINE.json configs[0]: DeCo-L/16 256px, batch 4, 10 Euler steps, CFG 3.2 on (0.1,1], fp32 CPU."""
    cfg = O.CFG_L
    m, P = build_ref(cfg)
    B = 4
    noise = seeded_noise(B, (3, 256, 256), 0)
    cond = torch.tensor([0, 250, 500, 750])
    unc = torch.full((B,), 1000)
    sch = RefSched()
    s = RefEuler(scheduler=sch, w_scheduler=sch, guidance_fn=ref_guidance, num_steps=10, guidance=3.2,
                 guidance_interval_min=0.1, guidance_interval_max=1.0, step_fn=ode_step_fn)
    t0 = time.time()
    x, xs, vs = s(m, noise, cond, unc, return_x_trajs=True, return_v_trajs=True)
    el = time.time() - t0
    print(f"[cfg1] reference L/16 10-step CFG sampling on {torch.get_num_threads()} threads: {el:.1f}s "
          f"-> {B / el:.4f} img/s")
    o = O.euler_sample(lambda a, b, c: O.denoiser_forward(P, cfg, a, b, c), noise, cond, unc, 10, 3.2, 0.1, 1.0)
    e = rel_l2(o, x)
    print(f"[cfg1] oracle vs reference 10-step trajectory rel-L2 = {e:.3e}")
    assert e < 1e-5
    # first forward of the trajectory: the CFG-batched network output at t=0
    np.savez_compressed(os.path.join(OUT, "cfg1_L16_euler10.npz"),
                        final_sub=x[:, :, ::4, ::4].numpy(), final_mean=np.float64(x.mean()),
                        final_std=np.float64(x.std()),
                        v0_sub=vs[0][:, :, ::4, ::4].numpy(),
                        x_step_norms=np.array([float(v.norm()) for v in xs]),
                        cond=cond.numpy(), ref_seconds=np.float64(el), ref_threads=np.int64(torch.get_num_threads()))


def build_ref_t2i(cfg):
    """The original text-to-image denoiser, re-assembled from the importable reference classes: encoder / text path
    from dit_t2i_pixnerd.py (Attention, FlattenDiTBlock, NerfEmbedder, TextRefineBlock, forward :276-297), decoder
    SimpleMLPAdaLN from dit_c2i_DeCo.py; member names as in the 3.10 bytecode of the original dit_t2i_DeCo.py."""
    import torch.nn as nn
    from src.models.transformer import dit_t2i_pixnerd as T
    from src.models.transformer.dit_c2i_DeCo import SimpleMLPAdaLN

    class RefT2I(nn.Module):
        def __init__(self):
            super().__init__()
            H = cfg.hidden_size
            self.cfg = cfg
            self.s_embedder = T.Embed(cfg.in_channels * cfg.patch_size ** 2, H, bias=True)
            self.x_embedder = T.NerfEmbedder(cfg.in_channels, cfg.decoder_hidden_size, max_freqs=8)
            self.t_embedder = T.TimestepEmbedder(H)
            self.y_embedder = T.Embed(cfg.txt_embed_dim, H, bias=True, norm_layer=T.Norm)
            self.y_pos_embedding = nn.Parameter(torch.randn(1, cfg.txt_max_length, H))
            self.blocks = nn.ModuleList([T.FlattenDiTBlock(H, cfg.num_groups) for _ in range(cfg.num_encoder_blocks)])
            self.dec_net = SimpleMLPAdaLN(in_channels=cfg.decoder_hidden_size, model_channels=cfg.decoder_hidden_size,
                                          out_channels=cfg.in_channels, z_channels=H,
                                          num_res_blocks=cfg.num_decoder_blocks, patch_size=cfg.patch_size)
            self.text_refine_blocks = nn.ModuleList([T.TextRefineBlock(H, cfg.num_groups)
                                                     for _ in range(cfg.num_text_blocks)])

        def forward(self, x, t, y):
            c = self.cfg
            B, _, Hh, Ww = x.shape
            p = c.patch_size
            x = torch.nn.functional.unfold(x, kernel_size=p, stride=p).transpose(1, 2)
            xpos = T.precompute_freqs_cis_2d(c.hidden_size // c.num_groups, Hh // p, Ww // p)
            t = self.t_embedder(t.view(-1)).view(B, -1, c.hidden_size)
            y = self.y_embedder(y).view(B, -1, c.hidden_size) + self.y_pos_embedding.to(y.dtype)
            condition = torch.nn.functional.silu(t)
            for blk in self.text_refine_blocks:
                y = blk(y, condition)
            s = self.s_embedder(x)
            for blk in self.blocks:
                s = blk(s, y, condition, xpos)
            s = torch.nn.functional.silu(t + s)
            bsz, length, _ = s.shape
            x = x.reshape(bsz * length, c.in_channels, p ** 2).transpose(1, 2)
            s = s.view(bsz * length, c.hidden_size)
            x = self.x_embedder(x)
            x = self.dec_net(x, s)
            x = x.transpose(1, 2).reshape(bsz, length, -1)
            return torch.nn.functional.fold(x.transpose(1, 2).contiguous(), (Hh, Ww), kernel_size=p, stride=p)

    m = RefT2I()
    P = O.t2i_seeded_params(cfg)
    sd = m.state_dict()
    assert set(sd.keys()) == set(P.keys()), set(sd.keys()) ^ set(P.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(P[k].shape), k
    m.load_state_dict(P)
    return m.eval(), P


def golden_t2i(name, cfg, B, res, seed):
    import warnings
    m, P = build_ref_t2i(cfg)
    x = seeded_noise(B, (cfg.in_channels, res, res), seed)
    t = torch.linspace(0.1, 0.9, B)
    y = torch.randn((B, cfg.txt_max_length, cfg.txt_embed_dim), generator=torch.Generator().manual_seed(seed + 1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")     # NerfEmbedder casts its complex table to real (discards the imaginary part)
        ref = m(x, t, y)
    ora = O.t2i_forward(P, cfg, x, t, y)
    e = rel_l2(ora, ref)
    print(f"[{name}] t2i oracle vs composed reference forward rel-L2 = {e:.3e}")
    assert e < 2e-6, e
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), x=x.numpy(), t=t.numpy(), y=y.numpy(), out=ref.numpy(),
                        cfg=np.array([cfg.in_channels, cfg.num_groups, cfg.hidden_size, cfg.decoder_hidden_size,
                                      cfg.num_encoder_blocks, cfg.num_decoder_blocks, cfg.num_text_blocks,
                                      cfg.patch_size, cfg.txt_embed_dim, cfg.txt_max_length]))


def golden_baseline(name, cfg, B, res, seed):
    """Patch-linear baseline denoiser (dit_c2i_baseline.FlattenDiT): oracle pinned, fixture written."""
    m = RefFlattenDiT(in_channels=cfg.in_channels, num_groups=cfg.num_groups, hidden_size=cfg.hidden_size,
                      num_blocks=cfg.num_blocks, patch_size=cfg.patch_size, num_classes=cfg.num_classes)
    P = O.baseline_seeded_params(cfg)
    sd = m.state_dict()
    assert set(sd.keys()) == set(P.keys()), set(sd.keys()) ^ set(P.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(P[k].shape), k
    m.load_state_dict(P)
    m.eval()
    x = seeded_noise(B, (cfg.in_channels, res, res), seed)
    t = torch.linspace(0.05, 0.95, B)
    y = torch.tensor([(5 * i + 2) % (cfg.num_classes + 1) for i in range(B)])
    y[-1] = cfg.num_classes
    ref = m(x, t, y)
    e = rel_l2(O.baseline_forward(P, cfg, x, t, y), ref)
    print(f"[{name}] oracle vs reference FlattenDiT forward rel-L2 = {e:.3e}")
    assert e < 2e-6, e
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ref_bf = m(x, t, y).float()
    print(f"[{name}] reference bf16-autocast vs fp32 rel-L2 = {rel_l2(ref_bf, ref):.3e} (noise floor)")
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), x=x.numpy(), t=t.numpy(), y=y.numpy(), out=ref.numpy(),
                        cfg=np.array([cfg.in_channels, cfg.num_groups, cfg.hidden_size, cfg.num_blocks, cfg.patch_size,
                                      cfg.num_classes]), bf16_floor=np.float64(rel_l2(ref_bf, ref)))


def golden_pixnerd(name, cfg, B, res, seed):
    """PixNerd baseline (dit_c2i_pixnerd.PixNerDiT: hyper-network NerfBlock decoder): oracle pinned ahead of its CUDA path."""
    m = RefPixNerd(in_channels=cfg.in_channels, num_groups=cfg.num_groups, hidden_size=cfg.hidden_size,
                   hidden_size_x=cfg.hidden_size_x, nerf_mlpratio=cfg.nerf_mlpratio, num_blocks=cfg.num_blocks,
                   num_cond_blocks=cfg.num_cond_blocks, patch_size=cfg.patch_size, num_classes=cfg.num_classes)
    P = O.pixnerd_seeded_params(cfg)
    sd = m.state_dict()
    assert set(sd.keys()) == set(P.keys()), set(sd.keys()) ^ set(P.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(P[k].shape), k
    m.load_state_dict(P)
    m.eval()
    x = seeded_noise(B, (cfg.in_channels, res, res), seed)
    t = torch.linspace(0.1, 0.9, B)
    y = torch.tensor([(3 * i + 1) % (cfg.num_classes + 1) for i in range(B)])
    y[-1] = cfg.num_classes
    ref = m(x, t, y)
    e = rel_l2(O.pixnerd_forward(P, cfg, x, t, y), ref)
    print(f"[{name}] oracle vs reference PixNerd forward rel-L2 = {e:.3e}")
    assert e < 2e-6, e
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), x=x.numpy(), t=t.numpy(), y=y.numpy(), out=ref.numpy(),
                        cfg=np.array([cfg.in_channels, cfg.num_groups, cfg.hidden_size, cfg.hidden_size_x, cfg.nerf_mlpratio,
                                      cfg.num_blocks, cfg.num_cond_blocks, cfg.patch_size, cfg.num_classes]))


def golden_samplers_ext():
    """EulerSamplerJiT and the SDE step functions of EulerSampler, pinned with analytic 'networks'.  The Gaussian increments
    of sde_step_fn / sde_preserve_step_fn are the reference's own torch.randn_like calls under torch.manual_seed; they are
    recorded so a consumer can replay them."""
    def toy_net(x, t, y):
        return torch.tanh(x * (0.5 + t.view(-1, 1, 1, 1))) - 0.1 * y.view(-1, 1, 1, 1).float()

    def toy_xnet(x, t, y):      # an x-prediction: bounded image estimate
        return torch.tanh(0.7 * x + 0.3 * t.view(-1, 1, 1, 1)) - 0.05 * y.view(-1, 1, 1, 1).float()

    noise = seeded_noise(3, (3, 8, 8), 200)
    cond, unc = torch.tensor([1, 2, 3]), torch.tensor([10, 10, 10])
    sch = RefSched()
    out = {"noise": noise.numpy()}
    for n, g, lo, hi, shift in [(12, 2.5, 0.1, 1.0, 1.0), (30, 1.5, 0.0, 0.8, 2.0)]:
        j = RefEulerJiT(scheduler=sch, w_scheduler=sch, guidance_fn=ref_guidance, num_steps=n, guidance=g,
                        guidance_interval_min=lo, guidance_interval_max=hi, timeshift=shift, step_fn=ode_step_fn)
        r = j(toy_xnet, noise, cond, unc)
        o = O.euler_sample_ex(toy_xnet, noise, cond, unc, n, g, lo, hi, shift, x_prediction=True)
        assert rel_l2(o, r) < 1e-6, rel_l2(o, r)
        out[f"jit_{n}"] = r.numpy()
    fns = {"sde_mean": sde_mean_step_fn, "sde": sde_step_fn, "sde_preserve": sde_preserve_step_fn}
    for kind, fn in fns.items():
        for n, g, shift, last in [(10, 2.0, 1.0, "ode"), (6, 1.0, 2.0, kind)]:
            e = RefEuler(scheduler=sch, w_scheduler=sch, guidance_fn=ref_guidance, num_steps=n, guidance=g,
                         guidance_interval_min=0.1, guidance_interval_max=1.0, timeshift=shift, step_fn=fn,
                         last_step_fn=(ode_step_fn if last == "ode" else fn))
            incs = []
            real = torch.randn_like

            def rec(x):
                z = real(x)
                incs.append(z)
                return z
            torch.manual_seed(77)
            torch.randn_like = rec
            try:
                r = e(toy_net, noise, cond, unc)
            finally:
                torch.randn_like = real
            torch.manual_seed(77)
            o = O.euler_sample_ex(toy_net, noise, cond, unc, n, g, 0.1, 1.0, shift, step=kind, last=last)
            assert rel_l2(o, r) < 1e-6, (kind, rel_l2(o, r))
            out[f"{kind}_{n}"] = r.numpy()
            if incs:
                out[f"{kind}_{n}_increments"] = torch.stack(incs).numpy()
    # HeunSampler with the SDE step functions: scores at (x, t_cur) and (x_hat, t_next) averaged (sampling.py:283-291)
    for kind, fn in fns.items():
        for n, g, shift, last, exact in [(8, 2.0, 1.0, "ode", False), (5, 1.5, 2.0, kind, True)]:
            h = RefHeun(scheduler=sch, w_scheduler=sch, exact_henu=exact, guidance_fn=ref_guidance, num_steps=n, guidance=g,
                        guidance_interval_min=0.1, guidance_interval_max=1.0, timeshift=shift, step_fn=fn,
                        last_step_fn=(ode_step_fn if last == "ode" else fn))
            torch.manual_seed(99)
            r = h(toy_net, noise, cond, unc)
            torch.manual_seed(99)
            o = O.heun_sample_ex(toy_net, noise, cond, unc, n, g, 0.1, 1.0, shift, exact_henu=exact, step=kind, last=last)
            assert rel_l2(o, r) < 1e-6, ("heun", kind, rel_l2(o, r))
            out[f"heun_{kind}_{n}"] = r.numpy()
    np.savez_compressed(os.path.join(OUT, "samplers_ext_toy.npz"), **out)
    print("[samplers-ext] EulerSamplerJiT / sde_mean / sde / sde_preserve (Euler and Heun) oracle == reference")


def golden_trainstep():
    """BaseTrainer.__call__ + REPATrainer._impl_trainstep of the reference (label dropout, the 90/10 timestep mixture,
    time shift, x_t / v_t, FM loss) on CPU under fixed seeds, with an analytic recording net: pins oracle.trainstep (same
    global-generator draw order) and stores t / labels / the loss as a fixture."""
    def make_net(rec):
        def net(x_t, t, y):
            rec.update(x_t=x_t.clone(), t=t.clone(), y=y.clone())
            return torch.tanh(x_t * (0.5 + t.view(-1, 1, 1, 1))) - 0.1 * y.view(-1, 1, 1, 1).float()
        return net
    res = {}
    for name, B, shape, p, shift, seed in [("a", 6, (3, 16, 24), 0.5, 1.0, 101), ("b", 4, (3, 8, 8), 0.2, 2.5, 202)]:
        x = torch.tanh(torch.randn((B,) + shape, generator=torch.Generator().manual_seed(seed)))
        cond = torch.arange(B) % 10
        unc = torch.full((B,), 10)
        tr = RefTrainer(scheduler=RefSched(), encoder=torch.nn.Identity(), null_condition_p=p, timeshift=shift)
        rec, rec_o = {}, {}
        torch.manual_seed(seed)
        ref = tr(make_net(rec), None, None, x, cond, unc, metadata=dict(raw_image=x))
        torch.manual_seed(seed)
        got = O.trainstep(make_net(rec_o), x, cond, unc, null_condition_p=p, timeshift=shift, freq_loss_weight=0.0)
        for k in ("x_t", "t", "y"):
            assert torch.equal(rec[k], rec_o[k]), k
        assert torch.equal(ref["loss"], got["loss"]) and torch.equal(ref["fm_loss"], got["fm_loss"])
        print(f"[trainstep {name}] oracle == reference (t, y, x_t bit-equal); loss {float(ref['loss']):.6f}")
        res[f"{name}_cfg"] = np.array([B, *shape, seed])
        res[f"{name}_p_shift"] = np.array([p, shift])
        res[f"{name}_t"] = rec["t"].numpy()
        res[f"{name}_y"] = rec["y"].numpy()
        res[f"{name}_xt_sub"] = rec["x_t"][:, :, ::4, ::4].numpy()
        res[f"{name}_loss"] = np.float64(ref["loss"])
    np.savez_compressed(os.path.join(OUT, "trainstep.npz"), **res)


if __name__ == "__main__":
    which = sys.argv[1:] or ["dct", "samplers", "tiny", "t2i", "cfg1", "baseline", "samplers_ext", "pixnerd", "trainstep"]
    if "trainstep" in which:
        golden_trainstep()
    if "pixnerd" in which:
        golden_pixnerd("pixnerd_d64", O.PixNerdCfg(num_groups=4, hidden_size=256, hidden_size_x=64, nerf_mlpratio=2, num_blocks=4,
                                                   num_cond_blocks=2, num_classes=10), B=2, res=64, seed=51)
    if "baseline" in which:
        golden_baseline("baseline_d64", O.BaselineCfg(num_groups=4, hidden_size=256, num_blocks=3, num_classes=10),
                        B=3, res=64, seed=41)
    if "samplers_ext" in which:
        golden_samplers_ext()
    if "dct" in which:
        golden_dct()
    if "samplers" in which:
        golden_samplers()
    if "tiny" in which:
        # XL-like head_dim 72 (non power of two), L-like head_dim 64 with a ragged FFN width
        golden_forward("fwd_d72", O.DenoiserCfg(num_groups=8, hidden_size=576, num_blocks=5, num_cond_blocks=3,
                                                num_classes=10), B=4, res=64, seed=11)
        golden_forward("fwd_d64", O.DenoiserCfg(num_groups=4, hidden_size=256, num_blocks=4, num_cond_blocks=2,
                                                num_classes=10), B=2, res=96, seed=21)
    if "t2i" in which:
        # XXL-t2i-like: head_dim 64, joint [image || text] keys with a ragged text length, 2 text + 2 image blocks
        golden_t2i("t2i_d64", O.T2ICfg(num_groups=4, hidden_size=256, num_encoder_blocks=2, num_decoder_blocks=3,
                                       num_text_blocks=2, txt_embed_dim=96, txt_max_length=24), B=2, res=64, seed=31)
    if "cfg1" in which:
        golden_cfg1()
