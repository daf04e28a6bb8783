"""SURVEY 8f rank 4 on the GPU: the patch-linear baseline denoiser (FlattenDiT), EulerSamplerJiT and the SDE step functions,
against the golden fixtures made from the real reference and against the oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from helpers import (baseline_cfg_from_array, build_baseline_module, load_golden, psnr, rel_l2, seeded_noise, toy_net,
                     toy_xnet)
from oracle import deco_oracle as O

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def test_layernorm_modulate_and_unpatchify_kernels(cuda_dev):
    from deco_b200 import ops
    g = torch.Generator().manual_seed(3)
    for M, H, L in [(2 * 16, 256, 16), (3 * 9, 1024, 9), (2 * 4, 1536, 4)]:
        x = (torch.randn(M, H, generator=g) * 2 + 0.7).to(cuda_dev)
        mod = (torch.randn(M // L, 2 * H + 8, generator=g) * 0.5).to(cuda_dev).to(bf16)
        sh, sc = mod[:, :H], mod[:, H:2 * H]
        out = ops.layernorm_modulate(x, sh, sc, L)
        ref = torch.nn.functional.layer_norm(x, (H,), None, None, 1e-6).view(M // L, L, H) \
            * (1 + sc.float().unsqueeze(1)) + sh.float().unsqueeze(1)
        assert out.dtype == bf16 and rel_l2(out.float(), ref.reshape(M, H)) < 4e-3
        assert torch.equal(out, ref.reshape(M, H).to(bf16)) or rel_l2(out.float(), ref.reshape(M, H).to(bf16).float()) < 2e-3
    for B, C, Hh, Ww, p in [(2, 3, 64, 64, 16), (1, 3, 32, 96, 16), (2, 4, 16, 16, 8)]:
        tok = torch.randn(B * (Hh // p) * (Ww // p), C * p * p, generator=g).to(cuda_dev).to(bf16)
        img = ops.unpatchify(tok, B, C, Hh, Ww, p)
        ref = torch.nn.functional.fold(tok.view(B, -1, C * p * p).transpose(1, 2).float(), (Hh, Ww), kernel_size=p, stride=p)
        assert torch.equal(img.float(), ref)                        # pure data movement: bit-exact
        assert torch.equal(ops.patchify(img.float(), p), tok)       # fold o unfold round trip


def test_baseline_forward_vs_reference_golden(cuda_dev):
    """Tolerance (north_star): relative L2 <= 1e-2 per bf16 denoiser forward, against the fp32 reference output."""
    g = load_golden("baseline_d64.npz")
    cfg = baseline_cfg_from_array(g["cfg"])
    m, P = build_baseline_module(cfg, cuda_dev)
    x, t, y = (torch.from_numpy(g[k]).to(cuda_dev) for k in ("x", "t", "y"))
    out = m(x, t, y)
    assert out.dtype == bf16 and out.shape == x.shape
    ref = torch.from_numpy(g["out"])
    e = rel_l2(out.float(), ref)
    print(f"baseline_d64: rel-L2 vs reference fp32 = {e:.3e} (reference's own bf16 floor {float(g['bf16_floor']):.3e})")
    assert e <= 1e-2
    Pd = {k: v.to(cuda_dev) for k, v in P.items()}
    assert rel_l2(O.baseline_forward(Pd, cfg, x, t, y), ref) < 1e-4
    out2, s_out = m.forward_sx(x, t, y)
    assert torch.equal(out2, out) and s_out.shape == (x.shape[0], cfg.hidden_size, 4, 4)
    with pytest.raises(NotImplementedError):     # parameter gradients only (tests/test_gpu_backward.py covers .train())
        m.train()(x.clone().requires_grad_(True), t, y)
    m.eval()


def test_baseline_jit_config_shape_vs_oracle(cuda_dev):
    """configs_c2i/Baseline_DiT_JiT.yaml architecture (hidden 1024, 16 heads of 64, ragged FFN 2730), 4 of its 24 blocks,
    256 px: forward parity and a 6-step EulerSamplerJiT trajectory (graphed and eager, bit-identical) against the oracle."""
    from deco_b200 import EulerSamplerJiT, LinearScheduler, ode_step_fn, simple_guidance_fn
    from deco_b200 import sampling as S
    cfg = O.BaselineCfg(num_blocks=4, num_classes=10)
    m, P = build_baseline_module(cfg, cuda_dev)
    Pd = {k: v.to(cuda_dev) for k, v in P.items()}
    B = 2
    x = seeded_noise(B, (3, 256, 256), 9).to(cuda_dev)
    t = torch.tensor([0.15, 0.8], device=cuda_dev)
    y = torch.tensor([3, 10], device=cuda_dev)
    ref = O.baseline_forward(Pd, cfg, x, t, y)
    e = rel_l2(m(x, t, y).float(), ref)
    print(f"FlattenDiT (JiT config shape, 4 blocks) rel-L2 vs fp32 oracle = {e:.3e}")
    assert e <= 1e-2
    sch = LinearScheduler()
    kw = dict(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=6, guidance=2.0,
              guidance_interval_min=0.1, guidance_interval_max=1.0, timeshift=1.0, step_fn=ode_step_fn)
    cond, unc = torch.tensor([3, 7], device=cuda_dev), torch.full((2,), 10, device=cuda_dev)
    S_graph = S.GRAPH
    try:
        S.GRAPH = True
        sg = EulerSamplerJiT(**kw)
        xg, ug = sg.sample_uint8(m, x, cond, unc)
        assert any(v[1] is not None for v in sg._steppers.values()), "the JiT step was not captured into a CUDA graph"
        S.GRAPH = False
        xe, ue = EulerSamplerJiT(**kw).sample_uint8(m, x, cond, unc)
    finally:
        S.GRAPH = S_graph
    assert torch.equal(xg, xe) and torch.equal(ug, ue)
    xo = O.euler_sample_ex(lambda a, b, c: O.baseline_forward(Pd, cfg, a, b, c), x, cond, unc, 6, 2.0, 0.1, 1.0, 1.0,
                           x_prediction=True)
    q = psnr(xe, xo, 2.0)
    print(f"6-step EulerSamplerJiT trajectory vs fp32 oracle: rel-L2 {rel_l2(xe, xo):.3e}, PSNR {q:.1f} dB")
    assert q >= 35.0        # north_star trajectory tolerance (peak-to-peak 2.0 on x in [-1, 1])


def test_extended_samplers_vs_reference_golden(cuda_dev):
    """EulerSamplerJiT and the SDE step functions with analytic nets: fixtures from the reference; the stochastic step
    functions against the oracle consuming the same CUDA generator stream (torch.randn_like, the reference's own call)."""
    from deco_b200 import (EulerSampler, EulerSamplerJiT, LinearScheduler, ode_step_fn, sde_mean_step_fn, sde_preserve_step_fn,
                           sde_step_fn, simple_guidance_fn)
    g = load_golden("samplers_ext_toy.npz")
    noise = torch.from_numpy(g["noise"]).to(cuda_dev)
    cond, unc = torch.tensor([1, 2, 3], device=cuda_dev), torch.tensor([10, 10, 10], device=cuda_dev)
    sch = LinearScheduler()
    for n, gd, lo, hi, shift in [(12, 2.5, 0.1, 1.0, 1.0), (30, 1.5, 0.0, 0.8, 2.0)]:
        s = EulerSamplerJiT(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=n, guidance=gd,
                            guidance_interval_min=lo, guidance_interval_max=hi, timeshift=shift, step_fn=ode_step_fn)
        x = s(toy_xnet, noise, cond, unc)
        assert rel_l2(x, torch.from_numpy(g[f"jit_{n}"])) < 5e-6
        x2, u8 = s.sample_uint8(toy_xnet, noise, cond, unc)
        assert torch.equal(x2, x) and torch.equal(u8, O.fp2uint8(x))
    fns = {"sde_mean": sde_mean_step_fn, "sde": sde_step_fn, "sde_preserve": sde_preserve_step_fn}
    for kind, fn in fns.items():
        for n, gd, shift, last in [(10, 2.0, 1.0, "ode"), (6, 1.0, 2.0, kind)]:
            s = EulerSampler(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=n, guidance=gd,
                             guidance_interval_min=0.1, guidance_interval_max=1.0, timeshift=shift, step_fn=fn,
                             last_step_fn=(ode_step_fn if last == "ode" else fn))
            torch.manual_seed(123)
            x = s(toy_net, noise, cond, unc)
            torch.manual_seed(123)
            xo = O.euler_sample_ex(toy_net, noise, cond, unc, n, gd, 0.1, 1.0, shift, step=kind, last=last)
            assert rel_l2(x, xo) < 5e-6, (kind, n, rel_l2(x, xo))
            if kind == "sde_mean":      # deterministic: the reference fixture itself
                assert rel_l2(x, torch.from_numpy(g[f"{kind}_{n}"])) < 5e-6
            else:                       # the stochastic runs must actually be stochastic
                torch.manual_seed(124)
                assert rel_l2(s(toy_net, noise, cond, unc), xo) > 1e-3


def test_heun_sde_samplers_vs_oracle(cuda_dev):
    """HeunSampler with sde_mean / sde / sde_preserve step functions (flow_matching/sampling.py:266-293): velocities and
    scores of the predictor and corrector evaluations are averaged, exact and re-use variants, SDE or ODE last step;
    against the oracle (pinned to the reference on CPU) consuming the same CUDA generator stream; the deterministic
    sde_mean runs also against the reference fixture itself."""
    from deco_b200 import (HeunSampler, LinearScheduler, ode_step_fn, sde_mean_step_fn, sde_preserve_step_fn, sde_step_fn,
                           simple_guidance_fn)
    g = load_golden("samplers_ext_toy.npz")
    noise = torch.from_numpy(g["noise"]).to(cuda_dev)
    cond, unc = torch.tensor([1, 2, 3], device=cuda_dev), torch.tensor([10, 10, 10], device=cuda_dev)
    sch = LinearScheduler()
    fns = {"sde_mean": sde_mean_step_fn, "sde": sde_step_fn, "sde_preserve": sde_preserve_step_fn}
    for kind, fn in fns.items():
        for n, gd, shift, last, exact in [(8, 2.0, 1.0, "ode", False), (5, 1.5, 2.0, kind, True)]:
            s = HeunSampler(scheduler=sch, w_scheduler=sch, exact_henu=exact, guidance_fn=simple_guidance_fn, num_steps=n,
                            guidance=gd, guidance_interval_min=0.1, guidance_interval_max=1.0, timeshift=shift, step_fn=fn,
                            last_step_fn=(ode_step_fn if last == "ode" else fn))
            torch.manual_seed(321)
            x = s(toy_net, noise, cond, unc)
            torch.manual_seed(321)
            xo = O.heun_sample_ex(toy_net, noise, cond, unc, n, gd, 0.1, 1.0, shift, exact_henu=exact, step=kind, last=last)
            assert rel_l2(x, xo) < 5e-6, (kind, n, rel_l2(x, xo))
            if kind == "sde_mean":
                assert rel_l2(x, torch.from_numpy(g[f"heun_{kind}_{n}"])) < 5e-6
            else:
                torch.manual_seed(322)
                assert rel_l2(s(toy_net, noise, cond, unc), xo) > 1e-3
            x2, u8 = s.sample_uint8(toy_net, noise, cond, unc) if kind == "sde_mean" else (None, None)
            if x2 is not None:
                assert torch.equal(x2, x) and torch.equal(u8, O.fp2uint8(x))


def test_sde_step_statistics_full_size(cuda_dev):
    """Size-independent property at the bench shape (256 x 3 x 256 x 256 elements): one sde_step_fn update with a zero
    network output and g = 1 is x_out = x (1 - a_s / sden) + a_n z, so the residual has mean 0 and variance a_n^2."""
    from deco_b200 import ops
    B = 64
    x = torch.randn(B, 3, 256, 256, device=cuda_dev)
    net = torch.zeros(2 * B, 3, 256, 256, device=cuda_dev, dtype=bf16)
    z = torch.randn_like(x)
    kd, sden, a_s, a_n = 0.3, 0.7, 0.05, 0.2
    xo, _, _, _ = ops.cfg_step_ex(x, net, 1.0, 0.1, kd=kd, sden=sden, a_s=a_s, a_n=a_n, noise=z)
    r = xo - x * (1 - a_s / sden)
    assert abs(float(r.mean())) < 1e-3 and abs(float(r.var()) - a_n * a_n) < 1e-3
    assert rel_l2(r, a_n * z) < 1e-5


@pytest.mark.parametrize("hw", [(32, 96), (80, 48)])
def test_non_square_images_both_denoisers(cuda_dev, hw):
    """Edge case the reference supports (any H, W divisible by the patch size; the RoPE table is per (h, w),
    dit_c2i_DeCo.py:467-473): non-square token grids, for the DeCo denoiser and the patch-linear baseline, one image."""
    from helpers import build_module
    Hh, Ww = hw
    x = seeded_noise(1, (3, Hh, Ww), 77).to(cuda_dev)
    t = torch.tensor([0.35], device=cuda_dev)
    y = torch.tensor([4], device=cuda_dev)
    cfg = O.DenoiserCfg(num_groups=4, hidden_size=256, num_blocks=4, num_cond_blocks=2, num_classes=10)
    m, P = build_module(cfg, cuda_dev)
    ref = O.denoiser_forward({k: v.to(cuda_dev) for k, v in P.items()}, cfg, x, t, y)
    e = rel_l2(m(x, t, y).float(), ref)
    print(f"PixNerDiT {Hh}x{Ww}: rel-L2 vs fp32 oracle = {e:.3e}")
    assert e <= 1e-2
    bcfg = O.BaselineCfg(num_groups=4, hidden_size=256, num_blocks=2, num_classes=10)
    mb, Pb = build_baseline_module(bcfg, cuda_dev)
    refb = O.baseline_forward({k: v.to(cuda_dev) for k, v in Pb.items()}, bcfg, x, t, y)
    eb = rel_l2(mb(x, t, y).float(), refb)
    print(f"FlattenDiT {Hh}x{Ww}: rel-L2 vs fp32 oracle = {eb:.3e}")
    assert eb <= 1e-2


def test_pixnerd_forward_vs_reference_golden_and_oracle(cuda_dev):
    """dit_c2i_pixnerd.PixNerDiT (hyper-network NerfBlock decoder; configs_c2i/Baseline_PixNerd.yaml): the fixture produced by
    the live reference (rel-L2 <= 1e-2, north_star forward tolerance), then the YAML's architecture (hidden 1024, 22 DiT + 2
    NerfBlocks, hidden_size_x 64, ratio 2) at 256 px against the fp32 oracle; forward(s=...) isolates the decoder."""
    from helpers import build_pixnerd_module, pixnerd_cfg_from_array
    g = load_golden("pixnerd_d64.npz")
    cfg = pixnerd_cfg_from_array(g["cfg"])
    m, P = build_pixnerd_module(cfg, cuda_dev)
    x, t, y = (torch.from_numpy(g[k]).to(cuda_dev) for k in ("x", "t", "y"))
    out = m(x, t, y)
    assert out.dtype == torch.bfloat16 and out.shape == x.shape
    e = rel_l2(out.float(), torch.from_numpy(g["out"]))
    print(f"pixnerd_d64: rel-L2 vs reference fp32 = {e:.3e}")
    assert e <= 1e-2
    # the decoder alone: same condition s for the kernel and the oracle (ragged token count: 3 images of 48 px = 27 tokens)
    Pd = {k: v.to(cuda_dev) for k, v in P.items()}
    gen = torch.Generator().manual_seed(4)
    xs = torch.randn(3, 3, 48, 48, generator=gen).to(cuda_dev)
    s = (torch.randn(3, 9, cfg.hidden_size, generator=gen) * 0.7).to(cuda_dev).to(torch.bfloat16)
    tz, yz = torch.zeros(3, device=cuda_dev), torch.zeros(3, dtype=torch.long, device=cuda_dev)
    h = O.F.linear(torch.cat([O.F.unfold(xs, 16, stride=16).transpose(1, 2).reshape(27, 3, 256).transpose(1, 2),
                              O.nerf_pos_table(16).to(cuda_dev)[None].expand(27, -1, -1)], -1),
                   Pd["x_embedder.embedder.0.weight"], Pd["x_embedder.embedder.0.bias"])
    for i in range(cfg.num_cond_blocks, cfg.num_blocks):
        h = O.nerf_block(Pd, f"blocks.{i}.", h, s.float().reshape(27, -1), cfg.nerf_mlpratio)
    h = O.F.linear(O.rmsnorm(h, Pd["final_layer.norm.weight"]), Pd["final_layer.linear.weight"], Pd["final_layer.linear.bias"])
    ref = O.F.fold(h.transpose(1, 2).reshape(3, 9, -1).transpose(1, 2).contiguous(), (48, 48), kernel_size=16, stride=16)
    got = m(xs, tz, yz, s=s)
    e = rel_l2(got.float(), ref)
    print(f"pixnerd decoder alone: rel-L2 vs oracle = {e:.3e}")
    assert e <= 8e-3
    # the YAML's architecture
    cfg = O.PixNerdCfg()
    m, P = build_pixnerd_module(cfg, cuda_dev)
    Pd = {k: v.to(cuda_dev) for k, v in P.items()}
    xb = seeded_noise(2, (3, 256, 256), 17).to(cuda_dev)
    tb = torch.tensor([0.2, 0.8], device=cuda_dev)
    yb = torch.tensor([1000, 33], device=cuda_dev)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = O.pixnerd_forward(Pd, cfg, xb, tb, yb)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    e = rel_l2(m(xb, tb, yb).float(), ref)
    print(f"PixNerd-L/16 256px: rel-L2 vs fp32 oracle = {e:.3e}")
    assert e <= 1e-2
