"""CPU: the oracle against (a) the closed-form known answers of SURVEY.md section 4 and (b) the golden fixtures
generated from the unmodified reference by tests/golden/make_golden.py."""
import numpy as np
import torch

from helpers import cfg_from_array, load_golden, rel_l2, seeded_noise, toy_net
from oracle import deco_oracle as O


def test_known_answers_tables():
    C = O.dct_matrix(8)
    assert torch.allclose(C[0], torch.full((8,), 0.353553), atol=1e-6)
    assert torch.allclose(C[1, :4], torch.tensor([0.490393, 0.415735, 0.277785, 0.097545]), atol=1e-6)
    assert float((C @ C.t() - torch.eye(8)).abs().max()) < 1e-6
    w = O.freq_weight(85)
    assert w.shape == (3, 8, 8) and torch.allclose(w.mean((1, 2)), torch.ones(3), atol=1e-6)
    assert torch.allclose(w[0, 0], torch.tensor([2.013036, 3.355059, 3.355059, 2.013036, 1.437883, 0.838765,
                                                 0.671012, 0.559177]), atol=2e-6)
    assert abs(float(w[0].min()) - 0.279588) < 1e-6
    assert torch.allclose(w[1, 0, :4], torch.tensor([3.874020, 3.874020, 2.767157, 1.383579]), atol=2e-6)
    assert torch.allclose(w[1, 4:], torch.full((4, 8), 0.645670), atol=1e-6)
    e = O.timestep_embedding(torch.tensor([0.5]))[0]
    assert torch.allclose(e[:3], torch.tensor([0.8775826, 0.8818213, 0.8859162]), atol=1e-6)
    assert abs(float(e[127]) - 0.9987045) < 1e-6 and abs(float(e[128 + 127]) - 0.0508856) < 1e-6
    a = O.rope_table_2d(72, 16, 16)
    assert a.shape == (256, 36)
    assert abs(float(torch.cos(a[1, 0])) - 0.4830455) < 1e-6 and abs(float(torch.sin(a[1, 0])) - 0.8755953) < 1e-6
    assert float(a[1, 1]) == 0.0 and abs(float(torch.cos(a[1, 2])) - 0.8024242) < 1e-6
    assert float(a[16, 0]) == 0.0 and abs(float(torch.cos(a[16, 1])) - 0.4830455) < 1e-6
    assert abs(float(torch.cos(O.rope_table_2d(72, 32, 32)[1, 0])) - 0.8697361) < 1e-6
    tab = O.nerf_pos_table(16, 8)
    assert tab.shape == (256, 64) and abs(float(tab.sum()) - 339.41028) < 2e-3
    assert torch.allclose(tab[17, :4], torch.tensor([1.0, 0.97149, 0.8875858, 0.7530714]), atol=1e-5)


def test_known_answers_schedules_and_sizes():
    ts = O.make_timesteps(100)
    assert ts.shape == (101,) and float(ts[10]) == 0.09999999403953552 and not bool(ts[10] > 0.1)
    assert sum(bool(t > 0.1) and bool(t <= 1.0) for t in ts[:-1]) == 89
    ts, deltas, coeffs = O.adam_coeffs(25, 2, 3.0)
    assert np.allclose(ts[:4].numpy(), [0, 0.0136986, 0.0281690, 0.0434783], atol=1e-6)
    assert coeffs[0] == (1.0,)
    assert np.allclose(coeffs[1], (-0.52816892, 1.52816892), atol=2e-5)
    assert np.allclose(coeffs[-1], (-0.57999986, 1.57999980), atol=2e-5)
    assert sum(int(np.prod(s)) for s in O.param_shapes(O.CFG_XL).values()) == 682_282_851
    assert sum(int(np.prod(s)) for s in O.param_shapes(O.CFG_L).values()) == 426_938_019
    assert O.CFG_XL.ffn_hidden == 3072 and O.CFG_L.ffn_hidden == 2730


def test_forward_matches_reference_fixtures():
    for name in ("fwd_d72", "fwd_d64"):
        g = load_golden(name + ".npz")
        cfg = cfg_from_array(g["cfg"])
        P = O.seeded_params(cfg)
        out = O.denoiser_forward(P, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), torch.from_numpy(g["y"]))
        assert rel_l2(out, torch.from_numpy(g["out"])) < 2e-6
        assert float(torch.from_numpy(g["out"]).abs().mean()) > 1e-2   # non-vacuous: default init would give zeros


def test_t2i_forward_matches_reference_fixture():
    """The t2i oracle against the output of the reference classes composed in tests/golden/make_golden.py::build_ref_t2i."""
    from helpers import t2i_cfg_from_array
    g = load_golden("t2i_d64.npz")
    cfg = t2i_cfg_from_array(g["cfg"])
    P = O.t2i_seeded_params(cfg)
    out = O.t2i_forward(P, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), torch.from_numpy(g["y"]))
    assert rel_l2(out, torch.from_numpy(g["out"])) < 2e-6
    assert float(torch.from_numpy(g["out"]).abs().mean()) > 1e-2
    # known answers: XXL t2i sizes (README "1.1B"; configs_t2i/sft_res512.yaml:45-56)
    n = sum(int(np.prod(s)) for s in O.t2i_param_shapes(O.CFG_XXL_T2I).values())
    assert 1.10e9 < n < 1.16e9, n
    assert O.CFG_XXL_T2I.head_dim == 64 and O.CFG_XXL_T2I.ffn_hidden == 6144
    # the NerfEmbedder table is the real part of the complex ex2d table: first pixel has angle 0 -> all ones
    assert torch.equal(O.t2i_nerf_pos_table(16)[0], torch.ones(64))


def test_samplers_match_reference_fixtures():
    g = load_golden("samplers_toy.npz")
    noise = torch.from_numpy(g["noise"])
    cond, unc = torch.tensor([1, 2, 3]), torch.tensor([10, 10, 10])
    for n, gd, lo, hi, shift in [(10, 3.2, 0.1, 1.0, 1.0), (7, 2.0, 0.0, 0.6, 3.0)]:
        assert np.array_equal(O.make_timesteps(n, shift).numpy(), g[f"euler_{n}_ts"])
        assert rel_l2(O.euler_sample(toy_net, noise, cond, unc, n, gd, lo, hi, shift), torch.from_numpy(g[f"euler_{n}"])) < 1e-6
        assert rel_l2(O.heun_sample(toy_net, noise, cond, unc, n, gd, lo, hi, shift), torch.from_numpy(g[f"heun_{n}"])) < 1e-6
        assert rel_l2(O.heun_sample(toy_net, noise, cond, unc, n, gd, lo, hi, shift, exact_henu=True),
                      torch.from_numpy(g[f"heun_exact_{n}"])) < 1e-6
    for n, order, shift, gd in [(25, 2, 3.0, 4.0), (8, 3, 1.0, 2.0), (6, 4, 2.0, 1.5)]:
        assert rel_l2(O.adam_sample(toy_net, noise, cond, unc, n, gd, order, 0.0, 1.0, shift),
                      torch.from_numpy(g[f"adam_{n}_{order}"])) < 1e-4


def test_dct_loss_matches_reference_fixtures():
    g = load_golden("dct_loss.npz")
    assert np.array_equal(O.freq_weight(85).numpy(), g["freq_w"]) and np.array_equal(O.dct_matrix().numpy(), g["dct_mat"])
    # SURVEY.md section 4 seeded check
    assert abs(float(g["256_fm"]) - 2.005540) < 1e-6 and abs(float(g["256_freq"]) - 0.847859) < 1e-6
    assert abs(float(g["256_grad_norm"]) - 6.70335e-3) < 1e-8
    for name in ("256", "ragged", "one"):
        shape, seed = tuple(int(v) for v in g[f"{name}_shape"]), int(g[f"{name}_seed"])
        gen = torch.Generator().manual_seed(seed)
        out = torch.randn(shape, generator=gen).requires_grad_(True)
        v = torch.randn(shape, generator=gen)
        d = O.dct_fm_loss(out, v)
        d["loss"].backward()
        assert abs(float(d["loss"]) - float(g[f"{name}_loss"])) < 1e-6 * float(g[f"{name}_loss"])
        assert abs(float(out.grad.norm()) - float(g[f"{name}_grad_norm"])) < 1e-6 * float(g[f"{name}_grad_norm"])


def test_cfg1_first_step_matches_reference_fixture():
    """configs[0] (L/16, batch 4, CFG): the first CFG-batched forward of the trajectory; the full 10-step run
    (33 s of CPU) is pinned by make_golden.py itself (4.5e-7) and re-checked on the GPU against the CUDA path."""
    g = load_golden("cfg1_L16_euler10.npz")
    P = O.seeded_params(O.CFG_L)
    noise = seeded_noise(4, (3, 256, 256), 0)
    cond, unc = torch.from_numpy(g["cond"]), torch.full((4,), 1000)
    with torch.no_grad():
        out = O.denoiser_forward(P, O.CFG_L, torch.cat([noise, noise]), torch.zeros(8), torch.cat([unc, cond]))
    v0 = O.cfg_combine(out, 1.0)   # t = 0 is outside the (0.1, 1] guidance window
    assert rel_l2(v0[:, :, ::4, ::4], torch.from_numpy(g["v0_sub"])) < 1e-5


def test_baseline_forward_matches_reference_fixture():
    """FlattenDiT (dit_c2i_baseline.py) restatement against the live reference's output (make_golden.py::golden_baseline)."""
    from helpers import baseline_cfg_from_array
    g = load_golden("baseline_d64.npz")
    cfg = baseline_cfg_from_array(g["cfg"])
    P = O.baseline_seeded_params(cfg)
    out = O.baseline_forward(P, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), torch.from_numpy(g["y"]))
    assert rel_l2(out, torch.from_numpy(g["out"])) < 2e-6
    assert float(torch.from_numpy(g["out"]).abs().mean()) > 1e-2
    # known answers: configs_c2i/Baseline_DiT_JiT.yaml (hidden 1024, 24 blocks, 16 heads of 64, FFN 2730)
    big = O.BaselineCfg()
    assert big.head_dim == 64 and big.ffn_hidden == 2730
    n = sum(int(np.prod(s)) for s in O.baseline_param_shapes(big).values())
    assert 4.5e8 < n < 4.7e8, n


def test_extended_euler_samplers_match_reference_fixtures():
    """EulerSamplerJiT and the SDE step functions (sampling.py:17-24, :109-188); the stochastic ones replay the reference's
    recorded Gaussian increments."""
    from helpers import toy_xnet
    g = load_golden("samplers_ext_toy.npz")
    noise = torch.from_numpy(g["noise"])
    cond, unc = torch.tensor([1, 2, 3]), torch.tensor([10, 10, 10])
    for n, gd, lo, hi, shift in [(12, 2.5, 0.1, 1.0, 1.0), (30, 1.5, 0.0, 0.8, 2.0)]:
        o = O.euler_sample_ex(toy_xnet, noise, cond, unc, n, gd, lo, hi, shift, x_prediction=True)
        assert rel_l2(o, torch.from_numpy(g[f"jit_{n}"])) < 1e-6
    for kind in ("sde_mean", "sde", "sde_preserve"):
        for n, gd, shift, last in [(10, 2.0, 1.0, "ode"), (6, 1.0, 2.0, kind)]:
            incs = list(torch.from_numpy(g[f"{kind}_{n}_increments"])) if f"{kind}_{n}_increments" in g else []
            o = O.euler_sample_ex(toy_net, noise, cond, unc, n, gd, 0.1, 1.0, shift, step=kind, last=last,
                                  randn=lambda x: incs.pop(0))
            assert rel_l2(o, torch.from_numpy(g[f"{kind}_{n}"])) < 1e-6, kind
            assert not incs


def test_heun_sde_oracle_matches_reference_fixtures():
    """HeunSampler with the SDE step functions (sampling.py:266-293: scores averaged like velocities): the oracle under the
    seed the fixture was made with (CPU generator: the same draws as the reference's torch.randn_like calls)."""
    g = load_golden("samplers_ext_toy.npz")
    noise = torch.from_numpy(g["noise"])
    cond, unc = torch.tensor([1, 2, 3]), torch.tensor([10, 10, 10])
    for kind in ("sde_mean", "sde", "sde_preserve"):
        for n, gd, shift, last, exact in [(8, 2.0, 1.0, "ode", False), (5, 1.5, 2.0, kind, True)]:
            torch.manual_seed(99)
            o = O.heun_sample_ex(toy_net, noise, cond, unc, n, gd, 0.1, 1.0, shift, exact_henu=exact, step=kind, last=last)
            assert rel_l2(o, torch.from_numpy(g[f"heun_{kind}_{n}"])) < 1e-6, (kind, n)


def test_pixnerd_forward_matches_reference_fixture():
    """dit_c2i_pixnerd.PixNerDiT (hyper-network NerfBlock decoder, configs_c2i/Baseline_PixNerd.yaml) restated ahead of its
    CUDA path: the oracle against the live reference's output (make_golden.py::golden_pixnerd)."""
    g = load_golden("pixnerd_d64.npz")
    a = [int(v) for v in g["cfg"]]
    cfg = O.PixNerdCfg(in_channels=a[0], num_groups=a[1], hidden_size=a[2], hidden_size_x=a[3], nerf_mlpratio=a[4],
                       num_blocks=a[5], num_cond_blocks=a[6], patch_size=a[7], num_classes=a[8])
    P = O.pixnerd_seeded_params(cfg)
    out = O.pixnerd_forward(P, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), torch.from_numpy(g["y"]))
    assert rel_l2(out, torch.from_numpy(g["out"])) < 2e-6
    assert float(torch.from_numpy(g["out"]).abs().mean()) > 1e-2
    # known answers: Baseline_PixNerd.yaml sizes -- each NerfBlock generates 2 * 64 * 128 weights per patch from H = 1024
    big = O.PixNerdCfg()
    sh = O.pixnerd_param_shapes(big)
    assert sh["blocks.22.param_generator1.0.weight"] == (16384, 1024) and "blocks.21.attn.qkv.weight" in sh
    assert "blocks.22.attn.qkv.weight" not in sh and sh["final_layer.linear.weight"] == (3, 64)


def test_trainstep_oracle_vs_reference_fixture():
    """oracle.trainstep (label dropout, 90/10 timestep mixture, time shift, x_t / v_t, FM loss) reproduces what the
    reference's BaseTrainer.__call__ + REPATrainer._impl_trainstep drew and computed under the same CPU seeds
    (tests/golden/make_golden.py::golden_trainstep)."""
    g = load_golden("trainstep.npz")
    for name in ("a", "b"):
        B, c, h, w, seed = (int(v) for v in g[f"{name}_cfg"])
        p, shift = (float(v) for v in g[f"{name}_p_shift"])
        x = torch.tanh(torch.randn((B, c, h, w), generator=torch.Generator().manual_seed(seed)))
        cond, unc = torch.arange(B) % 10, torch.full((B,), 10)
        rec = {}

        def net(x_t, t, y):
            rec.update(x_t=x_t, t=t, y=y)
            return torch.tanh(x_t * (0.5 + t.view(-1, 1, 1, 1))) - 0.1 * y.view(-1, 1, 1, 1).float()
        torch.manual_seed(seed)
        d = O.trainstep(net, x, cond, unc, null_condition_p=p, timeshift=shift, freq_loss_weight=0.0)
        assert np.array_equal(rec["t"].numpy(), g[f"{name}_t"]) and np.array_equal(rec["y"].numpy(), g[f"{name}_y"])
        assert np.array_equal(rec["x_t"][:, :, ::4, ::4].numpy(), g[f"{name}_xt_sub"])
        assert abs(float(d["loss"]) - float(g[f"{name}_loss"])) <= 1e-7 * abs(float(g[f"{name}_loss"]))
