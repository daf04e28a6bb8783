"""CPU: host-side logic of the drop-in modules -- parameter/checkpoint contract, YAML wiring, sampler schedules and
control flow, weight packing, data-parallel sharding (gloo, world_size 2).  No CUDA kernel runs here."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_golden, rel_l2, toy_net
from oracle import deco_oracle as O

XL_YAML = """
model:
  vae:
    class_path: src.models.autoencoder.pixel.PixelAE
    init_args: {scale: 1.0}
  denoiser:
    class_path: src.models.transformer.dit_c2i_DeCo.PixNerDiT
    init_args:
      in_channels: 3
      patch_size: 16
      num_groups: 4
      hidden_size: &hidden_dim 256
      hidden_size_x: 32
      num_blocks: 5
      num_cond_blocks: 2
      nerf_mlpratio: 2
      num_classes: &num_classes 10
  conditioner:
    class_path: src.models.conditioner.class_label.LabelConditioner
    init_args: {num_classes: *num_classes}
  diffusion_trainer:
    class_path: src.diffusion.flow_matching.training_repa_DeCo.REPATrainer
    init_args:
      lognorm_t: true
      encoder:
        class_path: src.models.encoder.DINOv2
        init_args: {weight_path: /nonexistent}
      align_layer: 8
      proj_denoiser_dim: *hidden_dim
      proj_hidden_dim: *hidden_dim
      proj_encoder_dim: 768
      null_condition_p: 0.2
      scheduler: &scheduler src.diffusion.flow_matching.scheduling.LinearScheduler
  diffusion_sampler:
    class_path: src.diffusion.flow_matching.sampling.EulerSampler
    init_args:
      num_steps: 100
      guidance: 3.2
      guidance_interval_min: 0.1
      guidance_interval_max: 1.0
      scheduler: *scheduler
      w_scheduler: src.diffusion.flow_matching.scheduling.LinearScheduler
      guidance_fn: src.diffusion.base.guidance.simple_guidance_fn
      step_fn: src.diffusion.flow_matching.sampling.ode_step_fn
"""


@pytest.mark.parametrize("cfg", [O.CFG_XL, O.CFG_L])
def test_state_dict_is_the_reference_checkpoint_contract(cfg):
    from deco_b200 import PixNerDiT
    with torch.device("meta"):
        m = PixNerDiT(in_channels=3, num_groups=cfg.num_groups, hidden_size=cfg.hidden_size, hidden_size_x=32,
                      num_blocks=cfg.num_blocks, num_cond_blocks=cfg.num_cond_blocks, patch_size=16, num_classes=1000)
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert sd == O.param_shapes(cfg)
    assert m.weight_path is None and m.load_ema is False


def test_t2i_state_dict_is_the_reference_checkpoint_contract():
    from deco_b200 import config
    from deco_b200.denoiser_t2i import PixNerDiT as T2I
    cfg = O.CFG_XXL_T2I
    assert config.resolve("src.models.transformer.dit_t2i_DeCo.PixNerDiT") is T2I
    with torch.device("meta"):
        m = T2I(in_channels=3, patch_size=16, num_groups=24, hidden_size=1536, txt_embed_dim=2048, txt_max_length=128,
                num_text_blocks=4, decoder_hidden_size=32, num_encoder_blocks=16, num_decoder_blocks=3)
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert sd == O.t2i_param_shapes(cfg)
    assert m.weight_path is None and m.load_ema is False and m.num_blocks == 19
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T2I(in_channels=3, patch_size=16, num_groups=4, hidden_size=256, txt_embed_dim=32, txt_max_length=8,
            num_text_blocks=1, decoder_hidden_size=32, num_encoder_blocks=1, num_decoder_blocks=1).eval()(
            torch.zeros(1, 3, 32, 32), torch.zeros(1), torch.zeros(1, 8, 32))


def test_default_init_follows_reference():
    from deco_b200 import PixNerDiT
    torch.manual_seed(0)
    m = PixNerDiT(in_channels=3, num_groups=4, hidden_size=256, hidden_size_x=32, num_blocks=4, num_cond_blocks=1,
                  patch_size=16, num_classes=10)
    assert float(m.dec_net.final_layer.linear.weight.abs().max()) == 0.0
    assert all(float(b.adaLN_modulation[1].weight.abs().max()) == 0.0 for b in m.dec_net.res_blocks)
    assert float(m.s_embedder.proj.bias.abs().max()) == 0.0
    assert abs(float(m.y_embedder.embedding_table.weight.std()) - 0.02) < 2e-3


def test_yaml_wiring(tmp_path):
    from deco_b200 import EulerSampler, LinearScheduler, PixNerDiT, REPATrainer, config, ode_step_fn, simple_guidance_fn
    p = tmp_path / "cfg.yaml"
    p.write_text(XL_YAML)
    parts = config.load_model_section(str(p))
    assert isinstance(parts["denoiser"], PixNerDiT) and isinstance(parts["diffusion_sampler"], EulerSampler)
    s, tr = parts["diffusion_sampler"], parts["diffusion_trainer"]
    assert isinstance(tr, REPATrainer) and tr.null_condition_p == 0.2 and isinstance(tr.scheduler, LinearScheduler)
    assert s.guidance_fn is simple_guidance_fn and s.step_fn is ode_step_fn and isinstance(s.scheduler, LinearScheduler)
    assert s.num_steps == 100 and s.guidance == 3.2
    assert float(s.timesteps[10]) == 0.09999999403953552
    cond, unc = parts["conditioner"]([1, 2, 3], device="cpu")
    assert cond.dtype == torch.int64 and unc.tolist() == [10, 10, 10]


def test_no_cpu_fallback():
    from deco_b200 import EulerSampler, LinearScheduler, PixNerDiT, REPATrainer, ops
    m = PixNerDiT(in_channels=3, num_groups=4, hidden_size=256, hidden_size_x=32, num_blocks=4, num_cond_blocks=1,
                  patch_size=16, num_classes=10).eval()
    x = torch.zeros(1, 3, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, torch.zeros(1), torch.zeros(1, dtype=torch.long))
    s = EulerSampler(scheduler=LinearScheduler(), num_steps=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        s(toy_net, x, torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        REPATrainer(scheduler=LinearScheduler()).loss(x, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.fp2uint8(x)


def test_training_path_dispatch_and_no_cpu_fallback():
    """Every trainable denoiser goes through the same autograd node with its own forward / backward pair, and in .train()
    mode a CPU input still fails loudly (no fallback to PyTorch autograd)."""
    from deco_b200 import PixNerDiT, autograd as A, ops
    from deco_b200.denoiser_baseline import FlattenDiT
    from deco_b200.denoiser_t2i import PixNerDiT as T2I
    c2i = PixNerDiT(in_channels=3, num_groups=4, hidden_size=256, hidden_size_x=32, num_blocks=4, num_cond_blocks=1,
                    patch_size=16, num_classes=10)
    base = FlattenDiT(in_channels=3, num_groups=4, hidden_size=256, num_blocks=2, patch_size=16, num_classes=10)
    t2i = T2I(in_channels=3, num_groups=4, hidden_size=256, decoder_hidden_size=32, num_encoder_blocks=2, num_decoder_blocks=1,
              num_text_blocks=1, patch_size=16, txt_embed_dim=64, txt_max_length=8)
    assert A._train_fns(c2i) == (A.train_forward, A.train_backward)
    assert A._train_fns(base) == (A.baseline_train_forward, A.baseline_train_backward)
    assert A._train_fns(t2i) == (A.t2i_train_forward, A.t2i_train_backward)
    x = torch.zeros(1, 3, 32, 32)
    for m, y in ((c2i, torch.zeros(1, dtype=torch.long)), (base, torch.zeros(1, dtype=torch.long)), (t2i, torch.zeros(1, 8, 64))):
        m.train()
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(x, torch.zeros(1), y)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.center_rows(torch.zeros(4, 8))


def _torch_cfg_step(x, net_out, g, dt, c0=1.0, prev=(), coeffs=(), x_out=None, want_pred=False, want_v=False,
                    want_u8=False):
    """Test-only torch statement of csrc/sampler.cu so the samplers' host control flow can be checked without a GPU."""
    u, c = net_out.float().chunk(2)
    pred = u + g * (c - u)
    v = c0 * pred
    for p, cf in zip(prev, coeffs):
        v = v + cf * p
    xo = x + dt * v
    return xo, (pred if want_pred else None), (v if want_v else None), (O.fp2uint8(xo) if want_u8 else None)


def test_sampler_control_flow_matches_reference(monkeypatch):
    from deco_b200 import AdamLMSampler, EulerSampler, HeunSampler, LinearScheduler, ode_step_fn, ops, sampling, simple_guidance_fn
    monkeypatch.setattr(ops, "cfg_step", _torch_cfg_step)
    monkeypatch.setattr(sampling, "_prep_inputs", lambda n, c, u: (n.float().contiguous(), torch.cat([u, c], 0)))
    g = load_golden("samplers_toy.npz")
    noise = torch.from_numpy(g["noise"])
    cond, unc = torch.tensor([1, 2, 3]), torch.tensor([10, 10, 10])
    sch = LinearScheduler()
    for n, gd, lo, hi, shift in [(10, 3.2, 0.1, 1.0, 1.0), (7, 2.0, 0.0, 0.6, 3.0)]:
        kw = dict(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=n, guidance=gd,
                  guidance_interval_min=lo, guidance_interval_max=hi, timeshift=shift, step_fn=ode_step_fn)
        assert rel_l2(EulerSampler(**kw)(toy_net, noise, cond, unc), torch.from_numpy(g[f"euler_{n}"])) < 1e-6
        assert rel_l2(HeunSampler(**kw)(toy_net, noise, cond, unc), torch.from_numpy(g[f"heun_{n}"])) < 1e-6
        assert rel_l2(HeunSampler(exact_henu=True, **kw)(toy_net, noise, cond, unc),
                      torch.from_numpy(g[f"heun_exact_{n}"])) < 1e-6
    for n, order, shift, gd in [(25, 2, 3.0, 4.0), (8, 3, 1.0, 2.0), (6, 4, 2.0, 1.5)]:
        a = AdamLMSampler(order=order, timeshift=shift, scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=n,
                          guidance=gd, guidance_interval_min=0.0, guidance_interval_max=1.0)
        assert np.array_equal(a.timesteps.numpy(), g[f"adam_{n}_{order}_ts"])
        ref = g[f"adam_{n}_{order}_coeffs"]
        for i in range(n):
            assert np.allclose(a.solver_coeffs[i], ref[i][: len(a.solver_coeffs[i])], rtol=2e-4, atol=5e-5)
        assert rel_l2(a(toy_net, noise, cond, unc), torch.from_numpy(g[f"adam_{n}_{order}"])) < 1e-4
    with pytest.raises(NotImplementedError):
        EulerSampler(scheduler=sch, step_fn=lambda *a, **k: None, num_steps=2)


def test_decoder_fragment_packing_roundtrip():
    """_frag must place W[n][k] where mma.m16n8k16 expects B[k][n] (see csrc/decoder.cu)."""
    from deco_b200.denoiser import _frag
    w = torch.arange(96 * 32, dtype=torch.float32).reshape(96, 32).to(torch.bfloat16)
    for permuted in (False, True):
        f = _frag(w, permuted).float()                  # [12, 2, 32, 4]
        for j in (0, 5, 11):
            for s in (0, 1):
                for lane in (0, 7, 18, 31):
                    g, t = lane // 4, lane % 4
                    for e in range(4):
                        half, lo = e // 2, e % 2
                        k = (8 * t + 4 * s + 2 * half + lo) if permuted else (16 * s + 8 * half + 2 * t + lo)
                        assert float(f[j, s, lane, e]) == float(w[8 * j + g, k])
    # every (n, k) appears exactly once
    assert sorted(_frag(w, True).float().reshape(-1).tolist()) == sorted(w.float().reshape(-1).tolist())


def test_rank_sharding_matches_distributed_sampler():
    from torch.utils.data import DistributedSampler
    from deco_b200.data import ClassLabelRandomNDataset, rank_indices, seeded_noise
    ds = ClassLabelRandomNDataset(latent_shape=(3, 8, 8), num_classes=10, max_num_instances=25)
    for world in (1, 2, 4, 8):
        for r in range(world):
            ref = list(DistributedSampler(range(len(ds)), num_replicas=world, rank=r, shuffle=False))
            assert rank_indices(len(ds), r, world) == ref
    lat, cond, meta = ds[7]
    assert cond == 7 // ds.num_seeds
    ref = torch.randn((3, 8, 8), generator=torch.Generator().manual_seed(meta["seed"]))
    assert torch.equal(lat, ref)
    assert torch.equal(seeded_noise([meta["seed"]], (3, 8, 8), pin=False)[0], ref)


def _gather_worker(rank, world, port, n_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from deco_b200 import distributed as D
    from deco_b200.data import rank_indices
    r, w, _ = D.init_from_env("gloo")
    idx = rank_indices(n_total, r, w)
    local = torch.tensor(idx, dtype=torch.uint8).view(-1, 1, 1, 1).expand(-1, 3, 2, 2).contiguous()
    out = D.all_gather_images(local, w, total=n_total)
    if r == 0:
        ret.put(out[:, 0, 0, 0].tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_restores_global_order_gloo_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, 10, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert got == list(range(10))


def _grad_worker(rank, world, port, ret):
    import torch.distributed as dist
    from deco_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    D.init_from_env("gloo")
    ps = [torch.nn.Parameter(torch.zeros(s)) for s in [(3,), (2, 5), (1,)]]
    for i, p in enumerate(ps):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    D.all_reduce_gradients(ps, world)
    if rank == 0:
        ret.put([float(p.grad.mean()) for p in ps] + [float(p.grad.min()) for p in ps])
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_all_reduce_averages_over_ranks_gloo_world2():
    """Data-parallel training step (bench.py --workload train256 at N > 1): .grad <- mean over ranks, in place."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert got == [1.5, 3.0, 4.5, 1.5, 3.0, 4.5]


def test_model_loader_prefix_contract(tmp_path):
    """ModelLoader (src/utils/model_loader.py:10-28): `ema_denoiser.` / `denoiser.` prefixed Lightning checkpoints."""
    from deco_b200 import PixNerDiT
    from deco_b200.io import ModelLoader
    kw = dict(in_channels=3, num_groups=4, hidden_size=256, hidden_size_x=32, num_blocks=3, num_cond_blocks=1,
              patch_size=16, num_classes=10)
    src = PixNerDiT(**kw)
    ema = {k: torch.full_like(v, 0.25) for k, v in src.state_dict().items()}
    raw = {k: torch.full_like(v, -0.5) for k, v in src.state_dict().items()}
    ckpt = {"state_dict": {**{"ema_denoiser." + k: v for k, v in ema.items()},
                           **{"denoiser." + k: v for k, v in raw.items()}, "diffusion_trainer.x": torch.zeros(1)}}
    path = tmp_path / "last.ckpt"
    torch.save(ckpt, path)
    m = ModelLoader().load(PixNerDiT(**kw, weight_path=str(path), load_ema=True))
    assert all(float(v.min()) == 0.25 == float(v.max()) for v in m.state_dict().values())
    m = ModelLoader().load(PixNerDiT(**kw, weight_path=str(path), load_ema=False))
    assert all(float(v.min()) == -0.5 == float(v.max()) for v in m.state_dict().values())
    # missing entries are skipped, not fatal (the reference logs and continues)
    del ckpt["state_dict"]["ema_denoiser.s_embedder.proj.bias"]
    m2 = PixNerDiT(**kw, load_ema=True)
    before = m2.s_embedder.proj.bias.clone()
    ModelLoader().load(m2, ckpt)
    assert torch.equal(m2.s_embedder.proj.bias, before) and float(m2.s_embedder.proj.weight.max()) == 0.25


def test_image_sink_png_and_npz(tmp_path):
    """SaveImagesHook contract (src/callbacks/save_images.py:31-64): NHWC uint8 `arr_0` in output.npz, one PNG per image."""
    import struct
    import zlib
    from deco_b200.io import ImageSink, encode_png
    g = torch.Generator().manual_seed(0)
    imgs = torch.randint(0, 256, (12, 3, 16, 24), generator=g, dtype=torch.uint8)
    mds = [dict(filename=f"{i % 3}_{i}", seed=i, condition=i % 3) for i in range(12)]
    sink = ImageSink(str(tmp_path / "val"), save_compressed=True)
    sink.process_batch(imgs[:8], mds[:8])
    sink.process_batch(imgs[8:], mds[8:])
    npz = sink.close()
    arr = np.load(npz)["arr_0"]
    assert arr.shape == (12, 16, 24, 3) and np.array_equal(arr, imgs.permute(0, 2, 3, 1).numpy())
    files = sorted(p.name for p in (tmp_path / "val").glob("*.png"))
    assert len(files) == 12 and "0_0.png" in files      # first batch (8 < 10) and second (8 < 10 at entry) are both written
    # the PNG decodes back to the same pixels
    data = (tmp_path / "val" / "1_4.png").read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and data == encode_png(imgs[4].permute(1, 2, 0).numpy())
    pos, idat = 8, b""
    while pos < len(data):
        n, tag = struct.unpack(">I", data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        if tag == b"IDAT":
            idat += data[pos + 8:pos + 8 + n]
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(16, 1 + 24 * 3)
    assert np.array_equal(raw[:, 1:].reshape(16, 24, 3), imgs[4].permute(1, 2, 0).numpy())


def test_backward_fragment_packing_places_transposed_weights():
    """_frag_t packs W^T for the decoder dgrad MMAs: lane (g, t) of (n-tile j, k-step s) must hold
    W^T[n][16s + 2t (+1), 16s + 8 + 2t (+1)] with n = 8j + g, or the permuted input channel 8 (g / 2) + 2 j + g % 2."""
    from deco_b200.autograd import _frag_t
    for K in (32, 96):
        wt = torch.arange(32 * K, dtype=torch.float32).reshape(32, K).to(torch.bfloat16)   # [n = inputs, k = outputs]
        for perm in (False, True):
            f = _frag_t(wt, perm_n=perm)
            assert f.shape == (4, K // 16, 32, 4)
            for j, s, lane in [(0, 0, 0), (1, 1, 5), (3, K // 16 - 1, 31), (2, 0, 18)]:
                g, t = lane // 4, lane % 4
                n = 8 * (g // 2) + 2 * j + g % 2 if perm else 8 * j + g
                ks = [16 * s + 2 * t, 16 * s + 2 * t + 1, 16 * s + 8 + 2 * t, 16 * s + 9 + 2 * t]
                assert f[j, s, lane].tolist() == [wt[n, k].item() for k in ks]
    # every element of W^T appears exactly once
    wt = torch.randperm(32 * 96).reshape(32, 96).to(torch.float32)
    assert sorted(_frag_t(wt, perm_n=True).reshape(-1).tolist()) == sorted(wt.reshape(-1).tolist())


def test_training_blobs_have_the_sizes_the_kernels_expect():
    from deco_b200 import PixNerDiT
    from deco_b200.autograd import pack_decoder_bwd, pack_decoder_train
    m = PixNerDiT(in_channels=3, num_groups=2, hidden_size=144, hidden_size_x=32, num_blocks=4, num_cond_blocks=1,
                  patch_size=16, num_classes=10)
    blob = pack_decoder_train(m, "cpu")
    assert blob.dtype == torch.float32 and blob.numel() == 1152 + 3 * 5344 + 132
    # input_proj bias sits at [1120, 1152); the final layer's fourth row is padding
    assert torch.equal(blob[1120:1152], m.dec_net.input_proj.bias.detach())
    assert float(blob[1152 + 3 * 5344 + 96:1152 + 3 * 5344 + 128].abs().sum()) == 0.0
    assert pack_decoder_bwd(m, "cpu").numel() == 4 * (512 + 3 * 2560 + 96)


def test_graph_schedule_tables_match_the_eager_loops():
    """The device tables the CUDA-graphed stepper walks (deco_b200/sampling.py::_graph_rows) hold exactly the per-step
    scalars of the eager loops: Euler (window test `min < t <= max`, sampling.py:93) and Adams order 2 (strict upper bound,
    fp32-accumulated t, adam_sampling.py:96-118)."""
    from deco_b200 import AdamLMSampler, EulerSampler, HeunSampler, LinearScheduler, ode_step_fn, simple_guidance_fn
    sch = LinearScheduler()
    e = EulerSampler(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=100, guidance=3.2,
                     guidance_interval_min=0.1, guidance_interval_max=1.0, step_fn=ode_step_fn)
    rows, use_pred = e._graph_rows()
    assert not use_pred and len(rows) == 100
    ts = e.timesteps
    for i, r in enumerate(rows):
        g = 3.2 if (bool(ts[i] > 0.1) and bool(ts[i] <= 1.0)) else 1.0
        assert r[0] == g and r[1] == float(ts[i + 1] - ts[i]) and r[2] == 1.0 and r[3] == 0.0 and r[6] == float(ts[i])
    assert rows[10][0] == 1.0 and rows[11][0] == 3.2        # ts[10] = 0.0999999940 is NOT > 0.1
    a = AdamLMSampler(order=2, timeshift=3.0, scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=25, guidance=4.0,
                      guidance_interval_min=0.0, guidance_interval_max=1.0)
    rows, use_pred = a._graph_rows()
    assert use_pred and len(rows) == 25
    t = torch.zeros((), dtype=torch.float32)
    for i, r in enumerate(rows):
        cs = a.solver_coeffs[i]
        assert r[0] == (4.0 if (bool(t > 0.0) and bool(t < 1.0)) else 1.0)
        assert r[1] == float(a.timedeltas[i]) and r[2] == float(cs[-1]) and r[3] == (float(cs[0]) if len(cs) > 1 else 0.0)
        assert r[6] == float(t)
        t = t + a.timedeltas[i]
    assert rows[0][0] == 1.0 and rows[0][3] == 0.0          # first step: t = 0 is outside the open window, order 1
    assert AdamLMSampler(order=3, scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=8)._graph_rows() is None
    assert HeunSampler(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=4,
                       step_fn=ode_step_fn)._graph_rows() is None


def _overlap_avg_worker(rank, world, port, q):
    import os
    import torch
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from deco_b200 import autograd as A
    from deco_b200 import distributed as D
    import torch.distributed as dist
    D.init_from_env("gloo")
    big = torch.arange(24, dtype=torch.float32).reshape(4, 6) * (rank + 1)       # a batched buffer announced through views
    lone = torch.full((5,), float(rank + 1))
    late = torch.full((3,), 10.0 * (rank + 1))
    with D.overlap_gradient_average(world) as avg:
        assert A.GRAD_READY_HOOK is avg
        A._grads_ready([big[1], lone])          # "block" group: big is reduced once, through its base
        A._grads_ready([big[2:], lone, late, None])   # "tail": repeats are skipped
        n_started = len(avg._work)
    assert A.GRAD_READY_HOOK is None
    q.put((rank, n_started, big.tolist(), lone.tolist(), late.tolist()))
    dist.destroy_process_group()


def test_overlapped_gradient_averager_gloo_world2():
    """The hook path the denoiser backward drives (autograd._grads_ready -> OverlappedGradientAverager): every distinct
    buffer is all-reduced exactly once (views through their base) and holds the rank average after the context exits."""
    import socket
    import torch
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_overlap_avg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n_started, big, lone, late in res:
        assert n_started == 3
        assert torch.allclose(torch.tensor(big), torch.arange(24, dtype=torch.float32).reshape(4, 6) * 1.5)
        assert lone == [1.5] * 5 and late == [15.0] * 3


def _torch_cfg_step_ex(x, net_out, g=1.0, dt=0.0, c0=1.0, prev=(), coeffs=(), dev=None, xpred_den=0.0, kd=0.0, sden=1.0,
                       a_s=0.0, a_n=0.0, noise=None, x_out=None, pred_out=None, want_pred=False, want_v=False,
                       want_u8=False, u8_out=None):
    """Test-only torch statement of the extended form of csrc/sampler.cu (deco_cfg_step_ex)."""
    u, c = net_out.float().chunk(2)
    if xpred_den > 0:
        u, c = (u - x) / xpred_den, (c - x) / xpred_den
    pred = u + g * (c - u)
    v = c0 * pred
    for p, cf in zip(prev, coeffs):
        v = v + cf * p
    xo = x + dt * v
    if a_s != 0.0:
        xo = xo + a_s * (kd * v - x) / sden
    if a_n != 0.0:
        xo = xo + a_n * noise
    return xo, (pred if want_pred else None), (v if want_v else None), (O.fp2uint8(xo) if want_u8 else None)


def test_extended_sampler_control_flow_matches_reference(monkeypatch):
    """EulerSamplerJiT and the SDE step functions: host schedule / score scalars / step-function selection against the
    fixtures from the real reference, replaying its recorded Gaussian increments through torch.randn_like."""
    from helpers import toy_xnet
    from deco_b200 import (ConstScheduler, EulerSampler, EulerSamplerJiT, GVPScheduler, HeunSampler, LinearScheduler, ode_step_fn,
                           ops, sampling, sde_mean_step_fn, sde_preserve_step_fn, sde_step_fn, simple_guidance_fn)
    monkeypatch.setattr(ops, "cfg_step", _torch_cfg_step)
    monkeypatch.setattr(ops, "cfg_step_ex", _torch_cfg_step_ex)
    monkeypatch.setattr(sampling, "_prep_inputs", lambda n, c, u: (n.float().contiguous(), torch.cat([u, c], 0)))
    g = load_golden("samplers_ext_toy.npz")
    noise = torch.from_numpy(g["noise"])
    cond, unc = torch.tensor([1, 2, 3]), torch.tensor([10, 10, 10])
    sch = LinearScheduler()
    for n, gd, lo, hi, shift in [(12, 2.5, 0.1, 1.0, 1.0), (30, 1.5, 0.0, 0.8, 2.0)]:
        s = EulerSamplerJiT(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=n, guidance=gd,
                            guidance_interval_min=lo, guidance_interval_max=hi, timeshift=shift, step_fn=ode_step_fn)
        assert rel_l2(s(toy_xnet, noise, cond, unc), torch.from_numpy(g[f"jit_{n}"])) < 1e-6
        rows, use_pred = s._graph_rows()
        assert not use_pred and len(rows) == n
        for i, r in enumerate(rows):        # column 7 = clamp_min(1 - t, 0.05) in fp32 (sampling.py:170)
            assert r[7] == float((1.0 - s.timesteps[i]).clamp_min(5e-2)) and r[6] == float(s.timesteps[i])
    fns = {"sde_mean": sde_mean_step_fn, "sde": sde_step_fn, "sde_preserve": sde_preserve_step_fn}
    for kind, fn in fns.items():
        for n, gd, shift, last in [(10, 2.0, 1.0, "ode"), (6, 1.0, 2.0, kind)]:
            s = EulerSampler(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=n, guidance=gd,
                             guidance_interval_min=0.1, guidance_interval_max=1.0, timeshift=shift, step_fn=fn,
                             last_step_fn=(ode_step_fn if last == "ode" else fn))
            assert s._graph_rows() is None          # fresh noise per step: no graph replay
            incs = list(torch.from_numpy(g[f"{kind}_{n}_increments"])) if f"{kind}_{n}_increments" in g else []
            monkeypatch.setattr(torch, "randn_like", lambda x: incs.pop(0))
            out = s(toy_net, noise, cond, unc)
            monkeypatch.undo()
            monkeypatch.setattr(ops, "cfg_step", _torch_cfg_step)
            monkeypatch.setattr(ops, "cfg_step_ex", _torch_cfg_step_ex)
            monkeypatch.setattr(sampling, "_prep_inputs", lambda n, c, u: (n.float().contiguous(), torch.cat([u, c], 0)))
            assert not incs and rel_l2(out, torch.from_numpy(g[f"{kind}_{n}"])) < 2e-6, kind
    # score scalars of the other schedulers (flow_matching/scheduling.py:17-32): GVP at t = 0.5, constant w
    e = EulerSampler(scheduler=GVPScheduler(), w_scheduler=ConstScheduler(), step_fn=sde_step_fn, num_steps=4)
    kd, sden, a_s, a_n = e._sde_scalars(torch.tensor(0.5), torch.tensor(0.25), "sde_step_fn")
    # the reference's GVP derivatives carry no pi/2 factor: dalpha/alpha = -tan(pi/4) = -1, sden = sin^2 + sin cos = 1
    assert abs(kd + 1.0) < 1e-6 and abs(sden - 1.0) < 1e-6
    assert abs(a_s - 0.25) < 1e-7 and abs(a_n - np.sqrt(0.5)) < 1e-6
    # the reference's guards: an SDE step needs a w_scheduler (sampling.py:61, :225); Heun takes the SDE step functions too
    # (scores averaged, sampling.py:283-291), an unknown step function is refused
    with pytest.raises(AssertionError):
        EulerSampler(scheduler=sch, step_fn=sde_step_fn, num_steps=2)
    with pytest.raises(AssertionError):
        HeunSampler(scheduler=sch, step_fn=sde_step_fn, num_steps=2)
    assert HeunSampler(scheduler=sch, w_scheduler=sch, step_fn=sde_step_fn, num_steps=2)._kinds == ("sde_step_fn", "ode_step_fn")
    with pytest.raises(NotImplementedError):
        HeunSampler(scheduler=sch, w_scheduler=sch, step_fn=lambda x, v, dt, s, w: x, num_steps=2)


BASELINE_JIT_YAML = """
model:
  denoiser:
    class_path: src.models.transformer.dit_c2i_baseline.FlattenDiT
    init_args: {in_channels: 3, patch_size: 16, num_groups: 4, hidden_size: 256, num_blocks: 2, num_classes: 10}
  diffusion_sampler:
    class_path: src.diffusion.flow_matching.sampling.EulerSamplerJiT
    init_args:
      num_steps: 50
      guidance: 1.0
      guidance_interval_min: 0.1
      guidance_interval_max: 1.0
      scheduler: src.diffusion.flow_matching.scheduling.LinearScheduler
      w_scheduler: src.diffusion.flow_matching.scheduling.LinearScheduler
      guidance_fn: src.diffusion.base.guidance.simple_guidance_fn
      step_fn: src.diffusion.flow_matching.sampling.sde_step_fn
      last_step_fn: src.diffusion.flow_matching.sampling.ode_step_fn
"""


def test_baseline_yaml_wiring_and_checkpoint_contract(tmp_path):
    """configs_c2i/Baseline_DiT_JiT.yaml-shaped model section through the class map; FlattenDiT's state_dict is the
    reference's (names and shapes from oracle.baseline_param_shapes, pinned by make_golden.py::golden_baseline)."""
    from deco_b200 import EulerSamplerJiT, FlattenDiT, config, ode_step_fn, sde_step_fn
    p = tmp_path / "cfg.yaml"
    p.write_text(BASELINE_JIT_YAML)
    parts = config.load_model_section(str(p))
    m, s = parts["denoiser"], parts["diffusion_sampler"]
    assert isinstance(m, FlattenDiT) and isinstance(s, EulerSamplerJiT) and s.x_prediction
    assert s.step_fn is sde_step_fn and s.last_step_fn is ode_step_fn and s._kinds == ("sde_step_fn", "ode_step_fn")
    want = O.baseline_param_shapes(O.BaselineCfg(num_groups=4, hidden_size=256, num_blocks=2, num_classes=10))
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    # default init zeroes the whole output layer (dit_c2i_baseline.py:351-355)
    assert float(m.final_layer.linear.weight.abs().max()) == 0.0 and float(m.final_layer.adaLN_modulation[0].weight.abs().max()) == 0.0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.eval()(torch.zeros(1, 3, 32, 32), torch.zeros(1), torch.zeros(1, dtype=torch.long))


def test_checkpoint_layout_and_prefix_cleaning(tmp_path):
    """src/lightning_model.py:322-368: checkpoints hold `denoiser.*` then `ema_denoiser.*` (the trainer contributes nothing,
    training_repa_DeCo.py:290-291); keys written under torch.compile (`_orig_mod.`) or DDP (`.module.`) load into the bare
    modules; the optimizer state travels in torch.optim.AdamW's layout; `ModelLoader` reads the same file."""
    import copy
    from deco_b200 import LinearScheduler, PixNerDiT, REPATrainer
    from deco_b200.io import ModelLoader, clean_checkpoint_keys, lightning_state_dict, load_checkpoint, save_checkpoint
    kw = dict(in_channels=3, num_groups=2, hidden_size=64, hidden_size_x=32, num_blocks=3, num_cond_blocks=1, patch_size=16,
              num_classes=10)
    torch.manual_seed(0)
    net = PixNerDiT(**kw)
    ema = copy.deepcopy(net)
    with torch.no_grad():
        for p in ema.parameters():
            p.add_(0.25)
    tr = REPATrainer(scheduler=LinearScheduler())
    sd = lightning_state_dict(net, ema, tr)
    names = list(net.state_dict())
    assert list(sd) == ["denoiser." + n for n in names] + ["ema_denoiser." + n for n in names]
    assert not any(k.startswith("diffusion_trainer.") for k in sd)

    class FakeOpt:      # torch.optim.AdamW layout, as FusedAdamWEMA.state_dict() (GPU) produces it
        def __init__(self):
            self.loaded = None

        def state_dict(self):
            return dict(state={0: dict(step=torch.tensor(3.0), exp_avg=torch.ones(2), exp_avg_sq=torch.ones(2))},
                        param_groups=[dict(lr=1e-4, params=[0])])

        def load_state_dict(self, s):
            self.loaded = s
    path = save_checkpoint(str(tmp_path / "last.ckpt"), net, ema, tr, FakeOpt(), global_step=7, ema_decay=0.999)
    ck = torch.load(path, map_location="cpu")
    assert ck["global_step"] == 7 and ck["callbacks"]["SimpleEMA"] == dict(decay=0.999, every_n_steps=1)
    assert list(ck["state_dict"]) == list(sd) and float(ck["optimizer_states"][0]["state"][0]["step"]) == 3.0
    # the reference's own writers: torch.compile'd denoiser under DDP
    dirty = {}
    for k, v in ck["state_dict"].items():
        head, rest = k.split(".", 1)
        dirty[f"{head}.module._orig_mod.{rest}" if head == "denoiser" else f"{head}._orig_mod.{rest}"] = v
    assert list(clean_checkpoint_keys(dirty)) == list(sd)
    net2, ema2, opt2 = PixNerDiT(**kw), PixNerDiT(**kw), FakeOpt()
    out = load_checkpoint(dict(ck, state_dict=dirty), net2, ema2, opt2)
    assert out["global_step"] == 7 and opt2.loaded is not None
    for a, b in ((net, net2), (ema, ema2)):
        for (n1, p1), (n2, p2) in zip(a.state_dict().items(), b.state_dict().items()):
            assert n1 == n2 and torch.equal(p1, p2)
    # the sampling side reads the same file (src/utils/model_loader.py:14-27; app.py:56-63 always takes the EMA weights)
    net3 = PixNerDiT(weight_path=path, load_ema=True, **kw)
    ModelLoader().load(net3)
    assert torch.equal(net3.blocks[0].attn.qkv.weight, ema.blocks[0].attn.qkv.weight)
    with pytest.raises(RuntimeError):
        load_checkpoint(dict(state_dict={"denoiser.nope": torch.zeros(1)}), PixNerDiT(**kw))


def test_pixnerd_state_dict_is_the_reference_checkpoint_contract():
    """dit_c2i_pixnerd.PixNerDiT (configs_c2i/Baseline_PixNerd.yaml): DiT blocks and NerfBlocks share one `blocks` list; names
    and shapes as instantiated from the reference (oracle.pixnerd_param_shapes, pinned by tests/golden/pixnerd_d64.npz)."""
    from deco_b200 import config
    from deco_b200.denoiser_pixnerd import PixNerDiT
    assert config.resolve("src.models.transformer.dit_c2i_pixnerd.PixNerDiT") is PixNerDiT
    cfg = O.PixNerdCfg()
    with torch.device("meta"):
        m = PixNerDiT(in_channels=3, patch_size=16, num_groups=16, hidden_size=1024, hidden_size_x=64, num_blocks=24,
                      num_cond_blocks=22, nerf_mlpratio=2, num_classes=1000)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == O.pixnerd_param_shapes(cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PixNerDiT(in_channels=3, patch_size=16, num_groups=2, hidden_size=128, hidden_size_x=64, num_blocks=2, num_cond_blocks=1,
                  nerf_mlpratio=2, num_classes=10).eval()(torch.zeros(1, 3, 16, 16), torch.zeros(1), torch.zeros(1, dtype=torch.long))
