#!/usr/bin/env python
"""Benchmark of the DeCo sampling hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): DeCo-XL/16, 256 px class-conditional sampling, global batch 256 sharded over the
N ranks (strong scaling), Euler 100 steps x CFG (= 200 network evaluations per image), guidance 3.2 on (0.1, 1].
A *step* is one denoiser step: one CFG-batched forward of 2 x B_local rows plus the fused guidance/Euler update.
`value` = images/s at the fixed NFE = B_global / (100 x seconds per step); inputs are resident in HBM.
`e2e`   = the same metric through the public sampler API (EulerSampler.sample_uint8) with pinned-host noise/labels
          copied in and the uint8 images copied out (and all-gathered for N > 1) inside the timed region.
Synthetic data: seeded CPU randn noise per sample (src/data/dataset/randn.py:74-75), labels cycling 0..999, seeded
random-init weights (deco_b200.utils.randomize_: every tensor non-zero).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

XL = dict(in_channels=3, num_groups=16, hidden_size=1152, hidden_size_x=32, num_blocks=31, num_cond_blocks=28,
          patch_size=16, num_classes=1000, nerf_mlpratio=2)
LARGE = dict(XL, hidden_size=1024, num_blocks=25, num_cond_blocks=22)
XXL_T2I = dict(in_channels=3, patch_size=16, num_groups=24, hidden_size=1536, txt_embed_dim=2048, txt_max_length=128,
               num_text_blocks=4, decoder_hidden_size=32, num_encoder_blocks=16, num_decoder_blocks=3)
BASELINE_JIT = dict(in_channels=3, patch_size=16, num_groups=16, hidden_size=1024, num_blocks=24, num_classes=1000)
PIXNERD = dict(in_channels=3, patch_size=16, num_groups=16, hidden_size=1024, hidden_size_x=64, num_blocks=24,
               num_cond_blocks=22, nerf_mlpratio=2, num_classes=1000)
UNIT = "images/s"
# BASELINE.json configs; "xl256" (configs[1]) is the one the headline metric is quoted on and the default.
# gflop = algorithmic GFLOP per image-forward (2*MAC, no padding counted).
WORKLOADS = {
    "xl256": dict(kind="c2i", model=XL, res=256, batch=256, sampler="euler", steps=100, guidance=3.2, gmin=0.1, gmax=1.0,
                  timeshift=1.0, gflop=244.9,
                  metric="DeCo-XL/16 256px class-conditional sampling throughput (Euler 100 steps x CFG, fixed NFE)",
                  name="DeCo-XL/16 256px c2i (configs_c2i/DeCo_XL.yaml), Euler 100 steps x CFG 3.2 on (0.1,1]"),
    "xl512": dict(kind="c2i", model=XL, res=512, batch=64, sampler="euler", steps=100, guidance=5.0, gmin=0.1, gmax=1.0,
                  timeshift=1.0, gflop=1079.9,
                  metric="DeCo-XL/16 512px class-conditional sampling throughput (Euler 100 steps x CFG, fixed NFE)",
                  name="DeCo-XL/16 512px c2i (configs_c2i/DeCo_XL_512.yaml), Euler 100 steps x CFG 5.0 on (0.1,1]"),
    "l256": dict(kind="c2i", model=LARGE, res=256, batch=256, sampler="euler", steps=100, guidance=3.2, gmin=0.1, gmax=1.0,
                 timeshift=1.0, gflop=155.0,
                 metric="DeCo-L/16 256px class-conditional sampling throughput (Euler 100 steps x CFG, fixed NFE)",
                 name="DeCo-L/16 256px c2i (configs_c2i/DeCo_large.yaml architecture), Euler 100 steps x CFG 3.2 on (0.1,1]"),
    "t2i512": dict(kind="t2i", model=XXL_T2I, res=512, batch=32, sampler="adam2", steps=25, guidance=4.0, gmin=0.0, gmax=1.0,
                   timeshift=3.0, gflop=1450.6,
                   metric="DeCo-XXL/16 512px text-to-image sampling throughput (AdamLM order 2, 25 steps x CFG, fixed NFE)",
                   name="DeCo-XXL/16 512px t2i (configs_t2i/sft_res512.yaml), AdamLM order 2, 25 steps x CFG 4.0, "
                        "timeshift 3, synthetic text-encoder states [128 x 2048]"),
    # the patch-linear baseline (FlattenDiT) with the x-prediction Euler sampler; not a BASELINE.json config
    "jit256": dict(kind="baseline", model=BASELINE_JIT, res=256, batch=256, sampler="euler_jit", steps=50, guidance=1.0,
                   gmin=0.1, gmax=1.0, timeshift=1.0, gflop=162.2,
                   metric="Baseline DiT-L/16 (patch-linear head, x-prediction) 256px sampling throughput (EulerSamplerJiT 50 "
                          "steps x CFG, fixed NFE)",
                   name="FlattenDiT 1024 x 24 blocks /16 256px c2i (configs_c2i/Baseline_DiT_JiT.yaml denoiser and sampler "
                        "settings), EulerSamplerJiT 50 steps x CFG rows (guidance 1.0 as configured)"),
    # the PixNerd baseline (hyper-network NerfBlock decoder); not a BASELINE.json config.  GFLOP per
    # image-forward: 22 DiT blocks of the L/16 architecture 147.6 + parameter generators 2 x 2 x 1024 x 16384 x 256 tokens =
    # 17.2 + per-pixel MLPs 2 x 2 x (2 x 64 x 128) x 65536 = 4.3 + embedders / final layer 0.6
    "pixnerd256": dict(kind="pixnerd", model=PIXNERD, res=256, batch=256, sampler="euler", steps=50, guidance=1.0, gmin=0.1,
                       gmax=1.0, timeshift=1.0, gflop=169.7,
                       metric="PixNerd-L/16 (hyper-network pixel decoder) 256px sampling throughput (Euler 50 steps x CFG rows, "
                              "fixed NFE)",
                       name="PixNerDiT 1024 x (22 DiT + 2 NerfBlock) /16 256px c2i (configs_c2i/Baseline_PixNerd.yaml denoiser and "
                            "sampler settings), Euler 50 steps x CFG rows (guidance 1.0 as configured)"),
    # BASELINE configs[3]: training step (forward + backward of denoiser and DCT/FM loss), 32 images PER GPU (weak scaling)
    "train256": dict(kind="train", model=XL, res=256, batch=32, gflop=3 * 244.9,
                     metric="DeCo-XL/16 256px training step throughput (denoiser + DCT/FM loss, forward + backward)",
                     name="DeCo-XL/16 256px training step (configs_c2i/DeCo_XL.yaml trainer: REPATrainer with the 8x8 "
                          "block-DCT loss, null_condition_p 0.2), 32 synthetic images per GPU, forward + backward, "
                          "no optimizer step"),
}
# the reference arm / cpu_baseline always time the headline workload's CPU restatement
NUM_SAMPLING_STEPS = WORKLOADS["xl256"]["steps"]
RES = WORKLOADS["xl256"]["res"]
GUIDANCE, G_MIN, G_MAX = 3.2, 0.1, 1.0
METRIC = WORKLOADS["xl256"]["metric"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="deco_b200", choices=["deco_b200", "reference"])
    ap.add_argument("--workload", default="xl256", choices=sorted(WORKLOADS))
    ap.add_argument("--global-batch", type=int, default=0, help="0 = the workload's batch (BASELINE.json)")
    ap.add_argument("--no-hbm-kernels", action="store_true", help="skip the per-kernel HBM roofline section")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-steps", type=int, default=2)
    ap.add_argument("--cpu-batch", type=int, default=4, help="images per CFG-batched step of the CPU arm (configs[0] uses 4)")
    ap.add_argument("--torch-baseline", default="both", choices=["none", "eager", "both"],
                    help="time the reference's own GPU path (PyTorch bf16 autocast; eager and torch.compile) on this GPU")
    ap.add_argument("--torch-baseline-batch", type=int, default=0, help="images per step of that arm (0 = the bench batch)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity check against the fp32 oracle")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (use with ncu --profile-from-start off)")
    return ap.parse_args()


def gemm_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch, averaged over the GEMM launches of ONE step, and the
    per-kernel table it comes from: parsed from the committed ncu summary profiles/traffic_r2_<workload>.txt (written by
    profiles/agg_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` over
    `bench.py --workload <workload> --steps 1 --profile`).  (None, None) until that capture exists for the workload."""
    path = os.path.join(ROOT, "profiles", f"traffic_r2_{workload}.txt")
    if not os.path.exists(path):
        return None, None
    for ln in open(path):
        if ln.startswith("#json "):
            d = json.loads(ln[6:])
            top = sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms"])[:8]
            return d["gemm_dram_bytes_per_launch"], dict(source=os.path.relpath(path, ROOT), gemm_launches=d["gemm_launches"],
                                                         kernels={k: v for k, v in top})
    return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs"), tf_burst=d.get("bf16_tflops"), tf_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock / throttle reasons sampled during the timed region (NVML, 100 ms period)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.nv:
            self.th.join(2)

    def report(self):
        if not self.samples:
            return None
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2], sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(s))


def labels_for(indices):
    return [i % 1000 for i in indices]


# ----------------------------------------------------------------------------------------------- reference (CPU) arm
def cpu_reference_step_seconds(state_dict_cpu, steps, warmup, batch=1):
    """The reference algorithm (oracle port, fp32, all host threads): seconds per CFG-batched denoiser step of `batch`
    images, i.e. one `net(cat[x,x], t, cat[uncond,cond])` + guidance + Euler update."""
    import torch
    from oracle import deco_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.CFG_XL
    P = state_dict_cpu if state_dict_cpu is not None else O.seeded_params(cfg)
    x = torch.stack([torch.randn((3, RES, RES), generator=torch.Generator().manual_seed(i)) for i in range(batch)])
    cond = torch.tensor(labels_for(range(batch)))
    unc = torch.full((batch,), 1000)
    ts = O.make_timesteps(NUM_SAMPLING_STEPS)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = O.denoiser_forward(P, cfg, torch.cat([x, x]), ts[i % NUM_SAMPLING_STEPS].repeat(2 * batch), torch.cat([unc, cond]))
            g = GUIDANCE if (ts[i % NUM_SAMPLING_STEPS] > G_MIN and ts[i % NUM_SAMPLING_STEPS] <= G_MAX) else 1.0
            x = x + O.cfg_combine(out, g) * (ts[i % NUM_SAMPLING_STEPS + 1] - ts[i % NUM_SAMPLING_STEPS])
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.cpu_batch
    sec, cores = cpu_reference_step_seconds(None, max(1, args.steps), max(0, args.warmup), batch)
    value = batch / (sec * NUM_SAMPLING_STEPS)
    sample = (f"{batch} images, {args.steps} CFG-batched Euler steps of the 100 timed after {args.warmup} warm-up; "
              f"images/s extrapolated linearly to 100 steps")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="fp32", data="synthetic",
                config=dict(workload=f"DeCo-XL/16 256px c2i, Euler 100 x CFG 3.2 on (0.1,1]; CPU sample batch {batch}",
                            global_batch=batch, num_sampling_steps=NUM_SAMPLING_STEPS),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def _oracle_forward(wl, net, dev):
    """(fn(x, t, cond) -> velocity, fp32 parameter dict): the oracle restatement of this workload's denoiser, holding the
    bench network's own weights (state_dict keys are the reference's, which is what the oracle indexes by)."""
    import torch
    from oracle import deco_oracle as O
    P = {k: v.detach().to(device=dev, dtype=torch.float32) for k, v in net.state_dict().items()}
    m = wl["model"]
    if wl["kind"] == "t2i":
        cfg = O.T2ICfg(in_channels=m["in_channels"], num_groups=m["num_groups"], hidden_size=m["hidden_size"],
                       decoder_hidden_size=m["decoder_hidden_size"], num_encoder_blocks=m["num_encoder_blocks"],
                       num_decoder_blocks=m["num_decoder_blocks"], num_text_blocks=m["num_text_blocks"],
                       patch_size=m["patch_size"], txt_embed_dim=m["txt_embed_dim"], txt_max_length=m["txt_max_length"])
        return (lambda x, t, c: O.t2i_forward(P, cfg, x, t, c.float())), P
    if wl["kind"] == "pixnerd":
        cfg = O.PixNerdCfg(in_channels=m["in_channels"], num_groups=m["num_groups"], hidden_size=m["hidden_size"],
                           hidden_size_x=m["hidden_size_x"], nerf_mlpratio=m["nerf_mlpratio"], num_blocks=m["num_blocks"],
                           num_cond_blocks=m["num_cond_blocks"], patch_size=m["patch_size"], num_classes=m["num_classes"])
        return (lambda x, t, c: O.pixnerd_forward(P, cfg, x, t, c)), P
    if wl["kind"] == "baseline":
        cfg = O.BaselineCfg(in_channels=m["in_channels"], num_groups=m["num_groups"], hidden_size=m["hidden_size"],
                            num_blocks=m["num_blocks"], patch_size=m["patch_size"], num_classes=m["num_classes"])
        return (lambda x, t, c: O.baseline_forward(P, cfg, x, t, c)), P
    cfg = O.DenoiserCfg(in_channels=m["in_channels"], num_groups=m["num_groups"], hidden_size=m["hidden_size"],
                        hidden_size_x=m["hidden_size_x"], num_blocks=m["num_blocks"], num_cond_blocks=m["num_cond_blocks"],
                        patch_size=m["patch_size"], num_classes=m["num_classes"])
    return (lambda x, t, c: O.denoiser_forward(P, cfg, x, t, c)), P


def parity_check(torch, wl, net, dev, x, t_cur, cfg_cond, B, tol=1e-2):
    """The number and the parity evidence on the SAME shape: one denoiser evaluation of the full bench batch (2B CFG rows)
    through the product path, and the fp32 oracle (checker only) on image 0's two CFG rows [uncond, cond]; rel-L2 of
    those two rows.  The bench aborts above the north_star tolerance (1e-2)."""
    fwd, _ = _oracle_forward(wl, net, dev)
    tt = torch.full((2 * B,), float(t_cur), device=dev)
    with torch.no_grad():
        out = net(torch.cat([x, x]), tt, cfg_cond).float()
        rows = [0, B]
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            ref = fwd(x[[0, 0]].float(), tt[rows], cfg_cond[rows]).float()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
    got = out[rows]
    e = float((got.double() - ref.double()).norm() / ref.double().norm())
    res = dict(rel_l2=e, tol=tol, rows="image 0: [uncond, cond] of the full-batch forward", t=float(t_cur),
               against="fp32 oracle (oracle/deco_oracle.py) on the same GPU, same weights", ok=bool(e <= tol))
    if not e <= tol:
        raise SystemExit(f"bench: parity check failed: rel-L2 {e:.3e} > {tol} -- refusing to report a throughput")
    return res


def torch_gpu_baseline(torch, args, wl, net, dev, x, t_cur, cfg_cond, B, nsteps):
    """The reference's own GPU path on THIS GPU (src/lightning_model.py:94-97 compiles the denoiser,
    src/diffusion/base/sampling.py:27 runs it under bf16 autocast): the oracle issues the reference's F.linear / SDPA /
    layer_norm calls, so under torch.autocast(bf16) it is the library-kernel implementation (cuBLAS, flash SDPA, eager or
    Inductor element-wise kernels) of the same step.  One CFG-batched denoiser evaluation + guidance + Euler update per step,
    the same batch as the timed product path when it fits."""
    import time as _time
    from oracle import deco_oracle as O
    fwd, _ = _oracle_forward(wl, net, dev)
    Bt = args.torch_baseline_batch or B
    Bt = min(Bt, B)
    xb = x[:Bt].clone()
    cond = torch.cat([cfg_cond[:Bt], cfg_cond[B:B + Bt]])
    tt = torch.full((2 * Bt,), float(t_cur), device=dev)
    g = float(wl["guidance"])

    def step(f):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = f(torch.cat([xb, xb]), tt, cond)
        return xb + O.cfg_combine(out.float(), g) * 0.01

    def timed(f, iters):
        for _ in range(2):
            step(f)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step(f)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    res = dict(batch=Bt, cfg_rows=2 * Bt, dtype="bf16 autocast (fp32 master weights)", unit=UNIT,
               note="oracle = the reference's torch calls; value = batch / (num_sampling_steps x seconds per step)")
    try:
        ms = timed(fwd, 3)
        res["eager"] = dict(ms_per_step=ms, value=Bt / (ms * 1e-3 * nsteps))
    except Exception as e:   # noqa: BLE001  (e.g. out of memory at the full batch)
        res["eager"] = dict(error=str(e)[:200])
    if args.torch_baseline == "both":
        try:
            # Inductor's C++ wrapper is built with -fopenmp; this image keeps libgomp.spec under gcc's own library directory,
            # which /opt/gcc/bin/g++ does not search by itself
            gomp = "/usr/lib/gcc/x86_64-linux-gnu/13"
            if os.path.exists(os.path.join(gomp, "libgomp.spec")) and gomp not in os.environ.get("LIBRARY_PATH", ""):
                os.environ["LIBRARY_PATH"] = gomp + (":" + os.environ["LIBRARY_PATH"] if os.environ.get("LIBRARY_PATH") else "")
            t0 = _time.perf_counter()
            cf = torch.compile(fwd)
            ms = timed(cf, 3)
            res["compiled"] = dict(ms_per_step=ms, value=Bt / (ms * 1e-3 * nsteps),
                                   compile_seconds=_time.perf_counter() - t0)
        except Exception as e:   # noqa: BLE001
            res["compiled"] = dict(error=str(e)[:200])
    return res


# ----------------------------------------------------------------------------------------------- deco_b200 arm
def _time_kernel(fn, iters, torch):
    """Average milliseconds per call of fn(i): `iters` calls captured into one CUDA graph (so that host-side launch cost
    does not pace kernels that run for tens of microseconds), replayed once after a warm replay, CUDA events around it."""
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def hbm_kernel_rooflines(torch, ops, net, dev, B2, res, hbm_gbs):
    """Achieved HBM GB/s of the memory-bound kernels at the bench shape (B2 CFG rows), each timed alone with CUDA events
    over buffers far larger than L2 (or rotated through > L2 of distinct buffers).  Algorithmic bytes per DESIGN.md."""
    from deco_b200.training import build_freq_weight
    bf = torch.bfloat16
    out = []
    H = net.hidden_size
    L = (res // net.patch_size) ** 2
    M = B2 * L
    B = B2 // 2
    npx = 3 * res * res

    def add(name, ms, nbytes, note, flops=None):
        d = dict(kernel=name, bound="hbm", ms=ms, algorithmic_bytes=nbytes, achieved=nbytes / (ms * 1e-3) / 1e9,
                 peak=hbm_gbs, unit="GB/s", frac=nbytes / (ms * 1e-3) / 1e9 / hbm_gbs, note=note)
        if flops:
            d["tflops"] = flops / (ms * 1e-3) / 1e12
        out.append(d)

    # sampler update: x fp32 r/w + two bf16 network rows = 12 B per element
    x = torch.randn(B, 3, res, res, device=dev)
    v = torch.randn(2 * B, 3, res, res, device=dev).to(bf)
    xo = torch.empty_like(x)
    ms = _time_kernel(lambda i: ops.cfg_step(x, v, 3.2, 0.01, x_out=xo), 20, torch)
    add("cfg_step_kernel", ms, 12.0 * B * npx, "x fp32 read+write, uncond/cond bf16 read")
    # rmsnorm + modulate on the fp32 stream: 4 B read + 2 B write per element
    s = torch.randn(M, H, device=dev)
    mod = torch.randn(B2, 2 * H, device=dev).to(bf)
    w = torch.ones(H, device=dev)
    hb = torch.empty(M, H, device=dev, dtype=bf)
    ms = _time_kernel(lambda i: ops.rmsnorm_modulate(s, w, mod[:, :H], mod[:, H:], L, out=hb), 20, torch)
    add("rmsnorm_modulate_kernel", ms, 6.0 * M * H, "fp32 stream read, bf16 write")
    del s, hb
    # q/k head norm + RoPE in place: q,k bf16 read + write = 8 B per (q,k) element pair -> 2*2*2H B per row
    d = H // net.num_groups
    qkv = torch.randn(M, 3 * H, device=dev).to(bf)
    pos = net.fetch_pos(res // net.patch_size, res // net.patch_size, dev)
    ms = _time_kernel(lambda i: ops.qknorm_rope_(qkv, w[:d], w[:d], pos, net.num_groups, d, L), 20, torch)
    add("qknorm_rope_kernel", ms, 8.0 * M * H, "q and k bf16, read + write in place")
    del qkv
    # fused pixel decoder: x fp32 12 B + condition 64 B + out bf16 6 B per pixel; 37 kFLOP per pixel
    P = net.prepare(dev)
    ycond = torch.randn(M, net.patch_size ** 2 * 32, device=dev).to(bf)
    xx = torch.randn(B2, 3, res, res, device=dev)
    nres = net.num_blocks - net.num_cond_blocks if hasattr(net, "num_cond_blocks") else net.num_decoder_blocks
    if "blob_tc" in P:      # the decoder the sampling step runs: tcgen05 kernel, sampler update fused into its epilogue
        xh = xx[:B].contiguous()
        xo = torch.empty_like(xh)
        ms = _time_kernel(lambda i: ops.pixel_decoder_tc_step(xh, ycond, P["blob_tc"], net.patch_size, 32, nres, g=3.2, dt=0.01,
                                                              x_out=xo), 5, torch)
        add("pixel_decoder_tc_kernel (CFG + Euler update fused)", ms, (64.0 * B2 + 8.0 * B) * res * res,
            "64 B/pixel condition per CFG row + x fp32 read/write per image; tensor/issue-bound (DESIGN.md 4): tflops on "
            "37.2 kFLOP/pixel", flops=37.2e3 * B2 * res * res)
        ms = _time_kernel(lambda i: ops.pixel_decoder_tc(xx, ycond, P["blob_tc"], net.patch_size, 32, nres), 5, torch)
        add("pixel_decoder_tc_kernel (plain)", ms, 82.0 * B2 * res * res, "82 B/pixel (64 condition + 12 x + 6 out)",
            flops=37.2e3 * B2 * res * res)
    ms = _time_kernel(lambda i: ops.pixel_decoder(xx, ycond, P["blob"], P["postab"], net.patch_size, 32, nres), 5, torch)
    add("pixel_decoder_kernel (legacy mma.sync, A/B)", ms, 82.0 * B2 * res * res,
        "82 B/pixel (64 condition + 12 x + 6 out); compute/issue-bound once fused (DESIGN.md): tflops on 37.2 kFLOP/pixel",
        flops=37.2e3 * B2 * res * res)
    del ycond, xx
    # DCT + FM loss, forward + backward in one launch: out, v_t read + grad write fp32 = 12 B per element.
    # BASELINE.json configs[3] is 32 images per GPU = 25 MB per tensor (L2-resident, ~10 us at the roofline: launch and
    # tail latency weigh in), so 8 distinct input sets rotate; a 256-image batch shows the streaming rate.
    fw = build_freq_weight(85).reshape(3, 8, 8).contiguous().to(dev)
    for Bt, nset, iters in ((32, 8, 24), (256, 2, 6)):
        outs = [torch.randn(Bt, 3, res, res, device=dev) for _ in range(nset)]
        vts = [torch.randn(Bt, 3, res, res, device=dev) for _ in range(nset)]
        ms = _time_kernel(lambda i: ops.dct_fm_loss(outs[i % nset], vts[i % nset], fw, 1.0, want_loss=True, want_grad=True),
                          iters, torch)
        add(f"dct_fm_loss_vec_kernel (fwd+bwd, {Bt} images)", ms, 12.0 * Bt * npx,
            f"out + v_t fp32 read, grad fp32 write, {nset} rotating input sets")
        ms = _time_kernel(lambda i: ops.dct_fm_loss(outs[i % nset], vts[i % nset], fw, 1.0, want_loss=True, want_grad=False),
                          iters, torch)
        add(f"dct_fm_loss_vec_kernel (fwd, {Bt} images)", ms, 8.0 * Bt * npx, f"out + v_t fp32 read, {nset} rotating input sets")
        del outs, vts
    return out


def baseline_head_rooflines(torch, ops, net, dev, B2, res, hbm_gbs):
    """Achieved HBM GB/s of the patch-linear head's memory-bound kernels and of the extended sampler update at the bench
    shape (jit256 workload), timed like hbm_kernel_rooflines."""
    bf = torch.bfloat16
    out = []
    H, p = net.hidden_size, net.patch_size
    M = B2 * (res // p) ** 2
    B = B2 // 2
    npx = 3 * res * res

    def add(name, ms, nbytes, note):
        out.append(dict(kernel=name, bound="hbm", ms=ms, algorithmic_bytes=nbytes, achieved=nbytes / (ms * 1e-3) / 1e9,
                        peak=hbm_gbs, unit="GB/s", frac=nbytes / (ms * 1e-3) / 1e9 / hbm_gbs, note=note))

    x = torch.randn(B, 3, res, res, device=dev)
    v = torch.randn(2 * B, 3, res, res, device=dev).to(bf)
    z = torch.randn_like(x)
    xo = torch.empty_like(x)
    ms = _time_kernel(lambda i: ops.cfg_step_ex(x, v, 1.0, 0.02, xpred_den=0.5, x_out=xo), 20, torch)
    add("cfg_step_kernel<ext> (x-prediction)", ms, 12.0 * B * npx, "x fp32 read+write, uncond/cond bf16 read")
    ms = _time_kernel(lambda i: ops.cfg_step_ex(x, v, 1.0, 0.02, kd=0.5, sden=0.5, a_s=0.01, a_n=0.2, noise=z, x_out=xo), 20, torch)
    add("cfg_step_kernel<ext> (sde_step_fn)", ms, 16.0 * B * npx, "x fp32 read+write, noise fp32 read, uncond/cond bf16 read")
    del x, v, z, xo
    s = torch.randn(M, H, device=dev)
    mod = torch.randn(B2, 2 * H, device=dev).to(bf)
    hb = torch.empty(M, H, device=dev, dtype=bf)
    ms = _time_kernel(lambda i: ops.layernorm_modulate(s, mod[:, :H], mod[:, H:], M // B2, out=hb), 20, torch)
    add("layernorm_modulate_kernel", ms, 6.0 * M * H, "fp32 stream read, bf16 write")
    del s, hb
    toks = [torch.randn(M, 3 * p * p, device=dev).to(bf) for _ in range(2)]
    ms = _time_kernel(lambda i: ops.unpatchify(toks[i % 2], B2, 3, res, res, p), 20, torch)
    add("unpatchify_kernel", ms, 4.0 * B2 * npx, "bf16 tokens read, bf16 image write, 2 rotating inputs")
    return out


def run_deco(args):
    import torch
    import torch.distributed as dist
    from deco_b200 import (AdamLMSampler, EulerSampler, EulerSamplerJiT, FlattenDiT, LinearScheduler, PixNerDiT, _lib,
                           ode_step_fn, ops, simple_guidance_fn)
    from deco_b200 import distributed as D
    from deco_b200.data import rank_indices, seeded_noise
    from deco_b200.denoiser_t2i import PixNerDiT as PixNerDiTT2I
    from deco_b200.denoiser_pixnerd import PixNerDiT as PixNerdDiT
    from deco_b200.utils import GemmProbe, randomize_

    wl = WORKLOADS[args.workload]
    nsteps, res = wl["steps"], wl["res"]
    gbatch = args.global_batch or wl["batch"]
    rank, world, local = D.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: deco_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()
    assert world == args.gpus or world == 1, (world, args.gpus)

    idx = rank_indices(gbatch, rank, world)          # DistributedSampler(shuffle=False) shard
    B = len(idx)
    with torch.device("meta"):
        net = {"t2i": PixNerDiTT2I, "baseline": FlattenDiT, "pixnerd": PixNerdDiT}.get(wl["kind"], PixNerDiT)(**wl["model"])
    net = randomize_(net.to_empty(device=dev), seed=0).eval()
    net.prepare(dev)
    sch = LinearScheduler()
    if wl["sampler"] in ("euler", "euler_jit"):
        cls = EulerSamplerJiT if wl["sampler"] == "euler_jit" else EulerSampler
        sampler = cls(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=nsteps,
                      guidance=wl["guidance"], guidance_interval_min=wl["gmin"], guidance_interval_max=wl["gmax"],
                      timeshift=wl["timeshift"], step_fn=ode_step_fn)
    else:
        sampler = AdamLMSampler(order=2, timeshift=wl["timeshift"], scheduler=sch, guidance_fn=simple_guidance_fn,
                                num_steps=nsteps, guidance=wl["guidance"], guidance_interval_min=wl["gmin"],
                                guidance_interval_max=wl["gmax"])
    noise_host = seeded_noise(idx, (3, res, res))                # pinned host batch
    if wl["kind"] == "t2i":
        # synthetic text-encoder states: one seeded [T, D] block per prompt, one shared block for the null prompt
        T_, D_ = wl["model"]["txt_max_length"], wl["model"]["txt_embed_dim"]
        cond_host = torch.stack([torch.randn((T_, D_), generator=torch.Generator().manual_seed(10_000 + i))
                                 for i in idx]).to(torch.bfloat16).pin_memory()
        unc_row = torch.randn((T_, D_), generator=torch.Generator().manual_seed(9_999)).to(torch.bfloat16)
        make_unc = lambda: unc_row.to(dev).unsqueeze(0).repeat(B, 1, 1)   # noqa: E731
    else:
        cond_host = torch.tensor(labels_for(idx), dtype=torch.int64).pin_memory()
        make_unc = lambda: torch.full((B,), 1000, dtype=torch.int64, device=dev)   # noqa: E731
    x = noise_host.to(dev, non_blocking=True)
    cond = cond_host.to(dev, non_blocking=True)
    cfg_cond = torch.cat([make_unc(), cond])
    ts = sampler.timesteps
    state = dict(pred=None)

    stepper = sampler.graphed_stepper(net, x, cfg_cond)
    if stepper is not None:
        stepper.reset(x, cfg_cond)

    def one_step(x, i):
        """One sampling step through the sampler's CUDA-graphed stepper (what EulerSampler runs), else eagerly."""
        if stepper is not None:
            stepper.step()
            return stepper.x
        return one_step_eager(x, i)

    def one_step_eager(x, i):
        k = i % nsteps
        t_cur, t_next = ts[k], ts[k + 1]
        out = net(torch.cat([x, x]), torch.full((2 * B,), float(t_cur), device=dev), cfg_cond)
        if wl["sampler"] in ("euler", "euler_jit"):
            g = wl["guidance"] if (bool(t_cur > wl["gmin"]) and bool(t_cur <= wl["gmax"])) else 1.0
            if wl["sampler"] == "euler_jit":
                return ops.cfg_step_ex(x, out, g, float(t_next - t_cur), xpred_den=sampler._xpred_den(t_cur))[0]
            return ops.cfg_step(x, out, g, float(t_next - t_cur))[0]
        g = wl["guidance"] if (bool(t_cur > wl["gmin"]) and bool(t_cur < wl["gmax"])) else 1.0
        cs = sampler.solver_coeffs[k] if state["pred"] is not None else (1.0,)
        prev = (state["pred"],) if len(cs) > 1 else ()
        xn, pred, _, _ = ops.cfg_step(x, out, g, float(t_next - t_cur), c0=cs[-1], prev=prev, coeffs=tuple(cs[:-1]),
                                      want_pred=True)
        state["pred"] = pred
        return xn

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if not args.no_parity and not args.profile:
        parity = parity_check(torch, wl, net, dev, x, ts[min(len(ts) - 2, nsteps // 3)], cfg_cond, B)
    for i in range(args.warmup):
        x = one_step(x, i)
    barrier()
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        if args.profile:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(args.steps):
            x = one_step(x, args.warmup + i)
        e1.record()
        barrier()
        if args.profile:
            torch.cuda.profiler.stop()
    launches = _lib.launch_count - launches0
    ms_total = e0.elapsed_time(e1)
    # roofline pass: the same steps again with every GEMM launch bracketed by CUDA events on the launch stream (kept out of
    # the timed region above: ~350 event records per step serialise back-to-back launches, 3 % of a 17 ms step at 8 GPUs)
    probe = GemmProbe()
    if not args.profile:
        ops.gemm_probe = probe
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        xe = x.clone()
        for i in range(args.steps):
            xe = one_step_eager(xe, args.warmup + args.steps + i)
        p1.record()
        barrier()
        ops.gemm_probe = None
        ms_probe_total = p0.elapsed_time(p1)
    else:
        ms_probe_total = ms_total
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt)
    ms_per_step = ms_total / args.steps
    value = gbatch / (ms_per_step * 1e-3 * nsteps)
    gs = probe.summary()
    peaks = measured_peaks()
    traffic, traffic_tab = gemm_traffic(args.workload)

    # ---- e2e: public sampler API, host buffers in, uint8 images out (+ all-gather)
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty((gbatch if world > 1 else B, 3, res, res), dtype=torch.uint8).pin_memory()
        sampler.graphed_stepper(net, x, cfg_cond, to_uint8=True)   # capture the uint8 variant outside the timed region
        if world > 1:                   # NCCL sets up its all-gather channels on first use: not part of a trajectory
            D.all_gather_images(torch.zeros((B, 3, res, res), dtype=torch.uint8, device=dev), world)   # same size = same algorithm
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        xd = noise_host.to(dev, non_blocking=True)
        cd = cond_host.to(dev, non_blocking=True)
        ud = make_unc()
        _, u8 = sampler.sample_uint8(net, xd, cd, ud)
        if world > 1:
            u8 = D.all_gather_images(u8, world)
        if rank == 0:
            out_host.copy_(u8[: out_host.shape[0]], non_blocking=True)
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        h2d = noise_host.numel() * 4 + cond_host.numel() * cond_host.element_size()
        d2h = out_host.numel()
        e2e = dict(value=gbatch / (ms * 1e-3), unit=UNIT,
                   h2d_bytes_per_step=h2d / nsteps, d2h_bytes_per_step=d2h / nsteps,
                   seconds_per_trajectory=ms * 1e-3,
                   note=f"one sampler.sample_uint8 call = {nsteps} steps; bytes are per rank per trajectory / {nsteps}")

    def finish():
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        finish()
        return

    hbm = None
    if not args.no_hbm_kernels and world == 1 and not args.profile:
        del x
        torch.cuda.empty_cache()
        fn = baseline_head_rooflines if wl["kind"] == "baseline" else hbm_kernel_rooflines
        hbm = fn(torch, ops, net, dev, 2 * B, res, peaks["hbm_gbs"]) if wl["kind"] != "pixnerd" else None

    tgb = None
    if args.torch_baseline != "none" and world == 1 and not args.profile:
        torch.cuda.empty_cache()
        xb = noise_host.to(dev)
        tgb = torch_gpu_baseline(torch, args, wl, net, dev, xb, ts[nsteps // 3], cfg_cond, B, nsteps)
        del xb
        torch.cuda.empty_cache()

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sd = None
        if args.workload == "xl256":
            sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
        sec, cores = cpu_reference_step_seconds(sd, args.cpu_sample_steps, 1, args.cpu_batch)
        cpu = dict(value=args.cpu_batch / (sec * NUM_SAMPLING_STEPS), unit=UNIT, cores=cores, kind="port",
                   sample=f"XL/16 256px: {args.cpu_batch} images, {args.cpu_sample_steps} CFG-batched Euler steps (fp32 oracle "
                          f"port) timed after 1 warm-up, extrapolated linearly to 100 steps; {sec:.2f} s per step")

    line = dict(metric=wl["metric"], value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16",
                data="synthetic",
                config=dict(workload=wl["name"], global_batch=gbatch, per_gpu_batch=B, cfg_rows_per_gpu=2 * B,
                            num_sampling_steps=nsteps, step="one CFG-batched denoiser step + fused update",
                            launch="CUDA graph replay per step" if stepper is not None else "eager launches",
                            l2="inputs larger than L2 (>1.3 GB bf16 weights + >1 GB activations per step)",
                            parallelism=f"dp{world}"),
                ms_per_denoiser_step=ms_per_step,
                clocks=clk.report(), e2e=e2e, gpu_launches=launches,
                roofline=dict(kernel="gemm_bf16_tcgen05_kernel (all DiT / embed / cond_embed GEMMs)", bound="tensor",
                              achieved=gs["tflops"], peak=peaks["tf_sustained"], unit="TFLOP/s",
                              frac=(gs["tflops"] / peaks["tf_sustained"]) if peaks["tf_sustained"] else None,
                              traffic=traffic, traffic_source=traffic_tab, peak_source=peaks["source"] + ", sustained bf16",
                              launches=gs["launches"], avg_launch_ms=gs["avg_ms"],
                              avg_launch_algorithmic_gflop=gs["total_flops"] / max(1, gs["launches"]) / 1e9,
                              gemm_ms_per_step=gs["total_ms"] / args.steps,
                              gemm_share_of_step=(gs["total_ms"] / args.steps) / ms_per_step if ms_per_step else None,
                              measured="second pass of the same steps with CUDA events around every GEMM launch; "
                                       "share = GEMM ms per step of that pass / the timed ms_per_step",
                              per_gpu_step_tflops_algorithmic=(wl["gflop"] * 1e9 * 2 * B) / (ms_per_step * 1e-3) / 1e12,
                              step_frac_of_peak=(wl["gflop"] * 1e9 * 2 * B) / (ms_per_step * 1e-3) / 1e12 / peaks["tf_sustained"]),
                parity=parity, torch_gpu_baseline=tgb, hbm_kernels=hbm, cpu_baseline=cpu)
    print(json.dumps(line), flush=True)
    finish()


def run_train(args):
    """BASELINE configs[3]: one training step = REPATrainer(net, ...) forward (denoiser + DCT/FM loss) + loss.backward()
    on 32 synthetic images per GPU (weak scaling).  For N > 1 the parameter gradients are averaged with one NCCL
    all-reduce per step (what DDP does in the reference, src/lightning_model.py under strategy ddp)."""
    import torch
    import torch.distributed as dist
    from deco_b200 import LinearScheduler, PixNerDiT, REPATrainer, _lib, ops
    from deco_b200 import distributed as D
    from deco_b200.utils import GemmProbe, randomize_

    wl = WORKLOADS[args.workload]
    res = wl["res"]
    # overlapped gradient averaging (N > 1): the GEMMs leave DECO_B200_RESERVE_SMS SMs free and NCCL is kept inside them
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and int(os.environ.get("DECO_B200_RESERVE_SMS", "0")) > 0:
        os.environ.setdefault("NCCL_MAX_CTAS", os.environ["DECO_B200_RESERVE_SMS"])
    rank, world, local = D.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: deco_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()
    B = args.global_batch // world if args.global_batch else wl["batch"]
    with torch.device("meta"):
        net = PixNerDiT(**wl["model"])
    net = randomize_(net.to_empty(device=dev), seed=0).train()
    trainer = REPATrainer(scheduler=LinearScheduler(), null_condition_p=0.2, freq_loss_weight=1, timeshift=1.0).to(dev)
    params = [p for p in net.parameters() if p.requires_grad]
    gen = torch.Generator().manual_seed(1234 + rank)
    nbuf = 4
    x_host = [torch.tanh(torch.randn((B, 3, res, res), generator=gen)).pin_memory() for _ in range(nbuf)]
    y_host = [torch.randint(0, 1000, (B,), generator=gen).pin_memory() for _ in range(nbuf)]
    unc = torch.full((B,), 1000, dtype=torch.int64, device=dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    # N > 1: gradients are averaged like DDP does for the reference -- by default overlapped with the backward (per-block
    # asynchronous all-reduces started from inside the denoiser's backward), DECO_B200_OVERLAP=0 = one flat all-reduce after it
    overlap = world > 1 and os.environ.get("DECO_B200_OVERLAP", "1") != "0"

    def step(x, y, overlap=overlap):
        for p in params:
            p.grad = None
        d = trainer(net, None, None, x, y, unc)
        if overlap:
            with D.overlap_gradient_average(world):
                d["loss"].backward()
        else:
            d["loss"].backward()
            D.all_reduce_gradients(params, world)      # one bucket after the backward
        return d["loss"].detach()

    def train_parity():
        """parameter gradients of 2 images on the bench architecture against autograd over the fp32 oracle (checker only).
        north_star states no gradient tolerance; the bar of tests/test_gpu_backward.py (global rel-L2 <= 2e-2, worst tensor
        <= 6e-2: bf16 GEMM operands in forward AND backward) is applied and the per-tensor maxima are reported.  Runs after
        the timed regions, inside its own scope, so no autograd graph of it is alive while a step is captured or timed."""
        if args.no_parity or args.profile:
            return None
        from oracle import deco_oracle as O
        _, Pd = _oracle_forward(wl, net, dev)
        xs = x_host[0][:2].to(dev)
        ys = torch.tensor([17, 1000], device=dev)
        tt = torch.tensor([0.35, 0.8], device=dev)
        g2 = torch.Generator(device=dev).manual_seed(5)
        x_t, v_t = O.make_xt_vt(xs, torch.randn(xs.shape, device=dev, generator=g2), tt)
        for p in params:
            p.grad = None
        lo = trainer.loss(net(x_t, tt, ys), v_t)["loss"]
        lo.backward()
        Pr = {k: v.clone().requires_grad_(True) for k, v in Pd.items()}
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            lr_ = O.dct_fm_loss(O.denoiser_forward(Pr, O.CFG_XL, x_t, tt, ys), v_t)["loss"]
            lr_.backward()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        per = []
        num = den = 0.0
        for n_, p in net.named_parameters():
            a, b = p.grad.double(), Pr[n_].grad.double()
            dn, dd = float((a - b).pow(2).sum()), float(b.pow(2).sum())
            num, den = num + dn, den + dd
            per.append(((dn / max(dd, 1e-300)) ** 0.5, n_, p.numel()))
        per.sort(reverse=True)
        e = (num / den) ** 0.5
        big = [v for v in per if v[2] >= 4096]          # the per-tensor bar is applied to tensors with >= 4096 elements: a
        #                                                 3-element bias gradient is a sum over every pixel (cancellation noise)
        parity = dict(grad_rel_l2=e, worst_tensors=[dict(name=n_, rel_l2=v, numel=k) for v, n_, k in per[:3]],
                      worst_large_tensor=dict(name=big[0][1], rel_l2=big[0][0], numel=big[0][2]),
                      loss=float(lo.detach()), loss_oracle=float(lr_.detach()), tol=dict(global_rel_l2=2e-2, per_tensor=6e-2),
                      rows="2 images of the bench architecture (XL/16 256px), fixed t = (0.35, 0.8), labels (17, null)",
                      against="torch autograd over the fp32 oracle on the same GPU, same weights",
                      ok=bool(e <= 2e-2 and big[0][0] <= 6e-2))
        for p in params:
            p.grad = None
        if not parity["ok"]:
            raise SystemExit(f"bench: training parity check failed: {parity}")
        return parity

    overlap_check = None
    if overlap:       # same seeded step through both averaging paths: the gradients must agree (fp32 atomics aside)
        got = []
        for mode in (True, False):
            torch.manual_seed(777 + rank)
            step(x_host[0].to(dev), y_host[0].to(dev), overlap=mode)
            got.append([p.grad.detach().clone() for p in params])
        num = sum(float(((a.double() - b.double()) ** 2).sum()) for a, b in zip(*got))
        den = sum(float((b.double() ** 2).sum()) for b in got[1])
        overlap_check = (num / max(den, 1e-300)) ** 0.5
        del got
        if not overlap_check < 1e-4:
            raise SystemExit(f"overlapped gradient averaging disagrees with the flat all-reduce: rel-L2 {overlap_check:.3e}")
    torch.manual_seed(4321 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    xd = [t.to(dev) for t in x_host]
    yd = [t.to(dev) for t in y_host]
    # Static shapes: forward + backward is captured once into a CUDA graph and replayed per step (the reference compiles its
    # denoiser, src/lightning_model.py:94-97).  The RNG of the trainer (t, noise, label dropout) is graph-safe device RNG.
    # Multi-GPU steps stay eager (the NCCL all-reduce follows the backward).
    eager_step = step
    launch_mode = "eager launches"
    if world == 1 and os.environ.get("DECO_B200_GRAPH", "1") != "0" and not args.profile:
        try:
            sx, sy = xd[0].clone(), yd[0].clone()
            for _ in range(2):
                eager_step(sx, sy)                    # warm caches (prepared weights, function attributes)
            torch.cuda.synchronize()
            l0 = _lib.launch_count
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = eager_step(sx, sy)
            per_replay = _lib.launch_count - l0

            def step(x, y):                           # noqa: F811
                sx.copy_(x, non_blocking=True)
                sy.copy_(y, non_blocking=True)
                graph.replay()
                _lib.launch_count += per_replay
                return static_loss
            launch_mode = "CUDA graph replay of forward + backward"
        except Exception as e:   # noqa: BLE001
            print(f"bench: CUDA-graph capture of the training step failed ({e}); eager launches", file=sys.stderr)
            step = eager_step
    for i in range(args.warmup):
        step(xd[i % nbuf], yd[i % nbuf])
    barrier()
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        if args.profile:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(args.steps):
            step(xd[i % nbuf], yd[i % nbuf])
        e1.record()
        barrier()
        if args.profile:
            torch.cuda.profiler.stop()
    launches = _lib.launch_count - launches0
    ms_total = e0.elapsed_time(e1)
    probe = GemmProbe()            # roofline pass: same steps again with CUDA events around every GEMM launch
    if not args.profile:
        ops.gemm_probe = probe
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(args.steps):
            eager_step(xd[i % nbuf], yd[i % nbuf])
        p1.record()
        barrier()
        ops.gemm_probe = None
        ms_probe_total = p0.elapsed_time(p1)
    else:
        ms_probe_total = ms_total
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt)
    ms_per_step = ms_total / args.steps
    value = B * world / (ms_per_step * 1e-3)
    gs = probe.summary()
    peaks = measured_peaks()

    e2e = None
    if not args.no_e2e:
        k = max(2, min(args.steps, 5))
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(k):
            x = x_host[i % nbuf].to(dev, non_blocking=True)
            y = y_host[i % nbuf].to(dev, non_blocking=True)
            loss_host.copy_(step(x, y), non_blocking=True)
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1) / k
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        e2e = dict(value=B * world / (ms * 1e-3), unit=UNIT, h2d_bytes_per_step=B * 3 * res * res * 4 + B * 8,
                   d2h_bytes_per_step=4, ms_per_step=ms, loss=float(loss_host),
                   note="trainer(net, ...) + backward per step with pinned host images/labels copied in and the loss read back")
    # optimizer tail (not part of `value`: BASELINE configs[3] names forward + backward): the fused AdamW + EMA kernel
    # alone, and a full iteration = step + optimizer + re-preparation of the bf16 / packed weights the next forward needs
    from deco_b200 import FusedAdamWEMA
    ema = [p.detach().clone() for p in params]
    opt = FusedAdamWEMA(params, ema, lr=1e-4, weight_decay=0.0, ema_decay=0.9999)
    opt.step()
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for _ in range(5):
        opt.step()
    o1.record()
    barrier()
    opt_ms = o0.elapsed_time(o1) / 5
    nparam = sum(p.numel() for p in params)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(2):
        eager_step(xd[i % nbuf], yd[i % nbuf]); opt.step()
    barrier()
    f0.record()
    for i in range(3):
        eager_step(xd[i % nbuf], yd[i % nbuf])
        opt.step()
    f1.record()
    barrier()
    full_ms = f0.elapsed_time(f1) / 3
    optimizer = dict(kernel="adamw_ema_kernel (fused multi-tensor AdamW + EMA)", ms=opt_ms, parameters=nparam,
                     algorithmic_bytes=36.0 * nparam, achieved=36.0 * nparam / (opt_ms * 1e-3) / 1e9,
                     peak=measured_peaks()["hbm_gbs"], unit="GB/s",
                     frac=36.0 * nparam / (opt_ms * 1e-3) / 1e9 / measured_peaks()["hbm_gbs"],
                     iteration_with_optimizer_ms=full_ms,
                     note="iteration = forward + backward + optimizer + weight re-preparation (bf16 / packed copies)")
    parity = train_parity()         # last: it leaves no gradients behind and sees the weights the optimizer section produced
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    flops_step = wl["gflop"] * 1e9 * B
    line = dict(metric=wl["metric"], value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic",
                config=dict(workload=wl["name"], global_batch=B * world, per_gpu_batch=B,
                            step="trainer forward (denoiser + DCT/FM loss) + backward" + (
                                "" if world == 1 else " + gradient averaging: per-block asynchronous NCCL all-reduces overlapped "
                                "with the backward" if overlap else " + one flat NCCL all-reduce after the backward"),
                            overlap_check_rel_l2=overlap_check,
                            reserved_sms=int(os.environ.get("DECO_B200_RESERVE_SMS", "0")) if overlap else None,
                            nccl_max_ctas=os.environ.get("NCCL_MAX_CTAS"),
                            launch=launch_mode,
                            l2="inputs larger than L2 (>4 GB of weights + >10 GB of saved activations per step)",
                            parallelism=f"dp{world}"),
                clocks=clk.report(), e2e=e2e, gpu_launches=launches,
                roofline=dict(kernel="gemm_bf16_tcgen05_kernel (forward, dgrad and wgrad GEMMs of the DiT blocks)", bound="tensor",
                              achieved=gs["tflops"], peak=peaks["tf_sustained"], unit="TFLOP/s",
                              frac=(gs["tflops"] / peaks["tf_sustained"]) if peaks["tf_sustained"] else None,
                              traffic=None, peak_source=peaks["source"] + ", sustained bf16",
                              launches=gs["launches"], avg_launch_ms=gs["avg_ms"],
                              avg_launch_algorithmic_gflop=gs["total_flops"] / max(1, gs["launches"]) / 1e9,
                              gemm_ms_per_step=gs["total_ms"] / args.steps,
                              gemm_share_of_step=(gs["total_ms"] / args.steps) / ms_per_step if ms_per_step else None,
                              measured="second pass of the same steps with CUDA events around every GEMM launch; "
                                       "share = GEMM ms per step of that pass / the timed ms_per_step",
                              per_gpu_step_tflops_algorithmic=flops_step / (ms_per_step * 1e-3) / 1e12,
                              step_frac_of_peak=flops_step / (ms_per_step * 1e-3) / 1e12 / peaks["tf_sustained"]),
                full_iteration=dict(ms=full_ms, value=B * world / (full_ms * 1e-3), unit=UNIT,
                                    note="forward + backward + fused AdamW/EMA step + weight re-preparation (eager launches; the "
                                         "re-preparation is a CUDA-graph replay): what one optimisation step of a training run "
                                         "costs; `value` above is forward + backward only, as BASELINE configs[3] names it"),
                parity=parity, hbm_kernels=None, cpu_baseline=None, optimizer=optimizer)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif WORKLOADS[a.workload]["kind"] == "train":
        run_train(a)
    else:
        run_deco(a)
