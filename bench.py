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

NUM_SAMPLING_STEPS = 100
GLOBAL_BATCH = 256
RES = 256
GUIDANCE, G_MIN, G_MAX = 3.2, 0.1, 1.0
XL = dict(in_channels=3, num_groups=16, hidden_size=1152, hidden_size_x=32, num_blocks=31, num_cond_blocks=28,
          patch_size=16, num_classes=1000, nerf_mlpratio=2)
METRIC = "DeCo-XL/16 256px class-conditional sampling throughput (Euler 100 steps x CFG, fixed NFE)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="deco_b200", choices=["deco_b200", "reference"])
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-steps", type=int, default=2)
    ap.add_argument("--profile", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (use with ncu --profile-from-start off)")
    return ap.parse_args()


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
    batch = 1
    sec, cores = cpu_reference_step_seconds(None, max(1, args.steps), max(0, args.warmup), batch)
    value = batch / (sec * NUM_SAMPLING_STEPS)
    sample = (f"{batch} image, {args.steps} CFG-batched Euler steps of the 100 timed after {args.warmup} warm-up; "
              f"images/s extrapolated linearly to 100 steps")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="fp32", data="synthetic",
                config=dict(workload="DeCo-XL/16 256px c2i, Euler 100 x CFG 3.2 on (0.1,1]; CPU sample batch 1",
                            global_batch=batch, num_sampling_steps=NUM_SAMPLING_STEPS),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- deco_b200 arm
def run_deco(args):
    import torch
    import torch.distributed as dist
    from deco_b200 import EulerSampler, LinearScheduler, PixNerDiT, _lib, ode_step_fn, ops, simple_guidance_fn
    from deco_b200 import distributed as D
    from deco_b200.data import rank_indices, seeded_noise
    from deco_b200.utils import GemmProbe, randomize_

    rank, world, local = D.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: deco_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()
    assert world == args.gpus or world == 1, (world, args.gpus)

    idx = rank_indices(args.global_batch, rank, world)          # DistributedSampler(shuffle=False) shard
    B = len(idx)
    with torch.device("meta"):
        net = PixNerDiT(**XL)
    net = randomize_(net.to_empty(device=dev), seed=0).eval()
    net.prepare(dev)
    sch = LinearScheduler()
    sampler = EulerSampler(scheduler=sch, w_scheduler=sch, guidance_fn=simple_guidance_fn, num_steps=NUM_SAMPLING_STEPS,
                           guidance=GUIDANCE, guidance_interval_min=G_MIN, guidance_interval_max=G_MAX, step_fn=ode_step_fn)
    noise_host = seeded_noise(idx, (3, RES, RES))                # pinned host batch
    cond_host = torch.tensor(labels_for(idx), dtype=torch.int64).pin_memory()
    x = noise_host.to(dev, non_blocking=True)
    cond = cond_host.to(dev, non_blocking=True)
    unc = torch.full((B,), 1000, dtype=torch.int64, device=dev)
    cfg_cond = torch.cat([unc, cond])
    ts = sampler.timesteps

    def one_step(x, i):
        t_cur, t_next = ts[i % NUM_SAMPLING_STEPS], ts[i % NUM_SAMPLING_STEPS + 1]
        out = net(torch.cat([x, x]), torch.full((2 * B,), float(t_cur), device=dev), cfg_cond)
        g = GUIDANCE if (bool(t_cur > G_MIN) and bool(t_cur <= G_MAX)) else 1.0
        return ops.cfg_step(x, out, g, float(t_next - t_cur))[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        x = one_step(x, i)
    barrier()
    probe = GemmProbe()
    ops.gemm_probe = probe
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
    ops.gemm_probe = None
    launches = _lib.launch_count - launches0
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt)
    ms_per_step = ms_total / args.steps
    value = args.global_batch / (ms_per_step * 1e-3 * NUM_SAMPLING_STEPS)
    gs = probe.summary()
    peaks = measured_peaks()

    # ---- e2e: public sampler API, host buffers in, uint8 images out (+ all-gather)
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty((args.global_batch if world > 1 else B, 3, RES, RES), dtype=torch.uint8).pin_memory()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        xd = noise_host.to(dev, non_blocking=True)
        cd = cond_host.to(dev, non_blocking=True)
        ud = torch.full((B,), 1000, dtype=torch.int64, device=dev)
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
        h2d = noise_host.numel() * 4 + cond_host.numel() * 8
        d2h = out_host.numel()
        e2e = dict(value=args.global_batch / (ms * 1e-3), unit=UNIT,
                   h2d_bytes_per_step=h2d / NUM_SAMPLING_STEPS, d2h_bytes_per_step=d2h / NUM_SAMPLING_STEPS,
                   seconds_per_trajectory=ms * 1e-3,
                   note="one EulerSampler.sample_uint8 call = 100 steps; bytes are per rank per trajectory / 100")

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
        sec, cores = cpu_reference_step_seconds(sd, args.cpu_sample_steps, 1, 1)
        cpu = dict(value=1.0 / (sec * NUM_SAMPLING_STEPS), unit=UNIT, cores=cores, kind="port",
                   sample=f"1 image, {args.cpu_sample_steps} CFG-batched Euler steps (fp32 oracle port) timed after 1 warm-up, "
                          f"extrapolated linearly to 100 steps; {sec:.2f} s per step")

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="bf16",
                data="synthetic",
                config=dict(workload="DeCo-XL/16 256px c2i (configs_c2i/DeCo_XL.yaml), Euler 100 steps x CFG 3.2 on (0.1,1]",
                            global_batch=args.global_batch, per_gpu_batch=B, cfg_rows_per_gpu=2 * B,
                            num_sampling_steps=NUM_SAMPLING_STEPS, step="one CFG-batched denoiser step + fused update",
                            l2="inputs larger than L2 (1.36 GB bf16 weights + >1 GB activations per step)",
                            parallelism=f"dp{world}"),
                ms_per_denoiser_step=ms_per_step,
                clocks=clk.report(), e2e=e2e, gpu_launches=launches,
                roofline=dict(kernel="gemm_bf16_tcgen05_kernel (all DiT / embed / cond_embed GEMMs)", bound="tensor",
                              achieved=gs["tflops"], peak=peaks["tf_sustained"], unit="TFLOP/s",
                              frac=(gs["tflops"] / peaks["tf_sustained"]) if peaks["tf_sustained"] else None,
                              traffic=None, peak_source=peaks["source"] + ", sustained bf16",
                              launches=gs["launches"], avg_launch_ms=gs["avg_ms"],
                              gemm_share_of_step=gs["total_ms"] / ms_total if ms_total else None,
                              per_gpu_step_tflops_algorithmic=(244.9e9 * 2 * B) / (ms_per_step * 1e-3) / 1e12),
                cpu_baseline=cpu)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_deco(a)
