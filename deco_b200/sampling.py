"""CFG-batched flow-matching samplers with the reference's constructor / call interface, running the fused
sm_100a update kernel (csrc/sampler.cu) instead of ~8 eager kernels and a device->host sync per step.

Mirrors (reference paths):
  src/diffusion/base/sampling.py:8-37                BaseSampler (forward + trajectory return modes)
  src/diffusion/base/guidance.py:3-6                 simple_guidance_fn
  src/diffusion/flow_matching/sampling.py:11-15      shift_respace_fn, ode_step_fn
  src/diffusion/flow_matching/sampling.py:30-107     EulerSampler
  src/diffusion/flow_matching/sampling.py:17-24      sde_mean_step_fn, sde_step_fn, sde_preserve_step_fn (Euler only)
  src/diffusion/flow_matching/sampling.py:109-188    EulerSamplerJiT (x-prediction nets: configs_c2i/Baseline_DiT_JiT.yaml)
  src/diffusion/flow_matching/sampling.py:190-296    HeunSampler
  src/diffusion/flow_matching/adam_sampling.py:39-122 AdamLMSampler (+ src/diffusion/pre_integral.py:103-125)

The schedule is precomputed on the host in fp32 with the same torch ops as the reference (so e.g.
ts[10] = 0.09999999403953552 is *not* > 0.1) and the guidance-window test happens on the host
copy: no per-step synchronisation.  The Euler samplers also take the SDE step functions: the score coefficients are
host scalars from the scheduler, the Gaussian increment is `torch.randn_like(x)` (torch's CUDA generator, the very call the
reference makes, so a seeded run consumes the same Philox stream) and everything else is the same fused update kernel
(csrc/sampler.cu, extended form).  Heun and Adams stay ODE-only, as in every config of the reference.
"""
from __future__ import annotations

import logging
import os
from typing import Callable, List, Optional, Union

import torch
import torch.nn as nn

from . import _lib, ops
from .scheduling import BaseScheduler

logger = logging.getLogger(__name__)


def simple_guidance_fn(out, cfg):
    """Reference-compatible tensor form (rows [uncond || cond]); the samplers below fuse it into the step kernel."""
    uncondition, condition = out.chunk(2, dim=0)
    return uncondition + cfg * (condition - uncondition)


def shift_respace_fn(t, shift=3.0):
    return t / (t + (1 - t) * shift)


def ode_step_fn(x, v, dt, s, w):
    return x + v * dt


def sde_mean_step_fn(x, v, dt, s, w):
    return x + v * dt + s * w * dt


def sde_step_fn(x, v, dt, s, w):
    return x + v * dt + s * w * dt + torch.sqrt(2 * w * dt) * torch.randn_like(x)


def sde_preserve_step_fn(x, v, dt, s, w):
    return x + v * dt + 0.5 * s * w * dt + torch.sqrt(w * dt) * torch.randn_like(x)


_STEP_FNS = ("ode_step_fn", "sde_mean_step_fn", "sde_step_fn", "sde_preserve_step_fn")


def _step_kind(fn) -> str:
    name = getattr(fn, "__name__", "")
    if name not in _STEP_FNS:
        raise NotImplementedError(f"step function {name or fn!r} is not one of {_STEP_FNS}")
    return name


def sde_scalars(scheduler, w_scheduler, t_cur, dt, kind, t_w=None):
    """(kd, sden, a_s, a_n) of one step: the score s = (kd v - x) / sden (sampling.py:98) at t_cur and the coefficients of s
    and of the Gaussian increment in step function `kind` (:17-24) with w = w_scheduler.w(t_w or t_cur), evaluated with the
    reference's fp32 torch expressions on the scheduler's own methods."""
    tt = t_cur.reshape(1)
    sigma = scheduler.sigma(tt)
    kd = 1 / scheduler.dalpha_over_alpha(tt)
    sden = sigma ** 2 - kd * scheduler.dsigma_mul_sigma(tt)
    if kind == "ode_step_fn":
        return float(kd.reshape(-1)[0]), float(sden.reshape(-1)[0]), 0.0, 0.0
    tw = (t_cur if t_w is None else t_w).reshape(1)
    w = w_scheduler.w(tw) if w_scheduler else torch.zeros(1)
    w = torch.as_tensor(w, dtype=torch.float32).reshape(-1)[:1]
    if kind == "sde_mean_step_fn":
        a_s, a_n = w * dt, torch.zeros(1)
    elif kind == "sde_step_fn":
        a_s, a_n = w * dt, torch.sqrt(2 * w * dt)
    else:
        a_s, a_n = 0.5 * w * dt, torch.sqrt(w * dt)
    return float(kd.reshape(-1)[0]), float(sden.reshape(-1)[0]), float(a_s.reshape(-1)[0]), float(a_n.reshape(-1)[0])


def _make_timesteps(num_steps, last_step, timeshift):
    timesteps = torch.linspace(0.0, 1 - last_step, num_steps)
    timesteps = torch.cat([timesteps, torch.tensor([1.0])], dim=0)
    return shift_respace_fn(timesteps, timeshift)


class BaseSampler(nn.Module):
    def __init__(self, scheduler: BaseScheduler = None, guidance_fn: Callable = None, num_steps: int = 250,
                 guidance: Union[float, List[float]] = 1.0, *args, **kwargs):
        super().__init__()
        self.num_steps = num_steps
        self.guidance = guidance
        self.guidance_fn = guidance_fn
        self.scheduler = scheduler

    def _impl_sampling(self, net, noise, condition, uncondition, keep_x=False, keep_v=False, to_uint8=False):
        raise NotImplementedError

    def _graph_rows(self):
        """(rows, use_pred) for GraphedStepper, or None when this sampler's step cannot be replayed from a table."""
        return None

    def graphed_stepper(self, net, noise, cfg_condition, to_uint8=False):
        """GraphedStepper for (net, batch shape), built once and cached; None when the step cannot be graphed (a net that
        is not a deco_b200 denoiser, DECO_B200_GRAPH=0, an unsupported sampler, or a capture failure -- the eager loop is
        used then)."""
        if not GRAPH or not getattr(net, "cuda_graph_safe", False) or getattr(net, "training", False):
            return None
        spec = self._graph_rows()
        if spec is None:
            return None
        if not isinstance(self.guidance, (int, float)):
            return None               # per-sample guidance lists: eager loop (the graphed table holds one scalar per step)
        prep = net.prepare(noise.device) if hasattr(net, "prepare") else None
        # one stepper per (net, batch shape, output kind): a new weight version (id(prep) changes after an optimizer step
        # or a checkpoint load) REPLACES the entry, so the old graph, its memory pool and its bf16 weight copies are freed
        slot = (id(net), tuple(noise.shape), tuple(cfg_condition.shape), cfg_condition.dtype, bool(to_uint8),
                noise.device.index)
        version = (id(prep), float(self.guidance))
        cache = self.__dict__.setdefault("_steppers", {})
        hit = cache.get(slot)
        if hit is not None and hit[0] == version:
            return hit[1]
        cache.pop(slot, None)
        while len(cache) >= MAX_STEPPERS:
            cache.pop(next(iter(cache)))
        try:
            st = GraphedStepper(self, net, noise.shape[0], noise.shape[1:], cfg_condition[: noise.shape[0]],
                                to_uint8, spec[0], spec[1])
        except Exception as e:   # noqa: BLE001 -- capture is an optimisation; the eager loop runs the same kernels
            logger.warning("CUDA-graph capture of the sampling step failed (%s); using the eager loop", e)
            st = None
        cache[slot] = (version, st)
        return st

    def _run_graphed(self, net, x, cfg_condition, to_uint8):
        st = self.graphed_stepper(net, x, cfg_condition, to_uint8)
        if st is None:
            return None
        st.reset(x, cfg_condition)
        for _ in range(self.num_steps):
            st.step()
        return st.x.clone(), None, None, (st.u8.clone() if to_uint8 else None)

    def _check(self):
        if self.guidance_fn is not None and self.guidance_fn is not simple_guidance_fn \
                and getattr(self.guidance_fn, "__name__", "") != "simple_guidance_fn":
            raise NotImplementedError("only simple_guidance_fn is fused into the sampler kernel")

    @torch.no_grad()
    def forward(self, net, noise, condition, uncondition, return_x_trajs=False, return_v_trajs=False):
        """Same contract as src/diffusion/base/sampling.py:28-37.  Trajectories are only materialised on request
        (the reference always keeps every step: 30 GB at batch 256 x 100 steps)."""
        self._check()
        x, x_trajs, v_trajs, _ = self._impl_sampling(net, noise, condition, uncondition,
                                                     keep_x=return_x_trajs, keep_v=return_v_trajs)
        if return_x_trajs and return_v_trajs:
            return x, x_trajs, v_trajs
        elif return_x_trajs:
            return x, x_trajs
        elif return_v_trajs:
            return x, v_trajs
        return x

    @torch.no_grad()
    def sample_uint8(self, net, noise, condition, uncondition):
        """Sampling with fp2uint8 (autoencoder/base.py:32-34; PixelAE.decode is the identity for scale 1, shift 0)
        fused into the last update: returns (x_final fp32, images uint8)."""
        self._check()
        x, _, _, u8 = self._impl_sampling(net, noise, condition, uncondition, to_uint8=True)
        return x, u8


def _prep_inputs(noise, condition, uncondition):
    if not noise.is_cuda:
        raise RuntimeError("deco_b200 samplers run on CUDA tensors only (no CPU fallback)")
    x = noise.detach().to(torch.float32).contiguous()
    cfg_condition = torch.cat([uncondition, condition], dim=0)
    return x, cfg_condition


GRAPH = os.environ.get("DECO_B200_GRAPH", "1") != "0"
MAX_STEPPERS = 4     # captured graphs kept per sampler (distinct batch shapes / output kinds)


class GraphedStepper:
    """One CFG sampling step -- schedule advance, denoiser forward, fused update -- captured ONCE into a CUDA graph and replayed
    per step.  The step scalars (t, g, dt) live in a device table indexed by a device counter (csrc/sampler.cu
    sampler_advance_kernel), so a whole trajectory is `num_steps` replays with no host work in between: at small per-GPU
    batches (8-GPU sharding: 64 CFG rows) the ~210 Python/ctypes launches of a step cost as much host time as the step
    takes on the GPU.  Same kernels, same order, same numerics as the eager loops (Euler; Adams order <= 2, whose previous
    prediction lives in one buffer that the update kernel reads and overwrites in place).
    rows: per step {g, dt, c0, c1, 0, 0, t, 0} (the sampler's `_graph_rows`)."""

    def __init__(self, sampler, net, batch, shape, cond_like, to_uint8, rows, use_pred=False):
        dev = cond_like.device
        self.sampler, self.net, self.B = sampler, net, batch
        self.table = torch.tensor(rows, dtype=torch.float32).to(dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cur = torch.zeros(8, dtype=torch.float32, device=dev)
        self.t = torch.zeros(2 * batch, dtype=torch.float32, device=dev)
        self.x = torch.zeros((batch,) + tuple(shape), dtype=torch.float32, device=dev)
        self.cond = torch.zeros((2 * batch,) + tuple(cond_like.shape[1:]), dtype=cond_like.dtype, device=dev)
        self.u8 = torch.zeros(self.x.shape, dtype=torch.uint8, device=dev) if to_uint8 else None
        self.pred = torch.zeros_like(self.x) if use_pred else None
        self.xpred = bool(getattr(sampler, "x_prediction", False))
        cur_stream = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur_stream)
        with torch.cuda.stream(side):       # eager warm-up: lazily-set function attributes, weight / table caches
            self._body()
        cur_stream.wait_stream(side)
        torch.cuda.synchronize(dev)
        n0 = _lib.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = _lib.launch_count - n0
        self.prep_ref = getattr(net, "_prep", None)     # the graph reads these weight buffers: keep them alive

    def _body(self):
        ops.sampler_advance(self.table, self.counter, self.cur, self.t)
        if not self.xpred and _fused_step_ok(self.net):
            # decoder epilogue = guidance + update (csrc/decoder_tc.cu): no cat[x, x], no bf16 network output, no step kernel
            self.net.cfg_step(self.x, self.t, self.cond, dev=self.cur, p1=self.pred, x_out=self.x, pred_out=self.pred,
                              u8_out=self.u8)
            return
        out = self.net(torch.cat([self.x, self.x], dim=0), self.t, self.cond)
        if out.dtype not in (torch.bfloat16, torch.float32):
            out = out.float()
        if self.xpred:      # EulerSamplerJiT: column 7 of the schedule row is the x-prediction denominator
            ops.cfg_step_ex(self.x, out.contiguous(), dev=self.cur, x_out=self.x, u8_out=self.u8)
        else:
            ops.cfg_step_dev(self.x, out.contiguous(), self.cur, self.x, p1=self.pred, pred_out=self.pred, u8_out=self.u8)

    def reset(self, x, cfg_condition):
        self.x.copy_(x)
        self.cond.copy_(cfg_condition)
        self.counter.zero_()
        if self.pred is not None:
            self.pred.zero_()

    def step(self):
        self.graph.replay()
        _lib.launch_count += self.launches_per_step


class HeunGraphedStepper:
    """HeunSampler (ODE steps) on a denoiser with the fused decoder-epilogue update, as CUDA-graph replays: the schedule
    scalars live in a device table walked by csrc/sampler.cu::sampler_advance_kernel (one row per network evaluation or
    element-wise predictor), so a trajectory is `num_steps` replays with no host work in between.  Three captured bodies:
      full  (step 0, and every step when exact_henu)  predictor evaluation at t_cur (x_hat = x + dt pred, pred kept) +
                                                       corrector evaluation at t_next (x += dt (pred + pred_hat) / 2)
      mid   (re-use variant, steps 1 .. n-2)           predictor x_hat = x + dt v_hat (element-wise) + corrector evaluation
      last                                             predictor only: x = x_hat (+ fp2uint8)
    Same kernels, same order, same numerics as HeunSampler's eager loop."""

    def __init__(self, sampler, net, batch, shape, cond_like, to_uint8):
        dev = cond_like.device
        self.sampler, self.net, self.B, self.n = sampler, net, batch, sampler.num_steps
        ts = sampler.timesteps
        rows, self.kinds = [], []
        for i in range(self.n):
            t_cur, t_next = ts[i], ts[i + 1]
            dt = float(t_next - t_cur)
            in_window = bool(t_cur > sampler.guidance_interval_min) and bool(t_cur <= sampler.guidance_interval_max)
            g = float(sampler.guidance) if in_window else 1.0
            last = i == self.n - 1
            evaluates = i == 0 or sampler.exact_henu
            if evaluates:
                rows.append([g, dt, 1.0, 0.0, 0.0, 0.0, float(t_cur), 0.0])           # predictor from the network
            else:
                rows.append([1.0, dt, 0.0, 1.0, 0.0, 0.0, float(t_next), 0.0])        # predictor from the kept velocity
            if not last:
                rows.append([g, dt, 0.5, 0.5, 0.0, 0.0, float(t_next), 0.0])          # corrector
            self.kinds.append(("full" if evaluates else "mid") if not last else ("last_eval" if evaluates else "last"))
        self.table = torch.tensor(rows, dtype=torch.float32).to(dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cur = torch.zeros(8, dtype=torch.float32, device=dev)
        self.t = torch.zeros(2 * batch, dtype=torch.float32, device=dev)
        self.x = torch.zeros((batch,) + tuple(shape), dtype=torch.float32, device=dev)
        self.x_hat = torch.zeros_like(self.x)
        self.v = torch.zeros_like(self.x)            # predictor velocity of the current step
        self.v_hat = torch.zeros_like(self.x)        # corrector velocity, kept for the next predictor
        self.cond = torch.zeros((2 * batch,) + tuple(cond_like.shape[1:]), dtype=cond_like.dtype, device=dev)
        self.u8 = torch.zeros(self.x.shape, dtype=torch.uint8, device=dev) if to_uint8 else None
        self.zero_net = _zero_net_like(self.x)
        self.graphs, self.launches = {}, {}
        cur_stream = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        for kind in sorted(set(self.kinds)):
            side.wait_stream(cur_stream)
            with torch.cuda.stream(side):       # eager warm-up: lazily-set function attributes, weight / table caches
                self._body(kind)
            cur_stream.wait_stream(side)
            torch.cuda.synchronize(dev)
            n0 = _lib.launch_count
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                self._body(kind)
            self.graphs[kind], self.launches[kind] = g_, _lib.launch_count - n0
        self.prep_ref = getattr(net, "_prep", None)

    def _advance(self):
        ops.sampler_advance(self.table, self.counter, self.cur, self.t)

    def _body(self, kind):
        if kind in ("full", "last_eval"):
            self._advance()
            self.net.cfg_step(self.x, self.t, self.cond, dev=self.cur, x_out=(self.x if kind == "last_eval" else self.x_hat),
                              pred_out=self.v, u8_out=(self.u8 if kind == "last_eval" else None))
            if kind == "last_eval":
                return
            p1 = self.v
        elif kind in ("mid", "last"):
            self._advance()                     # predictor from the velocity the last corrector left: x_hat = x + dt v_hat
            ops.cfg_step_dev(self.x, self.zero_net, self.cur, self.x if kind == "last" else self.x_hat, p1=self.v_hat,
                             u8_out=(self.u8 if kind == "last" else None))
            if kind == "last":
                return
            p1 = self.v_hat                     # read as p1 and overwritten (pred_out) element by element
        self._advance()
        self.net.cfg_step(self.x_hat, self.t, self.cond, dev=self.cur, p1=p1, pred_out=self.v_hat, x_out=self.x, x_base=self.x)

    def reset(self, x, cfg_condition):
        self.x.copy_(x)
        self.cond.copy_(cfg_condition)
        self.counter.zero_()

    def run(self):
        for kind in self.kinds:
            self.graphs[kind].replay()
            _lib.launch_count += self.launches[kind]


FUSED_STEP = os.environ.get("DECO_B200_FUSED_STEP", "1") != "0"


def _fused_step_ok(net) -> bool:
    """The denoiser can run guidance + update inside its decoder epilogue (PixNerDiT.cfg_step)."""
    if not FUSED_STEP or not getattr(net, "supports_fused_cfg_step", False) or getattr(net, "training", False):
        return False
    from . import denoiser
    return denoiser.DECODER == "tc"


def _net_eval(net, x, t_scalar: float, cfg_condition, batch_size):
    cfg_x = torch.cat([x, x], dim=0)
    cfg_t = torch.full((2 * batch_size,), t_scalar, dtype=torch.float32, device=x.device)
    out = net(cfg_x, cfg_t, cfg_condition)
    if out.dtype not in (torch.bfloat16, torch.float32):
        out = out.float()
    return out.contiguous()


class EulerSampler(BaseSampler):
    x_prediction = False    # EulerSamplerJiT: the net predicts x instead of v

    def __init__(self, w_scheduler: BaseScheduler = None, timeshift=1.0, guidance_interval_min: float = 0.0,
                 guidance_interval_max: float = 1.0, step_fn: Callable = ode_step_fn, last_step=None,
                 last_step_fn: Callable = ode_step_fn, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.step_fn = step_fn
        self.last_step = last_step
        self.last_step_fn = last_step_fn
        self.w_scheduler = w_scheduler
        self.timeshift = timeshift
        self.guidance_interval_min = guidance_interval_min
        self.guidance_interval_max = guidance_interval_max
        if self.last_step is None or self.num_steps == 1:
            self.last_step = 1.0 / self.num_steps
        self.timesteps = _make_timesteps(self.num_steps, self.last_step, self.timeshift)
        assert self.last_step > 0.0
        assert self.scheduler is not None
        self._kinds = (_step_kind(self.step_fn), _step_kind(self.last_step_fn))
        assert self.w_scheduler is not None or self._kinds[0] == "ode_step_fn"      # sampling.py:61
        if self.w_scheduler is not None and self._kinds[0] == "ode_step_fn":
            logger.warning("current sampler is ODE sampler, but w_scheduler is enabled")

    def _xpred_den(self, t_cur) -> float:
        """(1 - t).clamp_min(5e-2) in fp32 (sampling.py:170); 0 = the net already predicts v."""
        return float((1.0 - t_cur).clamp_min(5e-2)) if self.x_prediction else 0.0

    def _sde_scalars(self, t_cur, dt, kind):
        return sde_scalars(self.scheduler, self.w_scheduler, t_cur, dt, kind)

    def _graph_rows(self):
        if self._kinds != ("ode_step_fn", "ode_step_fn"):
            return None               # SDE steps draw fresh noise per step: eager loop
        rows, ts = [], self.timesteps
        for i in range(self.num_steps):
            t_cur, t_next = ts[i], ts[i + 1]
            in_window = bool(t_cur > self.guidance_interval_min) and bool(t_cur <= self.guidance_interval_max)
            rows.append([float(self.guidance) if in_window else 1.0, float(t_next - t_cur), 1.0, 0.0, 0.0, 0.0, float(t_cur),
                         self._xpred_den(t_cur)])
        return rows, False

    def _impl_sampling(self, net, noise, condition, uncondition, keep_x=False, keep_v=False, to_uint8=False):
        B = noise.shape[0]
        x, cfg_condition = _prep_inputs(noise, condition, uncondition)
        steps = self.timesteps  # host fp32
        if not keep_x and not keep_v:
            res = self._run_graphed(net, x, cfg_condition, to_uint8)
            if res is not None:
                return res
        x_trajs, v_trajs, u8 = ([x] if keep_x else None), ([] if keep_v else None), None
        plain = not self.x_prediction and self._kinds == ("ode_step_fn", "ode_step_fn")
        fused = plain and not keep_v and _fused_step_ok(net)      # guidance + update inside the decoder epilogue
        for i in range(self.num_steps):
            t_cur, t_next = steps[i], steps[i + 1]
            dt = float(t_next - t_cur)
            in_window = bool(t_cur > self.guidance_interval_min) and bool(t_cur <= self.guidance_interval_max)
            g = float(self.guidance) if in_window else 1.0
            last = i == self.num_steps - 1
            if fused:
                cfg_t = torch.full((2 * B,), float(t_cur), dtype=torch.float32, device=x.device)
                u = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if (to_uint8 and last) else None
                x = net.cfg_step(x, cfg_t, cfg_condition, g=g, dt=dt, u8_out=u)
                if keep_x:
                    x_trajs.append(x)
                if u is not None:
                    u8 = u
                continue
            out = _net_eval(net, x, float(t_cur), cfg_condition, B)
            if plain:
                x, _, v, u = ops.cfg_step(x, out, g, dt, want_v=keep_v, want_u8=(to_uint8 and last))
            else:
                kd, sden, a_s, a_n = self._sde_scalars(t_cur, t_next - t_cur, self._kinds[1 if last else 0])
                z = torch.randn_like(x) if a_n != 0.0 else None
                x, _, v, u = ops.cfg_step_ex(x, out, g, dt, xpred_den=self._xpred_den(t_cur), kd=kd, sden=sden, a_s=a_s,
                                             a_n=a_n, noise=z, want_v=keep_v, want_u8=(to_uint8 and last))
            if keep_x:
                x_trajs.append(x)
            if keep_v:
                v_trajs.append(v)
            if u is not None:
                u8 = u
        if keep_v:
            v_trajs.append(torch.zeros_like(x))
        return x, x_trajs, v_trajs, u8


class EulerSamplerJiT(EulerSampler):
    """src/diffusion/flow_matching/sampling.py:109-188: the Euler loop for a net that predicts the clean image; the
    velocity is (out - x) / clamp_min(1 - t, 0.05) per CFG half (:170), folded into the update kernel."""
    x_prediction = True


class HeunSampler(BaseSampler):
    def __init__(self, scheduler: BaseScheduler = None, w_scheduler: BaseScheduler = None, exact_henu=False,
                 guidance_interval_min: float = 0.0, guidance_interval_max: float = 1.0, timeshift=1.0,
                 step_fn: Callable = ode_step_fn, last_step=None, last_step_fn: Callable = ode_step_fn,
                 *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.scheduler = scheduler
        self.exact_henu = exact_henu
        self.step_fn = step_fn
        self.last_step = last_step
        self.last_step_fn = last_step_fn
        self.w_scheduler = w_scheduler
        self.timeshift = timeshift
        self.guidance_interval_min = guidance_interval_min
        self.guidance_interval_max = guidance_interval_max
        if self.last_step is None or self.num_steps == 1:
            self.last_step = 1.0 / self.num_steps
        self.timesteps = _make_timesteps(self.num_steps, self.last_step, self.timeshift)
        assert self.last_step > 0.0
        assert self.scheduler is not None
        self._kinds = (_step_kind(self.step_fn), _step_kind(self.last_step_fn))
        assert self.w_scheduler is not None or self._kinds[0] == "ode_step_fn"      # sampling.py:225

    def _impl_sampling(self, net, noise, condition, uncondition, keep_x=False, keep_v=False, to_uint8=False):
        B = noise.shape[0]
        x, cfg_condition = _prep_inputs(noise, condition, uncondition)
        steps = self.timesteps
        x_trajs, v_trajs, u8 = ([x] if keep_x else None), ([] if keep_v else None), None
        if self._kinds != ("ode_step_fn", "ode_step_fn"):
            return self._impl_sampling_sde(net, x, cfg_condition, B, keep_x, keep_v, to_uint8)
        v_hat = None  # fp32 guided velocity at (x_hat, t_next) of the previous step
        fused = not keep_v and _fused_step_ok(net)     # guidance + predictor / corrector update inside the decoder epilogue
        if fused and not keep_x and GRAPH and getattr(net, "cuda_graph_safe", False) and isinstance(self.guidance, (int, float)):
            st = self._heun_stepper(net, x, cfg_condition, to_uint8)
            if st is not None:
                st.reset(x, cfg_condition)
                st.run()
                return st.x.clone(), None, None, (st.u8.clone() if to_uint8 else None)

        def full(t_scalar):
            return torch.full((2 * B,), float(t_scalar), dtype=torch.float32, device=x.device)
        for i in range(self.num_steps):
            t_cur, t_next = steps[i], steps[i + 1]
            dt = float(t_next - t_cur)
            in_window = bool(t_cur > self.guidance_interval_min) and bool(t_cur <= self.guidance_interval_max)
            g = float(self.guidance) if in_window else 1.0
            last = i == self.num_steps - 1
            u = None
            if i == 0 or self.exact_henu:
                if fused:       # predictor: x_hat = x + dt pred, pred kept in fp32 for the corrector
                    v = torch.empty_like(x)
                    u = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if (to_uint8 and last) else None
                    x_hat = net.cfg_step(x, full(t_cur), cfg_condition, g=g, dt=dt, pred_out=v, u8_out=u)
                else:
                    out = _net_eval(net, x, float(t_cur), cfg_condition, B)
                    x_hat, v, _, u = ops.cfg_step(x, out, g, dt, want_pred=True, want_u8=(to_uint8 and last))
            else:
                # predictor re-uses the corrector's velocity: x_hat = x + dt * v_hat  (c0 = 0 drops the net term)
                v = v_hat
                x_hat, _, _, u = ops.cfg_step(x, _zero_net_like(x), 1.0, dt, c0=0.0, prev=(v,), coeffs=(1.0,),
                                              want_u8=(to_uint8 and last))
            if not last:
                if fused:       # corrector: the net sees x_hat, x = x + dt (v + v_hat) / 2; v_hat kept for the next predictor
                    v_hat = torch.empty_like(x)
                    x = net.cfg_step(x_hat, full(t_next), cfg_condition, g=g, dt=dt, c0=0.5, c1=0.5, p1=v, pred_out=v_hat,
                                     x_base=x)
                else:
                    out = _net_eval(net, x_hat, float(t_next), cfg_condition, B)
                    # x = x + dt * (v + v_hat) / 2 ; v_hat is kept (fp32) for the next predictor
                    x, v_hat, v_avg, _ = ops.cfg_step(x, out, g, dt, c0=0.5, prev=(v,), coeffs=(0.5,), want_pred=True,
                                                      want_v=keep_v)
                    v = v_avg if keep_v else v
            else:
                x = x_hat
                u8 = u
            if keep_x:
                x_trajs.append(x)
            if keep_v:
                v_trajs.append(v)
        if keep_v:
            v_trajs.append(torch.zeros_like(x))
        return x, x_trajs, v_trajs, u8


    def _heun_stepper(self, net, x, cfg_condition, to_uint8):
        """HeunGraphedStepper for (net, batch shape), cached like BaseSampler.graphed_stepper; None on a capture failure."""
        prep = net.prepare(x.device) if hasattr(net, "prepare") else None
        slot = (id(net), tuple(x.shape), tuple(cfg_condition.shape), cfg_condition.dtype, bool(to_uint8), x.device.index)
        version = (id(prep), float(self.guidance))
        cache = self.__dict__.setdefault("_steppers", {})
        hit = cache.get(slot)
        if hit is not None and hit[0] == version:
            return hit[1]
        cache.pop(slot, None)
        while len(cache) >= MAX_STEPPERS:
            cache.pop(next(iter(cache)))
        try:
            st = HeunGraphedStepper(self, net, x.shape[0], x.shape[1:], cfg_condition[: x.shape[0]], to_uint8)
        except Exception as e:   # noqa: BLE001 -- capture is an optimisation; the eager loop runs the same kernels
            logger.warning("CUDA-graph capture of the Heun step failed (%s); using the eager loop", e)
            st = None
        cache[slot] = (version, st)
        return st

    def _impl_sampling_sde(self, net, x, cfg_condition, B, keep_x, keep_v, to_uint8):
        """Heun with an SDE step function (sampling.py:266-293): velocities AND scores of the two evaluations are averaged
        (csrc/sampler.cu heun_sde_step_kernel); the Gaussian increments are torch.randn_like draws in the reference's order
        (predictor, then corrector / last step), so a seeded run consumes the same Philox stream."""
        steps = self.timesteps
        x_trajs, v_trajs, u8 = ([x] if keep_x else None), ([] if keep_v else None), None
        v_hat = s_hat = None
        for i in range(self.num_steps):
            t_cur, t_next = steps[i], steps[i + 1]
            dtt = t_next - t_cur
            dt = float(dtt)
            in_window = bool(t_cur > self.guidance_interval_min) and bool(t_cur <= self.guidance_interval_max)
            g = float(self.guidance) if in_window else 1.0
            last = i == self.num_steps - 1
            kd, sden, a_s, a_n = sde_scalars(self.scheduler, self.w_scheduler, t_cur, dtt, self._kinds[0])
            kdh, sdenh, _, _ = sde_scalars(self.scheduler, self.w_scheduler, t_next, dtt, "ode_step_fn")
            if i == 0 or self.exact_henu:
                out = _net_eval(net, x, float(t_cur), cfg_condition, B)
                _, v, _, _ = ops.cfg_step(x, out, g, 0.0, want_pred=True)       # guided velocity in fp32
                s_in = None
            else:
                v, s_in = v_hat, s_hat
            z = torch.randn_like(x) if a_n != 0.0 else None
            x_hat, _, _, _, _ = ops.heun_sde_step(x, v, dt, a_s, a_n, kd, sden, s_in=s_in, noise=z)
            if not last:
                out = _net_eval(net, x_hat, float(t_next), cfg_condition, B)
                z = torch.randn_like(x) if a_n != 0.0 else None
                x, v_hat, s_hat, v_avg, _ = ops.heun_sde_step(x, v, dt, a_s, a_n, kd, sden, s_in=s_in, noise=z, net_out=out,
                                                              x_hat=x_hat, g=g, kdh=kdh, sdenh=sdenh, want_v_avg=keep_v)
                v = v_avg if keep_v else v
            else:
                _, _, a_sl, a_nl = sde_scalars(self.scheduler, self.w_scheduler, t_cur, dtt, self._kinds[1])
                z = torch.randn_like(x) if a_nl != 0.0 else None
                x, _, _, _, u8 = ops.heun_sde_step(x, v, dt, a_sl, a_nl, kd, sden, s_in=s_in, noise=z, want_u8=to_uint8)
            if keep_x:
                x_trajs.append(x)
            if keep_v:
                v_trajs.append(v)
        if keep_v:
            v_trajs.append(torch.zeros_like(x))
        return x, x_trajs, v_trajs, u8


_ZERO_NET = {}


def _zero_net_like(x):
    """A cached all-zero [2B, ...] bf16 tensor: the `net_out` argument of an update that drops the network term (c0 = 0)."""
    key = (tuple(x.shape), x.device)
    z = _ZERO_NET.get(key)
    if z is None:
        if len(_ZERO_NET) > 4:
            _ZERO_NET.clear()
        z = _ZERO_NET[key] = torch.zeros((2 * x.shape[0],) + tuple(x.shape[1:]), dtype=torch.bfloat16, device=x.device)
    return z


# ------------------------------------------------------------------ Adams linear multistep
def _lagrange_coeffs(order, ts, t0, t1):
    """Normalised integrals over [t0, t1] of the Lagrange basis polynomials on the last `order` nodes of ts
    (src/diffusion/pre_integral.py:4-125, orders 1-4).  Evaluated in fp32 torch arithmetic with the reference's
    expression order for orders 1-2 (used by every config) and in float64 polynomials for orders 3-4."""
    order = min(order, len(ts))
    if order == 1:
        return (1.0,)
    if order == 2:
        ta, tb = ts[-2], ts[-1]
        int1 = 0.5 / (ta - tb) * ((t1 - tb) ** 2 - (t0 - tb) ** 2)
        int2 = 0.5 / (tb - ta) * ((t1 - ta) ** 2 - (t0 - ta) ** 2)
        tot = int1 + int2
        return (float(int1 / tot), float(int2 / tot))
    import numpy as np
    nodes = [float(v) for v in ts[-order:]]
    a, b = float(t0), float(t1)
    ints = []
    for j, tj in enumerate(nodes):
        poly, den = np.poly1d([1.0]), 1.0
        for m, tm in enumerate(nodes):
            if m != j:
                poly = poly * np.poly1d([1.0, -tm])
                den *= (tj - tm)
        ip = poly.integ()
        ints.append((ip(b) - ip(a)) / den)
    tot = sum(ints)
    return tuple(v / tot for v in ints)


def nop(t):
    return t


class AdamLMSampler(BaseSampler):
    def __init__(self, order: int = 2, timeshift: float = 1.0, guidance_interval_min: float = 0.0,
                 guidance_interval_max: float = 1.0, lms_transform_fn: Callable = nop, last_step=None,
                 step_fn: Callable = ode_step_fn, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.step_fn = step_fn
        assert self.scheduler is not None
        assert getattr(self.step_fn, "__name__", "") == "ode_step_fn"
        assert 1 <= order <= 4, "Invalid order"
        self.order = order
        self.lms_transform_fn = lms_transform_fn
        self.last_step = last_step
        self.guidance_interval_min = guidance_interval_min
        self.guidance_interval_max = guidance_interval_max
        if self.last_step is None:
            self.last_step = 1.0 / self.num_steps
        self.timesteps = _make_timesteps(self.num_steps, self.last_step, timeshift)
        self.timedeltas = self.timesteps[1:] - self.timesteps[:-1]
        self._reparameterize_coeffs()

    def _reparameterize_coeffs(self):
        coeffs = []
        for i in range(self.num_steps):
            pre_ts = self.lms_transform_fn(self.timesteps[: i + 1])
            t0 = self.lms_transform_fn(self.timesteps[i])
            t1 = self.lms_transform_fn(self.timesteps[i + 1])
            coeffs.append(_lagrange_coeffs(min(self.order, i + 1), pre_ts, t0, t1))
        self.solver_coeffs = coeffs

    def _graph_rows(self):
        if self.order > 2:
            return None               # one in-place prediction buffer = order <= 2 (every DeCo config uses order 2)
        rows = []
        t_cur = torch.zeros((), dtype=torch.float32)
        for i in range(self.num_steps):
            in_window = bool(t_cur > self.guidance_interval_min) and bool(t_cur < self.guidance_interval_max)
            cs = self.solver_coeffs[i]
            rows.append([float(self.guidance) if in_window else 1.0, float(self.timedeltas[i]), float(cs[-1]),
                         float(cs[0]) if len(cs) > 1 else 0.0, 0.0, 0.0, float(t_cur), 0.0])
            t_cur = t_cur + self.timedeltas[i]
        return rows, self.order > 1

    def _impl_sampling(self, net, noise, condition, uncondition, keep_x=False, keep_v=False, to_uint8=False):
        B = noise.shape[0]
        x, cfg_condition = _prep_inputs(noise, condition, uncondition)
        if not keep_x and not keep_v:
            res = self._run_graphed(net, x, cfg_condition, to_uint8)
            if res is not None:
                return res
        x_trajs, v_trajs, u8 = ([x] if keep_x else None), ([] if keep_v else None), None
        preds: List[torch.Tensor] = []
        # the reference accumulates t_cur += dt in fp32 on the device (adam_sampling.py:96,118); same sums here
        t_cur = torch.zeros((), dtype=torch.float32)
        fused = self.order <= 2 and not keep_v and _fused_step_ok(net)
        for i in range(self.num_steps):
            in_window = bool(t_cur > self.guidance_interval_min) and bool(t_cur < self.guidance_interval_max)
            g = float(self.guidance) if in_window else 1.0
            cs = self.solver_coeffs[i]
            order = len(cs)
            prev = preds[-(order - 1):] if order > 1 else []
            last = i == self.num_steps - 1
            dt = float(self.timedeltas[i])
            if fused:       # decoder epilogue = guidance + multistep update (previous prediction read, new one written)
                cfg_t = torch.full((2 * B,), float(t_cur), dtype=torch.float32, device=x.device)
                u = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if (to_uint8 and last) else None
                pred = torch.empty_like(x) if self.order > 1 else None
                x = net.cfg_step(x, cfg_t, cfg_condition, g=g, dt=dt, c0=float(cs[-1]), c1=float(cs[0]) if order > 1 else 0.0,
                                 p1=prev[0] if order > 1 else None, pred_out=pred, u8_out=u)
                if self.order > 1:
                    preds = [pred]
                t_cur = t_cur + self.timedeltas[i]
                if keep_x:
                    x_trajs.append(x)
                if u is not None:
                    u8 = u
                continue
            out = _net_eval(net, x, float(t_cur), cfg_condition, B)
            # v = sum_j cs[j] * pred[-order:][j]; the newest prediction (cs[-1]) comes straight from the net output
            x, pred, v, u = ops.cfg_step(x, out, g, dt, c0=cs[-1], prev=tuple(prev), coeffs=tuple(cs[:-1]),
                                         want_pred=(self.order > 1), want_v=keep_v, want_u8=(to_uint8 and last))
            if self.order > 1:
                preds.append(pred)
                preds = preds[-(self.order - 1):]
            t_cur = t_cur + self.timedeltas[i]
            if keep_x:
                x_trajs.append(x)
            if keep_v:
                v_trajs.append(v)
            if u is not None:
                u8 = u
        if keep_v:
            v_trajs.append(torch.zeros_like(x))
        return x, x_trajs, v_trajs, u8
