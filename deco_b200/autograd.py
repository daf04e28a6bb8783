"""Training path of the class-conditional denoiser: forward that keeps what the backward needs, and the backward itself.

The reference trains `PixNerDiT` with plain PyTorch autograd (`src/diffusion/flow_matching/training_repa_DeCo.py:257`
calls `net(x_t, t, y)` and Lightning calls `loss.backward()`).  Here the whole denoiser is ONE autograd node
(`DenoiserFn`): its forward runs one kernel per reference op and stashes the activations, its backward walks the graph by
hand -- every contraction (dgrad `dX = dY.W`, wgrad `dW = dY^T.X`) on the tcgen05 GEMM, attention and the pixel decoder on
their own backward kernels (csrc/attention_bwd.cu, csrc/decoder_bwd.cu), the memory-bound glue in csrc/backward.cu.
PyTorch only allocates, slices and copies.  Gradients are returned per parameter in `named_parameters()` order.

Per block (dit_c2i_DeCo.py:206-210), forward keeps: the stream before each branch (fp32), the two modulated norms h1/h2,
the raw and the normalised+rotated qkv, the attention output o, the branch outputs a1/a2 (for the gate gradients), the
SwiGLU pre-activation y13 and its output u.
"""
from __future__ import annotations

import os
from typing import Dict, List

import torch

from . import _lib, ops

bf16 = torch.bfloat16
F32 = torch.float32
WGRAD = os.environ.get("DECO_B200_WGRAD", "tn")                # "tn" (product path) | "transpose" (A/B check)
DECODER_BWD = os.environ.get("DECO_B200_DECODER_BWD", "mma")   # "mma" (product path) | "scalar" (A/B check)
FUSE_GATE_NORM = os.environ.get("DECO_B200_FUSE_GATE_NORM", "1") != "0"   # residual add + next norm in one kernel
# SwiGLU inside GEMM epilogues: "fwd" = w13 GEMM leaves y13 and u; "1" = also the dgrad GEMM du -> dy13; "0" = stand-alone kernels
_FS = os.environ.get("DECO_B200_FUSE_SWIGLU", "fwd")
FUSE_SWIGLU, FUSE_SWIGLU_BWD = _FS != "0", _FS == "1"
WGRAD_STREAM = os.environ.get("DECO_B200_WGRAD_STREAM", "1") != "0"       # weight-gradient GEMMs on a second stream


# Called with lists of gradient tensors the moment they are final (per DiT block, walking backwards, then the tail): lets a
# data-parallel trainer start averaging them while the rest of the backward still runs (DDP's bucketed overlap for the
# reference; deco_b200.distributed.OverlappedGradientAverager).  None = no hook.
GRAD_READY_HOOK = None


def _grads_ready(tensors) -> None:
    if GRAD_READY_HOOK is not None:
        GRAD_READY_HOOK([t for t in tensors if t is not None])


def _rb(t: torch.Tensor) -> torch.Tensor:
    """fp32 copy holding bf16-rounded values (what the forward MMAs see)."""
    return t.detach().to(bf16).to(F32)


@torch.no_grad()
def pack_decoder_train(module, device):
    """fp32 weight blob of the pixel decoder in the layout of csrc/decoder_bwd.cu."""
    C = module.in_channels
    dn = module.dec_net
    wx = _rb(module.x_embedder.embedder[0].weight)
    parts = [wx[:, :C].reshape(-1), _rb(dn.input_proj.weight).reshape(-1), dn.input_proj.bias.detach().float()]
    for blk in dn.res_blocks:
        parts += [_rb(blk.adaLN_modulation[1].weight).reshape(-1), blk.adaLN_modulation[1].bias.detach().float(),
                  blk.in_ln.weight.detach().float(), blk.in_ln.bias.detach().float(),
                  _rb(blk.mlp[0].weight).reshape(-1), blk.mlp[0].bias.detach().float(),
                  _rb(blk.mlp[2].weight).reshape(-1), blk.mlp[2].bias.detach().float()]
    wf = torch.zeros(4, 32, device=wx.device)
    wf[:C] = _rb(dn.final_layer.linear.weight)
    bfin = torch.zeros(4, device=wx.device)
    bfin[:C] = dn.final_layer.linear.bias.detach().float()
    parts += [wf.reshape(-1), bfin]
    blob = torch.cat([p.to(device=device, dtype=F32).reshape(-1) for p in parts]).contiguous()
    assert blob.numel() == _lib.load().deco_decoder_train_blob_floats(len(dn.res_blocks)), blob.numel()
    return blob


_FRAG_T_INDEX: dict = {}


def _frag_t(w: torch.Tensor, perm_n: bool = False) -> torch.Tensor:
    """Pack the B operand of dX = dY . W, i.e. W^T given as wt [n = 32 inputs, K = outputs] (bf16), into m16n8k16
    B-fragment order [n_tile, k_step, lane, 4] (see denoiser.py::_frag).  perm_n: n-tile j, column c <-> input channel
    8 (c / 2) + 2 j + (c % 2), which makes a lane's four accumulator pairs the 16-byte chunk of the condition
    (csrc/decoder_bwd_mma.cu).  The gather runs on w's device (index tensors cached)."""
    n_out, K = w.shape
    key = (n_out, K, perm_n, str(w.device))
    if key not in _FRAG_T_INDEX:
        j = torch.arange(n_out // 8).view(-1, 1, 1, 1)
        s = torch.arange(K // 16).view(1, -1, 1, 1)
        lane = torch.arange(32).view(1, 1, -1, 1)
        e = torch.arange(4).view(1, 1, 1, -1)
        g, t = lane // 4, lane % 4
        half, lo = e // 2, e % 2
        k = 16 * s + 8 * half + 2 * t + lo
        row = (8 * (g // 2) + 2 * j + (g % 2)) if perm_n else (8 * j + g)
        shape = (n_out // 8, K // 16, 32, 4)
        _FRAG_T_INDEX[key] = (row.expand(shape).contiguous().to(w.device), k.expand(shape).contiguous().to(w.device))
    row, k = _FRAG_T_INDEX[key]
    return w[row, k].contiguous()


@torch.no_grad()
def pack_decoder_bwd(module, device):
    """Backward blob of csrc/decoder_bwd_mma.cu: W^T fragments [WinT | R x (WadaT n-permuted, W0T, W2T)] then fp32 Wf[3][32].
    Packed on `device` (no host synchronisation: a training loop re-packs after every optimizer step)."""
    dn = module.dec_net

    def rb16(t):
        return t.detach().to(device=device, dtype=torch.float32).to(bf16)

    frags = [_frag_t(rb16(dn.input_proj.weight).t().contiguous())]
    for blk in dn.res_blocks:
        frags += [_frag_t(rb16(blk.adaLN_modulation[1].weight).t().contiguous(), perm_n=True),
                  _frag_t(rb16(blk.mlp[0].weight).t().contiguous()), _frag_t(rb16(blk.mlp[2].weight).t().contiguous())]
    wf = rb16(dn.final_layer.linear.weight).float()
    assert wf.shape == (3, 32)
    blob = torch.cat([torch.cat([f.reshape(-1) for f in frags]).view(torch.uint8), wf.reshape(-1).contiguous().view(torch.uint8)])
    assert blob.numel() == _lib.load().deco_decoder_bwd_blob_bytes(len(dn.res_blocks)), blob.numel()
    return blob.contiguous()


@torch.no_grad()
def prepare_train(module, P: dict, device) -> dict:
    """Transposed bf16 weights for the dgrad GEMMs + the fp32 decoder blob; cached next to `prepare()`'s dict."""
    if "train" in P:
        return P["train"]
    T = {}

    def tr(w):   # [N, K] bf16 -> [K, N] bf16 (the "weight" operand of dX = dY . W)
        return ops.transpose_cast(w, rows_pad=w.shape[0])

    T["wt2T"], T["wadaT"], T["wcondT"] = tr(P["wt2"]), tr(P["wada"]), tr(P["wcond"])
    T["blocks"] = [dict(wqkvT=tr(bp["wqkv"]), wprojT=tr(bp["wproj"]), w13T=tr(bp["w13"]), w2T=tr(bp["w2"]))
                   for bp in P["blocks"]]
    T["dec_blob"] = pack_decoder_train(module, device)
    T["dec_bwd_blob"] = pack_decoder_bwd(module, device)
    key = ("nerf_tabT", str(device))
    if key not in module.precompute_pos:        # constant: transposed positional table [64, 256] bf16, cached on the device
        from .denoiser import nerf_pos_table
        module.precompute_pos[key] = ops.transpose_cast(nerf_pos_table(module.patch_size, module.x_embedder.max_freqs).to(device))
    T["tabT"] = module.precompute_pos[key]
    P["train"] = T
    return T


def _wgrad(dy: torch.Tensor, x: torch.Tensor, deinterleave16: bool = False) -> torch.Tensor:
    """dW [N, K] fp32 = dy^T [N, M] . x [M, K], M = tokens.  WGRAD = "tn": the GEMM reads dy and x as they lie in memory
    (MN-major operands); "transpose": both operands are first copied K-major (K = M padded to 8).  deinterleave16: the N
    rows are the [16 x w1 | 16 x w3] rows of the SwiGLU weight and come back as (dW1 ; dW3) stacked."""
    if WGRAD == "tn":
        if dy.dtype != bf16:
            dy = ops.cast_bf16(dy.contiguous())
        return ops.gemm_tn(dy, x, deinterleave16=deinterleave16)
    dw = ops.gemm(ops.transpose_cast(dy), ops.transpose_cast(x), None, ops.EPI_BIAS_F32)
    if deinterleave16:
        n, k = dw.shape
        dw = dw.view(n // 32, 2, 16, k).transpose(0, 1).reshape(n, k)
    return dw


_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


class WgradLane:
    """The weight-gradient GEMMs of the DiT blocks on a SECOND stream.  Nothing on the backward's critical path reads a
    dW (autograd only collects them at the end), while the chain dY -> dX -> norm / gate / SwiGLU / attention backward is
    half memory-bound glue that leaves the tensor pipe idle: the wgrad GEMMs (a third of the backward's MMA work) fill
    those holes instead of queueing behind them.  Works inside CUDA-graph capture (fork / join through events).

    Allocation discipline: results are allocated on the MAIN stream before the fork and operands are kept referenced
    until the main stream has waited for the GEMMs that read them (`sync(lag)`), so the caching allocator never hands a
    block to one stream while the other still uses it."""

    def __init__(self, dev: torch.device, enabled: bool):
        self.main = torch.cuda.current_stream(dev)
        self.side = None
        if enabled:
            key = dev.index if dev.index is not None else torch.cuda.current_device()
            if key not in _SIDE_STREAMS:
                _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
            self.side = _SIDE_STREAMS[key]
        self.pending: List[tuple] = []       # (event recorded on the side stream, tensors it covers)
        self.batch: List[torch.Tensor] = []

    def wgrad(self, dy: torch.Tensor, x: torch.Tensor, deinterleave16: bool = False) -> torch.Tensor:
        """dW [N, K] fp32 = dy^T . x (both bf16, as they lie in memory), launched on the side stream behind everything the
        main stream has enqueued so far.  deinterleave16: dy's columns are the interleaved [16 x w1 | 16 x w3] SwiGLU
        columns; the result is (dW1 ; dW3) stacked."""
        if self.side is None:
            return _wgrad(dy, x, deinterleave16)
        out = torch.empty((dy.shape[1], x.shape[1]), dtype=F32, device=dy.device)
        self.side.wait_stream(self.main)
        with torch.cuda.stream(self.side):
            ops.gemm_tn(dy, x, out=out, deinterleave16=deinterleave16)
        self.batch += [dy, x, out]
        return out

    def mark(self) -> None:
        """Close a batch (one DiT block): one event behind its GEMMs."""
        if self.side is None or not self.batch:
            return
        ev = torch.cuda.Event()
        ev.record(self.side)
        self.pending.append((ev, self.batch))
        self.batch = []

    def sync(self, lag: int = 0) -> None:
        """Main stream waits for all but the `lag` most recent batches; their operands may be released afterwards."""
        while len(self.pending) > lag:
            ev, _ = self.pending.pop(0)
            self.main.wait_event(ev)


def _colsum(x: torch.Tensor) -> torch.Tensor:
    return ops.colsum_(torch.zeros(x.shape[1], dtype=F32, device=x.device), x)


def _cond_forward(module, P: dict, t, y, S: dict):
    """t sinusoid -> MLP -> + label embedding -> silu -> the adaLN GEMM of every block (+ whatever rows the model appended
    to `wada`); keeps what `_cond_backward` needs.  Returns (temb, mod)."""
    tfreq = ops.timestep_freq(t, module.t_embedder.frequency_embedding_size)
    z1 = ops.gemm(tfreq, P["wt0"], P["bt0"], ops.EPI_BIAS)
    h1t = ops.gemm(tfreq, P["wt0"], P["bt0"], ops.EPI_BIAS_SILU)
    temb = ops.gemm(h1t, P["wt2"], P["bt2"], ops.EPI_BIAS)
    c = ops.cond_combine(temb, P["ytab"], y)
    mod = ops.gemm(c, P["wada"], P["bada"], ops.EPI_BIAS)
    S.update(tfreq=tfreq, z1=z1, h1t=h1t, temb=temb, c=c, mod=mod)
    return temb, mod


def _blocks_forward(bps: list, mod, s, B: int, L: int, H: int, heads: int, pos, joint=None):
    """The AdaLN DiT blocks (dit_c2i_DeCo.py:194-210 == dit_c2i_baseline.py:194-210; dit_t2i_pixnerd.py:65-81, :176-198) on
    the fp32 stream s [B*L, H], one kernel per reference op, keeping per block what the backward reads.  bps = the prepared
    weights of the blocks, mod = their [B, nb*6H] modulation columns, pos = RoPE table or None (the t2i text blocks).
    joint = (ytxt bf16 [B*T, H], T): the t2i joint attention (dit_t2i_pixnerd.py:43-59) -- every block adds the keys / values
    kv_y(ytxt) (k-normed, no RoPE) behind the image keys.  Returns (s_out, saved blocks)."""
    d = H // heads
    dev = s.device
    blocks = []
    nb = len(bps)

    def mod6(i):
        m = mod[:, i * 6 * H:(i + 1) * 6 * H]
        return tuple(m[:, j * H:(j + 1) * H] for j in range(6))

    h1 = None
    for i, bp in enumerate(bps):
        sh1, sc1, g1, sh2, sc2, g2 = mod6(i)
        if h1 is None:
            h1 = ops.rmsnorm_modulate(s, bp["n1"], sh1, sc1, L)
        qkv_raw = ops.gemm(h1, bp["wqkv"], None, ops.EPI_BIAS)
        qkv = torch.empty_like(qkv_raw)          # q, k normalised + rotated here; v is read from the raw GEMM output
        ops.qknorm_rope_to(qkv_raw, qkv, bp["qn"], bp["kn"], pos, heads, d, L)
        extra = {}
        if joint is None:
            o, lse = ops.attention_lse(qkv[:, :H], qkv[:, H:2 * H], qkv_raw[:, 2 * H:], B, heads, d)
        else:
            # keys / values = [image (L) || text (T)] per image, as one segment of L + T rows (copies: the backward kernels
            # take one key segment)
            ytxt, Tt = joint
            kvy_raw = ops.gemm(ytxt, bp["wkvy"], None, ops.EPI_BIAS)                  # [B*T, 2H]
            kvy = kvy_raw.clone()
            ops.headnorm_rope_(kvy, 0, bp["kn"], heads, d, Tt)
            kc = torch.cat([qkv.view(B, L, 3 * H)[:, :, H:2 * H], kvy.view(B, Tt, 2 * H)[:, :, :H]], 1).view(B * (L + Tt), H)
            vc = torch.cat([qkv_raw.view(B, L, 3 * H)[:, :, 2 * H:], kvy.view(B, Tt, 2 * H)[:, :, H:]], 1).view(B * (L + Tt), H)
            o, lse = ops.attention_lse(qkv[:, :H], kc, vc, B, heads, d)
            extra = dict(kvy_raw=kvy_raw, kc=kc, vc=vc)
        a1 = ops.gemm(o, bp["wproj"], bp["bproj"], ops.EPI_BIAS)
        # residual add + the norm that feeds the next GEMM in one pass over the row (FUSE_GATE_NORM=0: the two kernels)
        if FUSE_GATE_NORM:
            s_mid, h2 = ops.gate_residual_norm(s, a1, g1, L, bp["n2"], sh2, sc2)
        else:
            s_mid = ops.gate_residual(s, a1, g1, L)
            h2 = ops.rmsnorm_modulate(s_mid, bp["n2"], sh2, sc2, L)
        if FUSE_SWIGLU:     # one GEMM leaves both the pre-activation (for the backward) and u
            y13 = torch.empty((h2.shape[0], bp["w13"].shape[0]), dtype=bf16, device=dev)
            u = ops.gemm(h2, bp["w13"], None, ops.EPI_SWIGLU_DUAL, aux=y13)
        else:
            y13 = ops.gemm(h2, bp["w13"], None, ops.EPI_BIAS)
            u = ops.swiglu_fwd(y13)
        a2 = ops.gemm(u, bp["w2"], None, ops.EPI_BIAS)
        h1_next = None
        if FUSE_GATE_NORM and i + 1 < nb:
            nsh1, nsc1 = mod6(i + 1)[:2]
            s_out, h1_next = ops.gate_residual_norm(s_mid, a2, g2, L, bps[i + 1]["n1"], nsh1, nsc1)
        else:
            s_out = ops.gate_residual(s_mid, a2, g2, L)
        blocks.append(dict(s_in=s, h1=h1, qkv_raw=qkv_raw, qkv=qkv, o=o, lse=lse, a1=a1, s_mid=s_mid, h2=h2, y13=y13, u=u, a2=a2,
                           **extra))
        s, h1 = s_out, h1_next
    return s, blocks


def train_forward(module, x32, t, y):
    """Returns (out fp32 [B,3,H,W], saved dict)."""
    P = module.prepare(x32.device)
    prepare_train(module, P, x32.device)
    B, _, Hh, Ww = x32.shape
    p, H, heads = module.patch_size, module.hidden_size, module.num_groups
    L = (Hh // p) * (Ww // p)
    dev = x32.device
    pos = module.fetch_pos(Hh // p, Ww // p, dev)
    S = dict(B=B, L=L, x32=x32, y=y, pos=pos)
    xp = ops.patchify(x32, p)
    temb, mod = _cond_forward(module, P, t, y, S)
    s = ops.gemm(xp, P["ws"], P["bs"], ops.EPI_BIAS_F32)
    S.update(xp=xp)
    s, blocks = _blocks_forward(P["blocks"], mod, s, B, L, H, heads, pos)
    s2 = ops.silu_add_rows(s, temb, L)
    ycond = ops.gemm(s2, P["wcond"], P["bcond"], ops.EPI_BIAS)
    R = module.num_blocks - module.num_cond_blocks
    out = ops.pixel_decoder(x32, ycond, P["blob"], P["postab"], p, module.hidden_size_x, R, out_dtype=F32)
    S.update(blocks=blocks, s_final=s, s2=s2, ycond=ycond)
    return out, S


def _make_lane(dev) -> "WgradLane":
    # wgrad GEMMs on a second stream (not under the per-GEMM event probe of the bench's roofline pass, whose launch durations
    # must not overlap other work, and not for the A/B transpose variant)
    return WgradLane(dev, WGRAD_STREAM and WGRAD == "tn" and ops.gemm_probe is None)


def _blocks_backward(module, P: dict, T: dict, S: dict, ds: torch.Tensor, dmod: torch.Tensor, G: Dict[str, torch.Tensor],
                     lane: "WgradLane" = None, spec: dict = None) -> None:
    """Backward of `_blocks_forward`: walks the blocks in reverse, updates the stream gradient ds [B*L, H] fp32 in place,
    accumulates the modulation gradients into dmod[:, :nb*6H] and leaves the blocks' parameter gradients in G.
    spec (t2i): which block list this is -- bps / bts (prepared weights and their transposes), saved (the forward's per-block
    dicts), mod (the blocks' modulation columns; dmod is the matching view), L (rows per image), pos, prefix, t2i = True
    (parameter names of dit_t2i_pixnerd.py: qkv_x / qkv, w12, w3), joint = dict(ytxt, T, dyt fp32 [B*T, H] accumulated, ones)."""
    spec = spec or {}
    dev = ds.device
    B, L = S["B"], spec.get("L", S["L"])
    H, heads = module.hidden_size, module.num_groups
    d = H // heads
    bps, bts, saved = spec.get("bps", P.get("blocks")), spec.get("bts", T.get("blocks")), spec.get("saved", S.get("blocks"))
    prefix, t2i, joint = spec.get("prefix", "blocks."), spec.get("t2i", False), spec.get("joint")
    pos = spec["pos"] if "pos" in spec else S["pos"]
    qkv_name = ("attn.qkv_x.weight" if joint is not None else "attn.qkv.weight")
    nb = len(bps)
    mod = spec.get("mod", S["mod"])
    z = lambda *shape: torch.zeros(shape, dtype=F32, device=dev)   # noqa: E731
    if lane is None:
        lane = _make_lane(dev)
    ready: List[tuple] = []     # per finished block: its matrix gradients, announced once the lane delivered
    zblk = z(max(nb, 1), 3 * H + 2 * d)       # one fill for the per-block vector gradients (norm1/2, proj bias, q/k-norm)
    da2 = None
    for i in reversed(range(nb)):
        bp, bt, sv = bps[i], bts[i], saved[i]
        pre = f"{prefix}{i}."
        m = mod[:, i * 6 * H:(i + 1) * 6 * H]
        dm = dmod[:, i * 6 * H:(i + 1) * 6 * H]
        sc1, g1, sc2, g2 = m[:, H:2 * H], m[:, 2 * H:3 * H], m[:, 4 * H:5 * H], m[:, 5 * H:6 * H]
        dsh1, dsc1, dg1, dsh2, dsc2, dg2 = (dm[:, j * H:(j + 1) * H] for j in range(6))
        Fp = bp["w2"].shape[1]                       # SwiGLU width as prepared (padded to a multiple of 16)
        F_ = Fp if t2i else module.blocks[i].mlp.w1.weight.shape[0]
        # MLP branch (da2: gate_bwd of this block's second residual add -- computed by the fused kernel at the end of the
        # previous iteration, or here for the last block)
        if da2 is None:
            da2 = ops.gate_bwd(ds, sv["a2"], g2, dg2, L)
        gw2 = lane.wgrad(da2, sv["u"])
        if FUSE_SWIGLU_BWD:     # du never reaches memory: the dgrad GEMM's epilogue turns it into dy13
            dy13 = ops.gemm(da2, bt["w2T"], None, ops.EPI_SWIGLU_BWD, aux=sv["y13"])
            du = None
        else:
            du = ops.gemm(da2, bt["w2T"], None, ops.EPI_BIAS)
            dy13 = ops.swiglu_bwd(sv["y13"], du)
        gw13 = lane.wgrad(dy13, sv["h2"], deinterleave16=True)      # (dW1 ; dW3) stacked: no copy to pull them apart
        dh2 = ops.gemm(dy13, bt["w13T"], None, ops.EPI_BIAS)
        dn2, dn1, dbproj, dqn, dkn = zblk[i, :H], zblk[i, H:2 * H], zblk[i, 2 * H:3 * H], zblk[i, 3 * H:3 * H + d], zblk[i, 3 * H + d:]
        # norm2 backward + the gate backward of the attention branch's residual add, one pass over ds
        if FUSE_GATE_NORM:
            da1 = ops.rmsnorm_modulate_bwd_gate_(ds, dh2, sv["s_mid"], bp["n2"], sc2, dn2, dsh2, dsc2, L, sv["a1"], g1, dg1,
                                                 dbias=dbproj)
        else:
            ops.rmsnorm_modulate_bwd_(ds, dh2, sv["s_mid"], bp["n2"], sc2, dn2, dsh2, dsc2, L)
            da1 = ops.gate_bwd(ds, sv["a1"], g1, dg1, L, dbias=dbproj)
        G[pre + "norm2.weight"] = dn2
        del da2, du, dy13, dh2
        # attention branch
        G[pre + "attn.proj.bias"] = dbproj
        gwproj = lane.wgrad(da1, sv["o"])
        do = ops.gemm(da1, bt["wprojT"], None, ops.EPI_BIAS)
        qkv = sv["qkv"]
        dqkv = torch.empty_like(qkv)
        gwkvy = None
        if joint is None:
            ops.attention_bwd(qkv[:, :H], qkv[:, H:2 * H], sv["qkv_raw"][:, 2 * H:], sv["o"], do,
                              dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], B, heads, d, lse=sv["lse"])
        else:
            # one key segment [image || text]: its gradient is pulled apart again (copies); the text half goes back through
            # the k-norm (the SAME k_norm weight as the image keys: both accumulate into dkn) and kv_y into the text stream
            Tt = joint["T"]
            dkc, dvc = torch.empty_like(sv["kc"]), torch.empty_like(sv["vc"])
            ops.attention_bwd(qkv[:, :H], sv["kc"], sv["vc"], sv["o"], do, dqkv[:, :H], dkc, dvc, B, heads, d, lse=sv["lse"])
            dq3, dk3, dv3 = dqkv.view(B, L, 3 * H), dkc.view(B, L + Tt, H), dvc.view(B, L + Tt, H)
            dq3[:, :, H:2 * H].copy_(dk3[:, :L])
            dq3[:, :, 2 * H:].copy_(dv3[:, :L])
            dkvy = torch.empty((B * Tt, 2 * H), dtype=bf16, device=dev)
            dkvy.view(B, Tt, 2 * H)[:, :, :H].copy_(dk3[:, L:])
            dkvy.view(B, Tt, 2 * H)[:, :, H:].copy_(dv3[:, L:])
            ops.headnorm_rope_bwd_(dkvy, sv["kvy_raw"], 0, bp["kn"], None, dkn, heads, d, Tt)
            gwkvy = lane.wgrad(dkvy, joint["ytxt"])
            ops.gemm(dkvy, bt["wkvyT"], None, ops.EPI_GATE_RESIDUAL, out=joint["dyt"], resid=joint["dyt"], gate=joint["ones"],
                     rows_per_gate=B * Tt)            # dyt += dkvy . Wkvy, accumulated in fp32 over the blocks
            del dkc, dvc, dkvy
        ops.headnorm_rope_bwd_(dqkv, sv["qkv_raw"], 0, bp["qn"], pos, dqn, heads, d, L)
        ops.headnorm_rope_bwd_(dqkv, sv["qkv_raw"], H, bp["kn"], pos, dkn, heads, d, L)
        G[pre + "attn.q_norm.weight"], G[pre + "attn.k_norm.weight"] = dqn, dkn
        gwqkv = lane.wgrad(dqkv, sv["h1"])
        lane.mark()
        dh1 = ops.gemm(dqkv, bt["wqkvT"], None, ops.EPI_BIAS)
        da2 = None
        if FUSE_GATE_NORM and i > 0:      # norm1 backward + the gate backward of block i - 1's second residual add
            mp, dmp = mod[:, (i - 1) * 6 * H:i * 6 * H], dmod[:, (i - 1) * 6 * H:i * 6 * H]
            da2 = ops.rmsnorm_modulate_bwd_gate_(ds, dh1, sv["s_in"], bp["n1"], sc1, dn1, dsh1, dsc1, L, saved[i - 1]["a2"],
                                                 mp[:, 5 * H:6 * H], dmp[:, 5 * H:6 * H])
        else:
            ops.rmsnorm_modulate_bwd_(ds, dh1, sv["s_in"], bp["n1"], sc1, dn1, dsh1, dsc1, L)
        G[pre + "norm1.weight"] = dn1
        del da1, do, dqkv, dh1
        saved[i] = None   # release the block's activations (the lane keeps what its GEMMs still read)
        # The main stream joins the lane one block late, so that a block's wgrad GEMMs overlap the next block's chain; the
        # matrix gradients of a joined block are final (w1 / w3 de-interleaved on the main stream) and are announced; the
        # vector gradients live in zblk / dmod and follow with the tail.
        ready.append((i, F_, gw2, gw13, gwproj, gwqkv, gwkvy))
        del gw2, gw13, gwproj, gwqkv, gwkvy
        lane.sync(lag=1 if i > 0 else 0)
        while len(ready) > (1 if (i > 0 and lane.side is not None) else 0):
            j, Fj, w2g, w13g, wpg, wqg, wkg = ready.pop(0)
            prej = f"{prefix}{j}."
            if t2i:     # layers/swiglu.py:4-17: w12 = (w1 ; w2) stacked, w3 = the down projection
                names = ["mlp.w3.weight", "mlp.w12.weight", "attn.proj.weight", qkv_name]
                G[prej + "mlp.w3.weight"], G[prej + "mlp.w12.weight"] = w2g, w13g
                if wkg is not None:
                    G[prej + "attn.kv_y.weight"] = wkg
                    names.append("attn.kv_y.weight")
            else:
                names = ["mlp.w2.weight", "mlp.w1.weight", "mlp.w3.weight", "attn.proj.weight", qkv_name]
                G[prej + "mlp.w2.weight"] = w2g if Fj == Fp else w2g[:, :Fj].contiguous()
                G[prej + "mlp.w1.weight"], G[prej + "mlp.w3.weight"] = w13g[:Fj], w13g[Fp:Fp + Fj]
            G[prej + "attn.proj.weight"], G[prej + qkv_name] = wpg, wqg
            _grads_ready([G[prej + k] for k in names])



def _cond_backward(module, P: dict, T: dict, S: dict, dmod, dtemb, G: Dict[str, torch.Tensor], ada_names: List[str],
                   ada_rows: int, extra_ada: tuple = (), label_table: bool = True) -> None:
    """Backward of `_cond_forward`: the adaLN GEMM (rows of `wada`: `ada_rows` per name in `ada_names`, then the
    (name, rows) pairs of `extra_ada`), c = silu(temb + table[y]), the timestep MLP.  dtemb [B, H] fp32 holds whatever the
    rest of the model already sent to temb."""
    B = S["B"]
    H = module.hidden_size
    nb = len(ada_names)
    dev = dmod.device
    z = lambda *shape: torch.zeros(shape, dtype=F32, device=dev)   # noqa: E731
    # ---- adaLN of all blocks: mod = c . wada^T + bada
    if dmod.shape[1]:
        dwada = _wgrad(dmod, S["c"])                            # [nb*6H, H]
        dbada = _colsum(dmod)
        r0 = 0
        for name, rows in [(n, ada_rows) for n in ada_names] + list(extra_ada):
            G[name + ".weight"], G[name + ".bias"] = dwada[r0:r0 + rows], dbada[r0:r0 + rows]
            r0 += rows
        dc = ops.gemm_f32_splitk(ops.cast_bf16(dmod), T["wadaT"])          # M = batch, K = nb*6H: split-K
    else:
        dc = z(B, H)
    # ---- c = silu(temb + table[y])
    dtab = torch.zeros_like(P["ytab"])
    ops.cond_combine_bwd_(dc, S["temb"], P["ytab"], S["y"], dtemb, dtab)
    if label_table:      # (the t2i model conditions on silu(temb) alone: its "table" is one zero row)
        G["y_embedder.embedding_table.weight"] = dtab
    # ---- t_embedder: temb = silu(tfreq . wt0^T + bt0) . wt2^T + bt2
    G["t_embedder.mlp.2.weight"] = _wgrad(dtemb, S["h1t"])
    G["t_embedder.mlp.2.bias"] = _colsum(dtemb)
    dh1t = ops.gemm(ops.cast_bf16(dtemb), T["wt2T"], None, ops.EPI_BIAS)
    dz1 = ops.silu_bwd(S["z1"], dh1t)
    G["t_embedder.mlp.0.weight"] = _wgrad(dz1, S["tfreq"])
    G["t_embedder.mlp.0.bias"] = _colsum(dz1)


def _decoder_backward(module, P: dict, T: dict, S: dict, dout: torch.Tensor, G: Dict[str, torch.Tensor], hidden_x: int,
                      R: int) -> torch.Tensor:
    """Backward of the pixel decoder + NerfEmbedder (dit_c2i_DeCo.py:212-248, :288-415): fills G with their parameter
    gradients and returns dycond bf16 [B*L, p*p*32]."""
    x32 = S["x32"]
    dev = x32.device
    p = module.patch_size
    C = module.in_channels
    z = lambda *shape: torch.zeros(shape, dtype=F32, device=dev)   # noqa: E731
    # ---- pixel decoder (+ NerfEmbedder)
    if DECODER_BWD == "scalar":     # fp32 scalar kernel (csrc/decoder_bwd.cu): the check the MMA kernel is validated against
        dycond, gdec = ops.pixel_decoder_bwd(x32, S["ycond"], dout.to(F32).contiguous(), T["dec_blob"], P["postab"], p,
                                             hidden_x, R)
    else:
        dycond, gdec = ops.pixel_decoder_bwd_tc(x32, S["ycond"], dout.to(F32).contiguous(), P["blob"], T["dec_bwd_blob"],
                                                P["postab"], p, hidden_x, R)
    nW = T["dec_blob"].numel()
    dpostab = gdec[nW:].view(p * p, 32)
    gx = z(32, C + module.x_embedder.max_freqs ** 2)
    gx[:, :C] = gdec[0:96].view(32, 3)[:, :C]
    gx[:, C:] = ops.gemm(ops.transpose_cast(dpostab), T["tabT"], None, ops.EPI_BIAS_F32)      # [32, 64]
    G["x_embedder.embedder.0.weight"] = gx
    G["x_embedder.embedder.0.bias"] = _colsum(dpostab)
    G["dec_net.input_proj.weight"] = gdec[96:1120].view(32, 32)
    G["dec_net.input_proj.bias"] = gdec[1120:1152]
    for j in range(R):
        o0 = 1152 + j * 5344
        pre = f"dec_net.res_blocks.{j}."
        G[pre + "adaLN_modulation.1.weight"] = gdec[o0:o0 + 3072].view(96, 32)
        G[pre + "adaLN_modulation.1.bias"] = gdec[o0 + 3072:o0 + 3168]
        G[pre + "in_ln.weight"] = gdec[o0 + 3168:o0 + 3200]
        G[pre + "in_ln.bias"] = gdec[o0 + 3200:o0 + 3232]
        G[pre + "mlp.0.weight"] = gdec[o0 + 3232:o0 + 4256].view(32, 32)
        G[pre + "mlp.0.bias"] = gdec[o0 + 4256:o0 + 4288]
        G[pre + "mlp.2.weight"] = gdec[o0 + 4288:o0 + 5312].view(32, 32)
        G[pre + "mlp.2.bias"] = gdec[o0 + 5312:o0 + 5344]
    of = 1152 + R * 5344
    G["dec_net.final_layer.linear.weight"] = gdec[of:of + 128].view(4, 32)[:C]
    G["dec_net.final_layer.linear.bias"] = gdec[of + 128:of + 128 + C]

    return dycond


def train_backward(module, S: dict, dout: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Gradients of every parameter (by name) for upstream gradient dout [B,3,H,W]."""
    x32 = S["x32"]
    dev = x32.device
    P = module.prepare(dev)
    T = P["train"]
    B, L = S["B"], S["L"]
    p, H = module.patch_size, module.hidden_size
    nb = len(P["blocks"])
    R = module.num_blocks - module.num_cond_blocks
    C = module.in_channels
    G: Dict[str, torch.Tensor] = {}
    z = lambda *shape: torch.zeros(shape, dtype=F32, device=dev)   # noqa: E731

    dycond = _decoder_backward(module, P, T, S, dout, G, module.hidden_size_x, R)

    # ---- cond_embed: ycond = s2 . wcond^T + bcond (its weight gradient, the largest single wgrad GEMM of the step, rides on
    # the second stream under the first blocks' chain)
    lane = _make_lane(dev)
    G["dec_net.cond_embed.weight"] = lane.wgrad(dycond, S["s2"])
    lane.mark()
    G["dec_net.cond_embed.bias"] = _colsum(dycond)
    ds2 = ops.gemm(dycond, T["wcondT"], None, ops.EPI_BIAS)
    del dycond
    # ---- s2 = silu(s + temb)
    dtemb = z(B, H)
    ds = ops.silu_add_rows_bwd(ds2, S["s_final"], S["temb"], dtemb, L)
    del ds2

    # ---- DiT blocks
    dmod = z(B, nb * 6 * H)
    _blocks_backward(module, P, T, S, ds, dmod, G, lane)
    lane.sync(0)
    # ---- s_embedder: s0 = xp . ws^T + bs
    G["s_embedder.proj.weight"] = _wgrad(ds, S["xp"])
    G["s_embedder.proj.bias"] = _colsum(ds)
    _cond_backward(module, P, T, S, dmod, dtemb, G, [f"blocks.{i}.adaLN_modulation.0" for i in range(nb)], 6 * H)
    _grads_ready(list(G.values()))      # the tail: everything not announced yet (the hook skips what it has seen)
    return G


# ------------------------------------------------------------------------------------------------------------------
# Patch-linear baseline (`FlattenDiT`, dit_c2i_baseline.py:357-379): the same blocks between a patch-embedding head and the
# AdaLN FinalLayer (:70-83) + Linear + fold tail.
@torch.no_grad()
def prepare_train_baseline(module, P: dict, device) -> dict:
    if "train" in P:
        return P["train"]

    def tr(w):
        return ops.transpose_cast(w, rows_pad=w.shape[0])

    T = dict(wt2T=tr(P["wt2"]), wadaT=tr(P["wada"]), wfinT=tr(P["wfin"]))
    T["blocks"] = [dict(wqkvT=tr(bp["wqkv"]), wprojT=tr(bp["wproj"]), w13T=tr(bp["w13"]), w2T=tr(bp["w2"]))
                   for bp in P["blocks"]]
    T["ones"] = torch.ones(module.hidden_size, dtype=F32, device=device)
    P["train"] = T
    return T


def baseline_train_forward(module, x32, t, y):
    """Returns (out bf16 [B,C,H,W] -- the reference's dtype under autocast --, saved dict)."""
    dev = x32.device
    P = module.prepare(dev)
    T = prepare_train_baseline(module, P, dev)
    B, Cc, Hh, Ww = x32.shape
    p, H, heads = module.patch_size, module.hidden_size, module.num_groups
    L = (Hh // p) * (Ww // p)
    nb = len(P["blocks"])
    pos = module.fetch_pos(Hh // p, Ww // p, dev)
    S = dict(B=B, L=L, y=y, pos=pos, shape=(B, Cc, Hh, Ww))
    xp = ops.patchify(x32, p)
    _, mod = _cond_forward(module, P, t, y, S)
    s = ops.gemm(xp, P["wx"], P["bx"], ops.EPI_BIAS_F32)
    s, blocks = _blocks_forward(P["blocks"], mod, s, B, L, H, heads, pos)
    # FinalLayer: LayerNorm (no affine) = RMSNorm of the centred row, then modulate with the last 2H adaLN columns
    shift, scale = mod[:, nb * 6 * H:nb * 6 * H + H], mod[:, nb * 6 * H + H:]
    xc = ops.center_rows(s)
    hf = ops.rmsnorm_modulate(xc, T["ones"], shift, scale, L)
    tok = ops.gemm(hf, P["wfin"], P["bfin"], ops.EPI_BIAS)
    out = ops.unpatchify(tok, B, Cc, Hh, Ww, p)
    S.update(xp=xp, blocks=blocks, xc=xc, hf=hf)
    return out, S


def baseline_train_backward(module, S: dict, dout: torch.Tensor) -> Dict[str, torch.Tensor]:
    dev = dout.device
    P = module.prepare(dev)
    T = P["train"]
    B, L = S["B"], S["L"]
    p, H = module.patch_size, module.hidden_size
    nb = len(P["blocks"])
    G: Dict[str, torch.Tensor] = {}
    z = lambda *shape: torch.zeros(shape, dtype=F32, device=dev)   # noqa: E731
    # ---- fold^T = unfold, then tok = hf . wfin^T + bfin
    dtok = ops.patchify(dout.to(F32).contiguous(), p)
    G["final_layer.linear.weight"] = _wgrad(dtok, S["hf"])
    G["final_layer.linear.bias"] = _colsum(dtok)
    dhf = ops.gemm(dtok, T["wfinT"], None, ops.EPI_BIAS)
    # ---- hf = rmsnorm(xc) (1 + scale) + shift ; xc = s - mean(s)
    dmod = z(B, nb * 6 * H + 2 * H)
    scale = S["mod"][:, nb * 6 * H + H:]
    ds = z(B * L, H)
    ops.rmsnorm_modulate_bwd_(ds, dhf, S["xc"], T["ones"], scale, z(H), dmod[:, nb * 6 * H:nb * 6 * H + H],
                              dmod[:, nb * 6 * H + H:], L)
    ops.center_rows(ds, out=ds)
    del dtok, dhf
    S["xc"] = S["hf"] = None
    _blocks_backward(module, P, T, S, ds, dmod, G)
    G["x_embedder.proj.weight"] = _wgrad(ds, S["xp"])
    G["x_embedder.proj.bias"] = _colsum(ds)
    _cond_backward(module, P, T, S, dmod, z(B, H), G, [f"blocks.{i}.adaLN_modulation.0" for i in range(nb)], 6 * H,
                   extra_ada=(("final_layer.adaLN_modulation.0", 2 * H),))
    _grads_ready(list(G.values()))
    return G


# ------------------------------------------------------------------------------------------------------------------
# Text-to-image denoiser (deco_b200/denoiser_t2i.py; dit_t2i_pixnerd.py:276-297 + SimpleMLPAdaLN): text embedding ->
# text-refine blocks -> joint-attention image blocks -> the DeCo pixel decoder.
@torch.no_grad()
def prepare_train_t2i(module, P: dict, device) -> dict:
    if "train" in P:
        return P["train"]

    def tr(w):
        return ops.transpose_cast(w, rows_pad=w.shape[0])

    T = dict(wt2T=tr(P["wt2"]), wadaT=tr(P["wada"]), wcondT=tr(P["wcond"]))
    T["text"] = [dict(wqkvT=tr(bp["wqkv"]), wprojT=tr(bp["wproj"]), w13T=tr(bp["w13"]), w2T=tr(bp["w2"])) for bp in P["text"]]
    T["blocks"] = [dict(wqkvT=tr(bp["wqkv"]), wprojT=tr(bp["wproj"]), w13T=tr(bp["w13"]), w2T=tr(bp["w2"]), wkvyT=tr(bp["wkvy"]))
                   for bp in P["blocks"]]
    T["dec_blob"] = pack_decoder_train(module, device)
    T["dec_bwd_blob"] = pack_decoder_bwd(module, device)
    from .denoiser_t2i import rope_cos_sin_ex2d
    tab = rope_cos_sin_ex2d(module.x_embedder.max_freqs ** 2 * 2, module.patch_size, module.patch_size)[..., 0]
    T["tabT"] = ops.transpose_cast(tab.to(device=device, dtype=F32).contiguous())
    T["ones_row"] = torch.ones(1, module.hidden_size, dtype=bf16, device=device)
    T["zero_mod"] = torch.zeros(1, module.hidden_size, dtype=bf16, device=device)
    P["train"] = T
    return T


def t2i_train_forward(module, x32, t, y):
    """Returns (out fp32 [B,3,H,W], saved dict); y = text-encoder states [B, T, txt_embed_dim]."""
    dev = x32.device
    P = module.prepare(dev)
    prepare_train_t2i(module, P, dev)
    B, _, Hh, Ww = x32.shape
    p, H, heads = module.patch_size, module.hidden_size, module.num_groups
    L = (Hh // p) * (Ww // p)
    Tt = y.shape[1]
    assert y.dim() == 3 and y.shape[0] == B and y.shape[2] == module.txt_embed_dim and Tt == module.txt_max_length
    nt = len(P["text"])
    pos = module.fetch_pos(Hh // p, Ww // p, dev)
    S = dict(B=B, L=L, T=Tt, x32=x32, y=torch.zeros(B, dtype=torch.int64, device=dev), pos=pos)
    Pc = dict(P, ytab=P["zero_row"])                       # c = silu(temb): the label table is one zero row
    temb, mod = _cond_forward(module, Pc, t, S["y"], S)
    # ---- text path: Linear -> RMSNorm -> + y_pos_embedding (fp32 stream), text-refine blocks (no RoPE)
    y16 = y.detach().reshape(B * Tt, module.txt_embed_dim).to(bf16).contiguous()
    yraw = ops.gemm(y16, P["wy"], P["by"], ops.EPI_BIAS_F32)
    ys = ops.rmsnorm_addpos(yraw, P["yn"], P["ypos"])
    ys, tblocks = _blocks_forward(P["text"], mod[:, :nt * 6 * H], ys, B, Tt, H, heads, None)
    ytxt = ops.cast_bf16(ys)
    # ---- image path: joint attention over [image || text] keys
    xp = ops.patchify(x32, p)
    s = ops.gemm(xp, P["ws"], P["bs"], ops.EPI_BIAS_F32)
    s, blocks = _blocks_forward(P["blocks"], mod[:, nt * 6 * H:], s, B, L, H, heads, pos, joint=(ytxt, Tt))
    s2 = ops.silu_add_rows(s, temb, L)
    ycond = ops.gemm(s2, P["wcond"], P["bcond"], ops.EPI_BIAS)
    out = ops.pixel_decoder(x32, ycond, P["blob"], P["postab"], p, module.decoder_hidden_size, module.num_decoder_blocks,
                            out_dtype=F32)
    S.update(y16=y16, yraw=yraw, tblocks=tblocks, ytxt=ytxt, xp=xp, blocks=blocks, s_final=s, s2=s2, ycond=ycond)
    return out, S


def t2i_train_backward(module, S: dict, dout: torch.Tensor) -> Dict[str, torch.Tensor]:
    dev = dout.device
    P = module.prepare(dev)
    T = P["train"]
    B, L, Tt = S["B"], S["L"], S["T"]
    H = module.hidden_size
    nt, ni = len(P["text"]), len(P["blocks"])
    G: Dict[str, torch.Tensor] = {}
    z = lambda *shape: torch.zeros(shape, dtype=F32, device=dev)   # noqa: E731
    dycond = _decoder_backward(module, P, T, S, dout, G, module.decoder_hidden_size, module.num_decoder_blocks)
    lane = _make_lane(dev)
    G["dec_net.cond_embed.weight"] = lane.wgrad(dycond, S["s2"])
    lane.mark()
    G["dec_net.cond_embed.bias"] = _colsum(dycond)
    ds2 = ops.gemm(dycond, T["wcondT"], None, ops.EPI_BIAS)
    del dycond
    dtemb = z(B, H)
    ds = ops.silu_add_rows_bwd(ds2, S["s_final"], S["temb"], dtemb, L)
    del ds2
    dmod = z(B, (nt + ni) * 6 * H)
    mod = S["mod"]
    # ---- image blocks; the text stream's gradient collects every block's kv_y branch
    dyt = z(B * Tt, H)
    _blocks_backward(module, P, T, S, ds, dmod[:, nt * 6 * H:], G, lane,
                     spec=dict(bps=P["blocks"], bts=T["blocks"], saved=S["blocks"], mod=mod[:, nt * 6 * H:], L=L, pos=S["pos"],
                               prefix="blocks.", t2i=True, joint=dict(ytxt=S["ytxt"], T=Tt, dyt=dyt, ones=T["ones_row"])))
    lane.sync(0)
    G["s_embedder.proj.weight"] = _wgrad(ds, S["xp"])
    G["s_embedder.proj.bias"] = _colsum(ds)
    del ds
    # ---- text-refine blocks (rows per image = T, no RoPE), then y_embedder: Linear -> RMSNorm (+ y_pos_embedding)
    _blocks_backward(module, P, T, S, dyt, dmod[:, :nt * 6 * H], G, None,
                     spec=dict(bps=P["text"], bts=T["text"], saved=S["tblocks"], mod=mod[:, :nt * 6 * H], L=Tt, pos=None,
                               prefix="text_refine_blocks.", t2i=True))
    G["y_pos_embedding"] = dyt.view(B, Tt, H).sum(0, keepdim=True)
    dyn, dyraw = z(H), z(B * Tt, H)
    zs = z(1, 2 * H)
    ops.rmsnorm_modulate_bwd_(dyraw, ops.cast_bf16(dyt), S["yraw"], P["yn"], T["zero_mod"].expand(B, H), dyn,
                              zs[:, :H].expand(B, H), zs[:, H:].expand(B, H), Tt)
    G["y_embedder.norm.weight"] = dyn
    G["y_embedder.proj.weight"] = _wgrad(dyraw, S["y16"])
    G["y_embedder.proj.bias"] = _colsum(dyraw)
    names = [f"text_refine_blocks.{i}.adaLN_modulation.0" for i in range(nt)] + [f"blocks.{i}.adaLN_modulation.0" for i in range(ni)]
    _cond_backward(module, dict(P, ytab=P["zero_row"]), T, S, dmod, dtemb, G, names, 6 * H, label_table=False)
    _grads_ready(list(G.values()))
    return G


def _train_fns(module):
    if hasattr(module, "text_refine_blocks"):
        return t2i_train_forward, t2i_train_backward
    if hasattr(module, "final_layer"):
        return baseline_train_forward, baseline_train_backward
    return train_forward, train_backward


class DenoiserFn(torch.autograd.Function):
    """out = net(x, t, y) as one autograd node (PixNerDiT, or the patch-linear FlattenDiT); *params only tell autograd
    which leaves receive gradients."""

    @staticmethod
    def forward(ctx, module, names: List[str], x, t, y, *params):
        x32 = x.detach().to(F32).contiguous()
        fwd = _train_fns(module)[0]
        yy = y.detach() if hasattr(module, "text_refine_blocks") else y.detach().reshape(-1)
        out, S = fwd(module, x32, t.detach().reshape(-1).to(F32), yy)
        ctx.module, ctx.names, ctx.S = module, names, S
        ctx.meta = [(p.requires_grad, p.shape, p.dtype) for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.S is None:
            raise RuntimeError("deco_b200 denoiser: backward called twice (activations are released after one pass)")
        G = _train_fns(ctx.module)[1](ctx.module, ctx.S, dout)
        ctx.S = None
        # an overlapped gradient averager reduces G's buffers in place on another stream: order this stream behind it
        # before autograd reads (or clones) them
        finish = getattr(GRAD_READY_HOOK, "finish", None)
        if finish is not None:
            finish()
        grads = []
        for n, (need, shape, dtype) in zip(ctx.names, ctx.meta):
            g = G.get(n) if need else None
            if need and g is None:
                raise RuntimeError(f"deco_b200 denoiser backward produced no gradient for {n}")
            grads.append(None if g is None else g.reshape(shape).to(dtype))
        return (None, None, None, None, None) + tuple(grads)


def denoiser_train_apply(module, x, t, y):
    names, params = zip(*module.named_parameters())
    return DenoiserFn.apply(module, list(names), x, t, y, *params)
