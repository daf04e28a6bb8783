"""Tensor-level wrappers over the C ABI (one function per entry point of include/deco_b200.h).

PyTorch is used only for device memory and streams; every function launches the hand-written sm_100a kernel on
`torch.cuda.current_stream()` and raises if the tensor is not on a CUDA device (there is no CPU path).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import call, ptr

EPI_BIAS, EPI_BIAS_SILU, EPI_GATE_RESIDUAL, EPI_SWIGLU, EPI_BIAS_F32, EPI_SWIGLU_DUAL, EPI_SWIGLU_BWD = 0, 1, 2, 3, 4, 5, 6
bf16 = torch.bfloat16
gemm_probe = None   # set to a deco_b200.utils.GemmProbe to time every GEMM launch with CUDA events
ATTN_BWD = __import__("os").environ.get("DECO_B200_ATTN_BWD", "tc")     # "tc" (tcgen05) | "legacy" (mma.sync)


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("deco_b200 kernels need CUDA tensors (sm_100a); there is no CPU fallback")


def _st(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, epilogue: int = EPI_BIAS,
         out: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
         gate: Optional[torch.Tensor] = None, rows_per_gate: int = 1, tile_n: int = 0,
         aux: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epilogue(a @ w.T): a [M,K] bf16 (row stride allowed), w [N,K] bf16, bias fp32 [N].
    gate: bf16 2-D view [M/rows_per_gate, N] with arbitrary row stride; resid fp32 [M,N].
    Output is bf16 except for EPI_GATE_RESIDUAL / EPI_BIAS_F32 (the fp32 residual stream).
    Training epilogues: EPI_SWIGLU_DUAL also WRITES the bf16 pre-activation [M,N] into `aux`; EPI_SWIGLU_BWD treats
    a @ w.T as du [M,N], reads the pre-activation `aux` [M,2N] and returns dy13 [M,2N]."""
    _cuda(a, w, bias, out, resid, gate, aux)
    assert a.dtype == bf16 and w.dtype == bf16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K, (a.shape, w.shape)
    n_out = N // 2 if epilogue in (EPI_SWIGLU, EPI_SWIGLU_DUAL) else (2 * N if epilogue == EPI_SWIGLU_BWD else N)
    odt = torch.float32 if epilogue in (EPI_GATE_RESIDUAL, EPI_BIAS_F32) else bf16
    if out is None:
        out = torch.empty((M, n_out), dtype=odt, device=a.device)
    assert out.dtype == odt and out.shape == (M, n_out) and out.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    ldr = gs = 0
    if epilogue == EPI_GATE_RESIDUAL:
        assert resid is not None and gate is not None
        assert resid.dtype == torch.float32 and resid.shape == (M, N) and resid.stride(1) == 1
        assert gate.dtype == bf16 and gate.dim() == 2 and gate.shape[1] == N and gate.stride(1) == 1
        assert gate.shape[0] * rows_per_gate >= M
        ldr, gs = resid.stride(0), gate.stride(0)
    if epilogue in (EPI_SWIGLU_DUAL, EPI_SWIGLU_BWD):
        assert aux is not None and resid is None and aux.dtype == bf16 and aux.stride(1) == 1
        assert aux.shape == (M, N if epilogue == EPI_SWIGLU_DUAL else 2 * N)
        resid, ldr = aux, aux.stride(0)
    probe = gemm_probe
    ev = probe.before() if probe is not None else None
    call("deco_gemm_bf16", ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K, epilogue,
         ptr(bias), ptr(resid), ldr, ptr(gate), gs, rows_per_gate, tile_n, _st(a))
    if probe is not None:
        probe.after(ev, 2.0 * M * N * K)
    return out


def gemm_stream_parts(N: int, K: int) -> int:
    """Number of per-row partial sums the FE_STREAM epilogue of an [*, K] x [K, N] GEMM writes (its column tiles)."""
    return _lib.load().deco_gemm_stream_parts(int(N), int(K))


def gemm_stream(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor,
                resid: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None, rows_per_image: int = 1,
                next_w: Optional[torch.Tensor] = None, next_scale: Optional[torch.Tensor] = None,
                xg: Optional[torch.Tensor] = None, ssq: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Residual-stream GEMM (csrc/gemm_fused.cu FE_STREAM): out = [resid + gate *] (a @ w.T + bias) in fp32 (out may be
    resid), plus ssq[p, row] partial sums of squares of out and xg = bf16(out * next_w * (1 + next_scale))."""
    _cuda(a, w, bias, out, resid, gate, next_w, next_scale, xg, ssq)
    assert a.dtype == bf16 and w.dtype == bf16 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and out.dtype == torch.float32 and out.shape == (M, N) and out.stride(1) == 1
    if resid is not None:
        assert resid.dtype == torch.float32 and resid.shape == (M, N) and resid.stride(1) == 1
    if gate is not None:
        assert gate.dtype == bf16 and gate.shape[1] == N and gate.stride(1) == 1 and gate.shape[0] * rows_per_image >= M
    if next_w is not None:
        assert next_w.dtype == torch.float32 and next_w.numel() == N and next_w.is_contiguous()
        assert next_scale.dtype == bf16 and next_scale.shape[1] == N and next_scale.stride(1) == 1
        assert next_scale.shape[0] * rows_per_image >= M
        assert xg is not None and xg.dtype == bf16 and xg.shape == (M, N) and xg.stride(1) == 1
    if ssq is not None:
        assert ssq.dtype == torch.float32 and ssq.is_contiguous() and ssq.shape == (gemm_stream_parts(N, K), M)
    probe = gemm_probe
    ev = probe.before() if probe is not None else None
    call("deco_gemm_stream", ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, ptr(bias), ptr(resid),
         resid.stride(0) if resid is not None else 0, ptr(out), out.stride(0), ptr(gate),
         gate.stride(0) if gate is not None else 0, rows_per_image, ptr(next_w), ptr(next_scale),
         next_scale.stride(0) if next_scale is not None else 0, ptr(xg), xg.stride(0) if xg is not None else 0,
         ptr(ssq), _st(a))
    if probe is not None:
        probe.after(ev, 2.0 * M * N * K)
    return out


def gemm_norm_qkv(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, rows_per_image: int, heads: int, head_dim: int,
                  seg_w=(None, None, None), rope_mask: int = 0, rope: Optional[torch.Tensor] = None,
                  rope_tokens_per_row: int = 0, ssq: Optional[torch.Tensor] = None, norm_hidden: int = 0, shw: Optional[torch.Tensor] = None,
                  norm_eps: float = 1e-6, head_eps: float = 1e-6, out_head_pitch: int = 0) -> torch.Tensor:
    """out = headnorm/rope(rstd * (a @ w.T) + shw[row // rows_per_image]) (csrc/gemm_fused.cu FE_NORM_QKV).
    out_head_pitch: output columns per head (0 = head_dim); 80 for head_dim 72 writes zero-padded, sector-aligned heads
    (out is then [M, N / head_dim * 80])."""
    _cuda(a, w, out, rope, ssq, shw, *seg_w)
    assert a.dtype == bf16 and w.dtype == bf16 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    n_out = N // head_dim * (out_head_pitch or head_dim)
    assert w.shape[1] == K and out.dtype == bf16 and out.shape == (M, n_out) and out.stride(1) == 1
    if ssq is not None:
        assert ssq.dtype == torch.float32 and ssq.is_contiguous() and ssq.dim() == 2 and ssq.shape[1] == M
    if shw is not None:
        assert shw.dtype == torch.float32 and shw.shape[1] == N and shw.stride(1) == 1 and shw.shape[0] * rows_per_image >= M
    if rope is not None:
        assert rope.dtype == torch.float32 and rope.is_contiguous() and rope.shape == (rows_per_image, head_dim // 2, 2)
    sw = list(seg_w) + [None] * (3 - len(seg_w))
    for t in sw:
        assert t is None or (t.dtype == torch.float32 and t.numel() == head_dim and t.is_contiguous())
    probe = gemm_probe
    ev = probe.before() if probe is not None else None
    call("deco_gemm_norm_qkv", ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K, rows_per_image,
         ptr(ssq), ssq.shape[0] if ssq is not None else 0, norm_hidden, float(norm_eps), ptr(shw),
         shw.stride(0) if shw is not None else 0, heads, head_dim, ptr(sw[0]), ptr(sw[1]), ptr(sw[2]), rope_mask,
         ptr(rope), int(rope_tokens_per_row), float(head_eps), int(out_head_pitch), _st(a))
    if probe is not None:
        probe.after(ev, 2.0 * M * N * K)
    return out


def gemm_norm_swiglu(a: torch.Tensor, w13: torch.Tensor, out: torch.Tensor, rows_per_image: int,
                     ssq: Optional[torch.Tensor] = None, norm_hidden: int = 0, shw: Optional[torch.Tensor] = None,
                     norm_eps: float = 1e-6) -> torch.Tensor:
    """out = silu(y_a) * y_b with y = rstd * (a @ w13.T) + shw[row // rows_per_image] on the interleaved columns."""
    _cuda(a, w13, out, ssq, shw)
    assert a.dtype == bf16 and w13.dtype == bf16 and a.stride(1) == 1 and w13.stride(1) == 1
    M, K = a.shape
    N = w13.shape[0]
    assert w13.shape[1] == K and out.dtype == bf16 and out.shape == (M, N // 2) and out.stride(1) == 1
    if ssq is not None:
        assert ssq.dtype == torch.float32 and ssq.is_contiguous() and ssq.dim() == 2 and ssq.shape[1] == M
    if shw is not None:
        assert shw.dtype == torch.float32 and shw.shape[1] == N and shw.stride(1) == 1 and shw.shape[0] * rows_per_image >= M
    probe = gemm_probe
    ev = probe.before() if probe is not None else None
    call("deco_gemm_norm_swiglu", ptr(a), a.stride(0), ptr(w13), w13.stride(0), ptr(out), out.stride(0), M, N, K,
         rows_per_image, ptr(ssq), ssq.shape[0] if ssq is not None else 0, norm_hidden, float(norm_eps), ptr(shw),
         shw.stride(0) if shw is not None else 0, _st(a))
    if probe is not None:
        probe.after(ev, 2.0 * M * N * K)
    return out


def patchify(x: torch.Tensor, p: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(x, out)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty((B * (H // p) * (W // p), Cc * p * p), dtype=bf16, device=x.device)
    assert out.dtype == bf16 and out.is_contiguous() and out.shape == (B * (H // p) * (W // p), Cc * p * p)
    call("deco_patchify", ptr(x), ptr(out), B, Cc, H, W, p, _st(x))
    return out


def timestep_freq(t: torch.Tensor, dim: int = 256, max_period: float = 10.0) -> torch.Tensor:
    _cuda(t)
    t = t.reshape(-1).to(torch.float32).contiguous()
    out = torch.empty((t.numel(), dim), dtype=bf16, device=t.device)
    call("deco_timestep_freq", ptr(t), ptr(out), t.numel(), dim, float(max_period), _st(t))
    return out


def cond_combine(temb: torch.Tensor, table: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    _cuda(temb, table, labels)
    assert temb.dtype == bf16 and temb.is_contiguous() and table.dtype == torch.float32 and table.is_contiguous()
    labels = labels.reshape(-1).to(torch.int64).contiguous()
    B, Hd = temb.shape
    assert labels.numel() == B and table.shape[1] == Hd
    out = torch.empty_like(temb)
    call("deco_cond_combine", ptr(temb), ptr(table), ptr(labels), ptr(out), B, Hd, table.shape[0], _st(temb))
    return out


def rmsnorm_modulate(x: torch.Tensor, weight: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor,
                     rows_per_mod: int, eps: float = 1e-6, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [M,H] fp32 or bf16; shift/scale: bf16 views [M/rows_per_mod, H] sharing one row stride; out bf16."""
    _cuda(x, weight, shift, scale)
    assert x.dtype in (bf16, torch.float32) and x.is_contiguous() and weight.dtype == torch.float32
    M, Hd = x.shape
    assert shift.stride(0) == scale.stride(0) and shift.stride(1) == 1 and scale.stride(1) == 1
    if out is None:
        out = torch.empty(x.shape, dtype=bf16, device=x.device)
    call("deco_rmsnorm_modulate", ptr(x), int(x.dtype == torch.float32), ptr(weight), ptr(shift), ptr(scale), shift.stride(0), rows_per_mod,
         ptr(out), M, Hd, float(eps), _st(x))
    return out


def qknorm_rope_(qkv: torch.Tensor, q_weight: torch.Tensor, k_weight: torch.Tensor, rope: torch.Tensor,
                 heads: int, head_dim: int, L: int, eps: float = 1e-6) -> torch.Tensor:
    """In place on qkv [M, 3*heads*head_dim]; rope fp32 [L, head_dim/2, 2] (cos, sin)."""
    _cuda(qkv, q_weight, k_weight, rope)
    assert qkv.dtype == bf16 and qkv.is_contiguous() and qkv.shape[1] == 3 * heads * head_dim
    assert rope.dtype == torch.float32 and rope.is_contiguous() and rope.shape == (L, head_dim // 2, 2)
    call("deco_qknorm_rope", ptr(qkv), ptr(q_weight), ptr(k_weight), ptr(rope), qkv.shape[0], heads, head_dim, L,
         float(eps), _st(qkv))
    return qkv


def headnorm_rope_(buf: torch.Tensor, col0: int, w0: torch.Tensor, heads: int, head_dim: int, L: int,
                   col1: Optional[int] = None, w1: Optional[torch.Tensor] = None,
                   rope: Optional[torch.Tensor] = None, eps: float = 1e-6) -> torch.Tensor:
    """In place per-head RMSNorm (+ optional RoPE) on one or two column segments of buf [M, row_stride]
    (the t2i layouts: qkv_x, kv_y, text-refine qkv).  rope fp32 [L, head_dim/2, 2] or None."""
    _cuda(buf, w0, w1, rope)
    assert buf.dtype == bf16 and buf.dim() == 2 and buf.stride(1) == 1
    if rope is not None:
        assert rope.dtype == torch.float32 and rope.is_contiguous() and rope.shape == (L, head_dim // 2, 2)
    nseg = 1 if col1 is None else 2
    call("deco_headnorm_rope", ptr(buf), buf.stride(0), nseg, col0, 0 if col1 is None else col1, ptr(w0), ptr(w1),
         ptr(rope), buf.shape[0], heads, head_dim, L, float(eps), _st(buf))
    return buf


def qknorm_rope_to(raw: torch.Tensor, dst: torch.Tensor, q_weight: torch.Tensor, k_weight: torch.Tensor, rope: torch.Tensor,
                   heads: int, head_dim: int, L: int, eps: float = 1e-6) -> torch.Tensor:
    """dst[:, :2H] = q/k-norm + RoPE of raw[:, :2H] (raw, dst: [M, 3H] bf16, same layout); the v columns are not touched."""
    _cuda(raw, dst, q_weight, k_weight, rope)
    assert raw.dtype == bf16 and dst.dtype == bf16 and raw.is_contiguous() and dst.is_contiguous() and raw.shape == dst.shape
    assert raw.shape[1] == 3 * heads * head_dim
    call("deco_headnorm_rope_to", ptr(raw), ptr(dst), raw.stride(0), 2, 0, heads * head_dim, ptr(q_weight), ptr(k_weight),
         ptr(rope), raw.shape[0], heads, head_dim, L, float(eps), _st(raw))
    return dst


def rmsnorm_addpos(x: torch.Tensor, weight: torch.Tensor, pos: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """out[m] = weight * rms(x[m]) + pos[m % T]; x fp32 [M,H], pos fp32 [T,H]; fp32 output."""
    _cuda(x, weight, pos)
    assert x.dtype == torch.float32 and x.is_contiguous() and pos.dtype == torch.float32 and pos.is_contiguous()
    assert weight.dtype == torch.float32 and pos.shape[1] == x.shape[1] and x.shape[0] % pos.shape[0] == 0
    out = torch.empty_like(x)
    call("deco_rmsnorm_addpos", ptr(x), ptr(weight), ptr(pos), pos.shape[0], ptr(out), x.shape[0], x.shape[1],
         float(eps), _st(x))
    return out


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() % 8 == 0
    out = torch.empty(x.shape, dtype=bf16, device=x.device)
    call("deco_cast_f32_bf16", ptr(x), ptr(out), x.numel(), _st(x))
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, heads: int, head_dim: int,
              k2: Optional[torch.Tensor] = None, v2: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, head_pitch: int = 0, head_pitch2: int = 0) -> torch.Tensor:
    """q [B*Lq, heads*pitch] view, k/v [B*Lk, heads*pitch] views (row strides free), optional second KV segment.
    head_pitch: elements between consecutive heads of a q / k / v row (0 = head_dim, dense heads); head_pitch2: the same
    for k2 / v2.  Returns out [B*Lq, heads*d] (dense)."""
    _cuda(q, k, v, k2, v2)
    Hd = heads * head_dim
    Hp = heads * (head_pitch or head_dim)
    Lq, Lk = q.shape[0] // B, k.shape[0] // B
    assert q.dtype == bf16 and q.shape[1] == Hp and k.shape[1] == Hp and v.shape == k.shape
    assert k.stride(0) == v.stride(0)
    if out is None:
        out = torch.empty((q.shape[0], Hd), dtype=bf16, device=q.device)
    Lk2, s2 = 0, 0
    if k2 is not None:
        Lk2, s2 = k2.shape[0] // B, k2.stride(0)
        assert v2 is not None and v2.stride(0) == s2
    if head_pitch or head_pitch2:
        call("deco_attention_fwd_pitched", ptr(q), q.stride(0), head_pitch, ptr(k), ptr(v), k.stride(0), head_pitch, Lk,
             ptr(k2), ptr(v2), s2, head_pitch2, Lk2, ptr(out), out.stride(0), B, heads, Lq, head_dim,
             float(head_dim) ** -0.5, _st(q))
        return out
    call("deco_attention_fwd", ptr(q), q.stride(0), ptr(k), ptr(v), k.stride(0), Lk, ptr(k2), ptr(v2), s2, Lk2,
         ptr(out), out.stride(0), B, heads, Lq, head_dim, float(head_dim) ** -0.5, _st(q))
    return out


def silu_add_rows(x: torch.Tensor, row: torch.Tensor, rows_per: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(x, row)
    assert x.dtype in (bf16, torch.float32) and row.dtype == bf16 and x.is_contiguous() and row.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=bf16, device=x.device)
    assert out.dtype == bf16
    call("deco_silu_add_rows", ptr(x), int(x.dtype == torch.float32), ptr(row), ptr(out), x.shape[0], x.shape[1], rows_per, _st(x))
    return out


def pixel_decoder(x: torch.Tensor, ycond: torch.Tensor, blob: torch.Tensor, postab: torch.Tensor, patch: int,
                  hidden_x: int, num_res_blocks: int, out_dtype=bf16) -> torch.Tensor:
    _cuda(x, ycond, blob, postab)
    assert x.dtype == torch.float32 and x.is_contiguous() and ycond.dtype == bf16 and ycond.is_contiguous()
    B, Cc, H, W = x.shape
    assert Cc == 3, "pixel decoder is built for 3 image channels"
    assert blob.numel() * blob.element_size() == _lib.load().deco_decoder_blob_bytes(num_res_blocks)
    out = torch.empty((B, Cc, H, W), dtype=out_dtype, device=x.device)
    call("deco_pixel_decoder", ptr(x), ptr(ycond), ptr(blob), ptr(postab), ptr(out), int(out_dtype == bf16),
         B, H, W, patch, hidden_x, num_res_blocks, _st(x))
    return out


def pixel_decoder_tc(x: torch.Tensor, ysilu: torch.Tensor, blob: torch.Tensor, patch: int, hidden_x: int,
                     num_res_blocks: int, out_dtype=bf16) -> torch.Tensor:
    """tcgen05 pixel decoder (csrc/decoder_tc.cu): x fp32 [rows,3,H,W], ysilu = silu(cond_embed(s)) bf16 [rows*L, p*p*32]."""
    _cuda(x, ysilu, blob)
    assert x.dtype == torch.float32 and x.is_contiguous() and ysilu.dtype == bf16 and ysilu.is_contiguous()
    B, Cc, H, W = x.shape
    assert Cc == 3, "pixel decoder is built for 3 image channels"
    assert ysilu.shape == (B * (H // patch) * (W // patch), patch * patch * hidden_x)
    assert blob.numel() * blob.element_size() == _lib.load().deco_decoder_tc_blob_bytes(num_res_blocks)
    out = torch.empty((B, Cc, H, W), dtype=out_dtype, device=x.device)
    call("deco_pixel_decoder_tc", ptr(x), ptr(ysilu), ptr(blob), ptr(out), int(out_dtype == bf16), B, H, W, patch, hidden_x,
         num_res_blocks, 0, None, 0.0, 0.0, 0.0, 0.0, None, None, None, None, None, _st(x))
    return out


def pixel_decoder_tc_step(x: torch.Tensor, ysilu: torch.Tensor, blob: torch.Tensor, patch: int, hidden_x: int,
                          num_res_blocks: int, dev: Optional[torch.Tensor] = None, g: float = 1.0, dt: float = 0.0,
                          c0: float = 1.0, c1: float = 0.0, p1: Optional[torch.Tensor] = None,
                          x_out: Optional[torch.Tensor] = None, pred_out: Optional[torch.Tensor] = None,
                          u8_out: Optional[torch.Tensor] = None, x_base: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Decoder of both CFG rows [uncond || cond] of the image state x [B,3,H,W] + guidance + multistep update in one kernel:
    x_out = x_base + dt (c0 pred + c1 p1), pred = u + g (c - u), x_base = x unless given (Heun corrector: the net sees
    x = x_hat, the update starts from x_base).  dev = device {g, dt, c0, c1, ...} for graph replays."""
    _cuda(x, ysilu, blob, dev, p1, x_out, pred_out, u8_out, x_base)
    assert x.dtype == torch.float32 and x.is_contiguous() and ysilu.dtype == bf16 and ysilu.is_contiguous()
    B, Cc, H, W = x.shape
    assert Cc == 3 and ysilu.shape == (2 * B * (H // patch) * (W // patch), patch * patch * hidden_x)
    assert blob.numel() * blob.element_size() == _lib.load().deco_decoder_tc_blob_bytes(num_res_blocks)
    if x_out is None:
        x_out = torch.empty_like(x)
    for t_, dt_ in ((p1, torch.float32), (x_out, torch.float32), (pred_out, torch.float32), (u8_out, torch.uint8),
                    (x_base, torch.float32)):
        assert t_ is None or (t_.dtype == dt_ and t_.is_contiguous() and t_.shape == x.shape)
    if dev is not None:
        assert dev.dtype == torch.float32 and dev.numel() >= 4 and dev.is_contiguous()
    call("deco_pixel_decoder_tc", ptr(x), ptr(ysilu), ptr(blob), None, 0, 2 * B, H, W, patch, hidden_x, num_res_blocks, 1,
         ptr(dev), float(g), float(dt), float(c0), float(c1), ptr(x_base), ptr(p1), ptr(x_out), ptr(pred_out), ptr(u8_out),
         _st(x))
    return x_out


def nerf_decoder(x: torch.Tensor, params, blob: torch.Tensor, patch: int, hidden_x: int, mlp_ratio: int,
                 out_dtype=bf16) -> torch.Tensor:
    """PixNerd hyper-network decoder (csrc/nerf_decoder.cu): x fp32 [B,3,H,W]; params = list of bf16 [B*L, 2*Hx*Hx*ratio]
    generated weights, one per NerfBlock."""
    import ctypes
    _cuda(x, blob, *params)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3
    B, _, H, W = x.shape
    M = B * (H // patch) * (W // patch)
    for t_ in params:
        assert t_.dtype == bf16 and t_.is_contiguous() and t_.shape == (M, 2 * hidden_x * hidden_x * mlp_ratio)
    assert blob.numel() * blob.element_size() == _lib.load().deco_nerf_decoder_blob_bytes(len(params))
    out = torch.empty((B, 3, H, W), dtype=out_dtype, device=x.device)
    ptrs = (ctypes.c_void_p * len(params))(*[t_.data_ptr() for t_ in params])
    call("deco_nerf_decoder", ptr(x), ctypes.cast(ptrs, ctypes.c_void_p), len(params), ptr(blob), ptr(out),
         int(out_dtype == bf16), B, H, W, patch, hidden_x, mlp_ratio, _st(x))
    return out


def cfg_step(x: torch.Tensor, net_out: torch.Tensor, g: float, dt: float, c0: float = 1.0,
             prev=(), coeffs=(), x_out: Optional[torch.Tensor] = None, want_pred: bool = False,
             want_v: bool = False, want_u8: bool = False):
    """x_out = x + dt * (c0 * cfg(net_out, g) + sum_j coeffs[j] * prev[j]).  Returns (x_out, pred, v, u8)."""
    _cuda(x, net_out)
    assert x.dtype == torch.float32 and x.is_contiguous() and net_out.is_contiguous()
    assert net_out.shape[0] == 2 * x.shape[0] and net_out.shape[1:] == x.shape[1:]
    assert net_out.dtype in (bf16, torch.float32)
    assert len(prev) == len(coeffs) <= 3
    n = x.numel()
    if x_out is None:
        x_out = torch.empty_like(x)
    pred = torch.empty_like(x) if want_pred else None
    v = torch.empty_like(x) if want_v else None
    u8 = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_u8 else None
    ps = [None, None, None]
    cs = [0.0, 0.0, 0.0]
    for j, (p, c) in enumerate(zip(prev, coeffs)):
        assert p.dtype == torch.float32 and p.is_contiguous() and p.shape == x.shape
        ps[j], cs[j] = p, float(c)
    call("deco_cfg_step", ptr(x), ptr(net_out), int(net_out.dtype == bf16), ptr(ps[0]), ptr(ps[1]), ptr(ps[2]),
         float(g), float(dt), float(c0), cs[0], cs[1], cs[2], ptr(x_out), ptr(pred), ptr(v), ptr(u8), n, _st(x))
    return x_out, pred, v, u8


def sampler_advance(table: torch.Tensor, counter: torch.Tensor, cur: torch.Tensor, t_out: torch.Tensor) -> None:
    """Head of a graphed sampling step (csrc/sampler.cu): cur = table[counter % rows], t_out[:] = that row's t, ++counter."""
    _cuda(table, counter, cur, t_out)
    assert table.dtype == torch.float32 and table.is_contiguous() and table.dim() == 2 and table.shape[1] == 8
    assert counter.dtype == torch.int32 and counter.numel() == 1 and cur.dtype == torch.float32 and cur.numel() == 8
    assert t_out.dtype == torch.float32 and t_out.is_contiguous()
    call("deco_sampler_advance", ptr(table), table.shape[0], ptr(counter), ptr(cur), ptr(t_out), t_out.numel(), _st(table))


def cfg_step_dev(x: torch.Tensor, net_out: torch.Tensor, cur: torch.Tensor, x_out: torch.Tensor,
                 p1: Optional[torch.Tensor] = None, pred_out: Optional[torch.Tensor] = None,
                 u8_out: Optional[torch.Tensor] = None) -> None:
    """cfg_step with {g, dt, c0, c1, ...} read from the device vector `cur` (graph replays); x_out may be x, pred_out may
    be p1."""
    _cuda(x, net_out, cur, x_out, p1, pred_out, u8_out)
    assert x.dtype == torch.float32 and x.is_contiguous() and net_out.is_contiguous() and x_out.is_contiguous()
    assert net_out.shape[0] == 2 * x.shape[0] and net_out.dtype in (bf16, torch.float32)
    call("deco_cfg_step_dev", ptr(x), ptr(net_out), int(net_out.dtype == bf16), ptr(p1), None, None, ptr(cur),
         ptr(x_out), ptr(pred_out), None, ptr(u8_out), x.numel(), _st(x))


def cfg_step_ex(x: torch.Tensor, net_out: torch.Tensor, g: float = 1.0, dt: float = 0.0, c0: float = 1.0,
                prev=(), coeffs=(), dev: Optional[torch.Tensor] = None, xpred_den: float = 0.0,
                kd: float = 0.0, sden: float = 1.0, a_s: float = 0.0, a_n: float = 0.0,
                noise: Optional[torch.Tensor] = None, x_out: Optional[torch.Tensor] = None,
                pred_out: Optional[torch.Tensor] = None, want_pred: bool = False, want_v: bool = False,
                want_u8: bool = False, u8_out: Optional[torch.Tensor] = None):
    """Extended sampler update (csrc/sampler.cu, deco_cfg_step_ex): x-prediction nets (xpred_den > 0, EulerSamplerJiT) and
    the SDE step functions (score coefficient a_s, noise coefficient a_n).  dev: device vector {g, dt, c0, c1, c2, c3, t,
    xpred_den} for graph replays (then g .. c0 / xpred_den are ignored).  Returns (x_out, pred, v, u8)."""
    _cuda(x, net_out, noise, dev, x_out, pred_out, u8_out)
    assert x.dtype == torch.float32 and x.is_contiguous() and net_out.is_contiguous()
    assert net_out.shape[0] == 2 * x.shape[0] and net_out.shape[1:] == x.shape[1:]
    assert net_out.dtype in (bf16, torch.float32)
    assert len(prev) == len(coeffs) <= 3
    if noise is not None:
        assert noise.dtype == torch.float32 and noise.is_contiguous() and noise.shape == x.shape
    if x_out is None:
        x_out = torch.empty_like(x)
    pred = pred_out if pred_out is not None else (torch.empty_like(x) if want_pred else None)
    v = torch.empty_like(x) if want_v else None
    u8 = u8_out if u8_out is not None else (torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_u8 else None)
    ps = [None, None, None]
    cs = [0.0, 0.0, 0.0]
    for j, (p, c) in enumerate(zip(prev, coeffs)):
        assert p.dtype == torch.float32 and p.is_contiguous() and p.shape == x.shape
        ps[j], cs[j] = p, float(c)
    call("deco_cfg_step_ex", ptr(x), ptr(net_out), int(net_out.dtype == bf16), ptr(ps[0]), ptr(ps[1]), ptr(ps[2]), ptr(dev),
         float(g), float(dt), float(c0), cs[0], cs[1], cs[2], float(xpred_den), float(kd), float(sden), float(a_s),
         float(a_n), ptr(noise), ptr(x_out), ptr(pred), ptr(v), ptr(u8), x.numel(), _st(x))
    return x_out, pred, v, u8


def heun_sde_step(x: torch.Tensor, v: torch.Tensor, dt: float, a_s: float, a_n: float, kd: float = 0.0, sden: float = 1.0,
                  s_in: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                  net_out: Optional[torch.Tensor] = None, x_hat: Optional[torch.Tensor] = None, g: float = 1.0,
                  kdh: float = 0.0, sdenh: float = 1.0, want_v_avg: bool = False, want_u8: bool = False):
    """Heun predictor (net_out None) / corrector with an SDE step function (csrc/sampler.cu heun_sde_step_kernel).
    Returns (x_out, v_hat, s_hat, v_avg, u8); the middle three are None for a predictor."""
    _cuda(x, v, s_in, noise, net_out, x_hat)
    for t_ in (x, v, s_in, noise, x_hat):
        assert t_ is None or (t_.dtype == torch.float32 and t_.is_contiguous() and t_.shape == x.shape)
    corr = net_out is not None
    if corr:
        assert net_out.is_contiguous() and net_out.shape[0] == 2 * x.shape[0] and net_out.dtype in (bf16, torch.float32)
        assert x_hat is not None
    x_out = torch.empty_like(x)
    v_hat = torch.empty_like(x) if corr else None
    s_hat = torch.empty_like(x) if corr else None
    v_avg = torch.empty_like(x) if (corr and want_v_avg) else None
    u8 = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_u8 else None
    call("deco_heun_sde_step", ptr(x), ptr(v), ptr(s_in), ptr(net_out), int(corr and net_out.dtype == bf16), ptr(x_hat),
         ptr(noise), float(g), float(dt), float(kd), float(sden), float(kdh), float(sdenh), float(a_s), float(a_n), int(corr),
         ptr(x_out), ptr(v_hat), ptr(s_hat), ptr(v_avg), ptr(u8), x.numel(), _st(x))
    return x_out, v_hat, s_hat, v_avg, u8


def layernorm_modulate(x: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor, rows_per_mod: int, eps: float = 1e-6,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm (no affine) + modulate of the fp32 stream x [M, H] -> bf16 (dit_c2i_baseline.py:76-82); shift / scale: bf16
    views [M / rows_per_mod, H] sharing one row stride."""
    _cuda(x, shift, scale)
    assert x.dtype == torch.float32 and x.is_contiguous() and shift.dtype == bf16 and scale.dtype == bf16
    M, Hd = x.shape
    assert shift.stride(0) == scale.stride(0) and shift.stride(1) == 1 and scale.stride(1) == 1
    if out is None:
        out = torch.empty(x.shape, dtype=bf16, device=x.device)
    call("deco_layernorm_modulate", ptr(x), ptr(shift), ptr(scale), shift.stride(0), rows_per_mod, ptr(out), M, Hd,
         float(eps), _st(x))
    return out


def center_rows(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = x - x.mean(1, keepdim=True) for the fp32 stream x [M, H] (out may be x)."""
    _cuda(x, out)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    if out is None:
        out = torch.empty_like(x)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == x.shape
    call("deco_center_rows", ptr(x), ptr(out), x.shape[0], x.shape[1], _st(x))
    return out


def unpatchify(tok: torch.Tensor, B: int, C: int, H: int, W: int, p: int) -> torch.Tensor:
    """F.fold(kernel = stride = p) of bf16 tokens [B*L, C*p*p] -> [B, C, H, W] (dit_c2i_baseline.py:378)."""
    _cuda(tok)
    assert tok.dtype == bf16 and tok.is_contiguous() and tok.shape == (B * (H // p) * (W // p), C * p * p)
    out = torch.empty((B, C, H, W), dtype=bf16, device=tok.device)
    call("deco_unpatchify", ptr(tok), ptr(out), B, C, H, W, p, _st(tok))
    return out


def fp2uint8(x: torch.Tensor) -> torch.Tensor:
    _cuda(x)
    x = x.to(torch.float32).contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    call("deco_fp2uint8", ptr(x), ptr(out), x.numel(), _st(x))
    return out


_DCT_SCRATCH = {}


def _dct_scratch(device) -> torch.Tensor:
    """Zeroed scratch (deco_dct_scratch_doubles doubles) per (device, stream): the kernel leaves its first three words zero
    again (include/deco_b200.h)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _DCT_SCRATCH.get(key)
    if buf is None:
        buf = _DCT_SCRATCH[key] = torch.zeros(_lib.load().deco_dct_scratch_doubles(), dtype=torch.float64, device=device)
    return buf


def dct_fm_loss(out: torch.Tensor, v_t: torch.Tensor, freq_w: torch.Tensor, freq_loss_weight: float,
                want_loss: bool = True, want_grad: bool = False, upstream: Optional[torch.Tensor] = None):
    """Returns (losses fp32[3] = fm, freq, total | None, grad | None)."""
    _cuda(out, v_t, freq_w)
    assert out.dim() == 4 and out.shape[1] == 3 and out.shape == v_t.shape
    assert out.dtype in (bf16, torch.float32) and out.is_contiguous()
    assert v_t.dtype == torch.float32 and v_t.is_contiguous()
    assert freq_w.dtype == torch.float32 and freq_w.numel() == 192 and freq_w.is_contiguous()
    B, _, H, W = out.shape
    losses = torch.empty(3, dtype=torch.float32, device=out.device) if want_loss else None
    grad = torch.empty_like(out) if want_grad else None
    accum = _dct_scratch(out.device)
    if upstream is not None:
        upstream = upstream.reshape(()).to(torch.float32)
    call("deco_dct_fm_loss", ptr(out), int(out.dtype == bf16), ptr(v_t), ptr(freq_w), B, H, W,
         float(freq_loss_weight), ptr(losses), ptr(grad), ptr(upstream), ptr(accum), _st(out))
    return losses, grad


# ------------------------------------------------------------------------------------------------ training-step inputs
def train_timesteps(nt: torch.Tensor, u_uniform: torch.Tensor, u_select: torch.Tensor, timeshift: float, linear: bool):
    """t = time_shift(where(u_select <= 0.9, sigmoid(nt), u_uniform)) and, for the LinearScheduler, the per-image
    coefficients (alpha, sigma, dalpha, dsigma) [B, 4] (csrc/train_inputs.cu).  Returns (t, coef | None)."""
    _cuda(nt, u_uniform, u_select)
    for v in (nt, u_uniform, u_select):
        assert v.dtype == torch.float32 and v.is_contiguous() and v.shape == nt.shape and v.dim() == 1
    B = nt.numel()
    t = torch.empty_like(nt)
    coef = torch.empty((B, 4), dtype=torch.float32, device=nt.device) if linear else None
    call("deco_train_timesteps", ptr(nt), ptr(u_uniform), ptr(u_select), float(timeshift), int(linear), ptr(t), ptr(coef), B,
         _st(nt))
    return t, coef


def flow_pair(x: torch.Tensor, eps: torch.Tensor, coef: torch.Tensor):
    """(x_t, v_t) = (alpha x + sigma eps, dalpha x + dsigma eps) with coef [B, 4] = per-image (alpha, sigma, dalpha, dsigma)."""
    _cuda(x, eps, coef)
    assert x.dtype == torch.float32 and x.is_contiguous() and eps.dtype == torch.float32 and eps.is_contiguous()
    assert eps.shape == x.shape and coef.dtype == torch.float32 and coef.is_contiguous() and coef.shape == (x.shape[0], 4)
    x_t, v_t = torch.empty_like(x), torch.empty_like(x)
    call("deco_flow_pair", ptr(x), ptr(eps), ptr(coef), ptr(x_t), ptr(v_t), x.shape[0], x.numel() // x.shape[0], _st(x))
    return x_t, v_t


def label_dropout(cond: torch.Tensor, uncond: torch.Tensor, u: torch.Tensor, p: float) -> torch.Tensor:
    """out[i] = uncond[i] if u[i] < p else cond[i] (int64 labels; base/training.py:14-20)."""
    _cuda(cond, uncond, u)
    assert cond.dtype == torch.int64 and uncond.dtype == torch.int64 and cond.dim() == 1 and uncond.shape == cond.shape
    assert u.dtype == torch.float32 and u.shape == cond.shape
    cond, uncond, u = cond.contiguous(), uncond.contiguous(), u.contiguous()
    out = torch.empty_like(cond)
    call("deco_label_dropout", ptr(cond), ptr(uncond), ptr(u), float(p), ptr(out), cond.numel(), _st(cond))
    return out


# ------------------------------------------------------------------------------------------------ backward (training)
def transpose_cast(src: torch.Tensor, rows_pad: Optional[int] = None) -> torch.Tensor:
    """[R, C] (fp32 or bf16, row stride free) -> bf16 [C, Rp] with rows R..Rp zero-filled (Rp = R rounded up to 8):
    the K-major operands of a wgrad GEMM dW = dY^T . X."""
    _cuda(src)
    assert src.dim() == 2 and src.stride(1) == 1 and src.dtype in (bf16, torch.float32)
    R, Cc = src.shape
    Rp = (R + 7) // 8 * 8 if rows_pad is None else rows_pad
    out = torch.empty((Cc, Rp), dtype=bf16, device=src.device)
    call("deco_transpose_cast", ptr(src), int(src.dtype == torch.float32), src.stride(0), ptr(out), out.stride(0),
         R, Cc, Rp, _st(src))
    return out


def colsum_(out: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """out[c] += sum_r x[r, c]; out fp32 [N] (caller zeroes it)."""
    _cuda(out, x)
    assert x.dim() == 2 and x.stride(1) == 1 and out.dtype == torch.float32 and out.numel() == x.shape[1]
    call("deco_colsum", ptr(x), int(x.dtype == torch.float32), x.stride(0), ptr(out), x.shape[0], x.shape[1], _st(x))
    return out


def gate_residual(s: torch.Tensor, a: torch.Tensor, gate: torch.Tensor, L: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = s + gate[row // L] * a  (fp32 stream, bf16 branch output / gate); out may be s."""
    _cuda(s, a, gate, out)
    assert s.dtype == torch.float32 and s.is_contiguous() and a.dtype == bf16 and a.is_contiguous() and a.shape == s.shape
    assert gate.dtype == bf16 and gate.stride(1) == 1 and gate.shape[1] == s.shape[1]
    if out is None:
        out = torch.empty_like(s)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == s.shape
    call("deco_gate_residual", ptr(s), ptr(a), ptr(gate), gate.stride(0), ptr(out), L, s.shape[0], s.shape[1], _st(s))
    return out


def gate_residual_norm(s: torch.Tensor, a: torch.Tensor, gate: torch.Tensor, L: int, weight: torch.Tensor, shift: torch.Tensor,
                       scale: torch.Tensor, eps: float = 1e-6):
    """(s_out, h): s_out = s + gate[row // L] * a (fp32) and h = rms(s_out) * weight * (1 + scale) + shift (bf16) in one
    pass over the row -- gate_residual + rmsnorm_modulate without re-reading the stream."""
    _cuda(s, a, gate, weight, shift, scale)
    assert s.dtype == torch.float32 and s.is_contiguous() and a.dtype == bf16 and a.is_contiguous() and a.shape == s.shape
    assert gate.dtype == bf16 and gate.stride(1) == 1 and gate.shape[1] == s.shape[1] and weight.dtype == torch.float32
    assert shift.stride(0) == scale.stride(0) and shift.stride(1) == 1 and scale.stride(1) == 1
    s_out = torch.empty_like(s)
    h = torch.empty(s.shape, dtype=bf16, device=s.device)
    call("deco_gate_residual_norm", ptr(s), ptr(a), ptr(gate), gate.stride(0), ptr(s_out), ptr(weight), ptr(shift), ptr(scale),
         shift.stride(0), L, ptr(h), s.shape[0], s.shape[1], float(eps), _st(s))
    return s_out, h


def gate_bwd(ds: torch.Tensor, a: torch.Tensor, gate: torch.Tensor, dgate: torch.Tensor, L: int,
             dbias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """da = gate * ds (returned, bf16); dgate (fp32 view [B, H], row stride free) += sum_rows ds * a; dbias += sum da."""
    _cuda(ds, a, gate, dgate, dbias)
    assert ds.dtype == torch.float32 and ds.is_contiguous() and a.dtype == bf16 and a.is_contiguous() and a.shape == ds.shape
    assert gate.dtype == bf16 and gate.stride(1) == 1 and dgate.dtype == torch.float32 and dgate.stride(1) == 1
    da = torch.empty_like(a)
    ws = torch.zeros((ds.shape[0] // L, ds.shape[1]), dtype=torch.float32, device=ds.device) if dbias is not None else None
    call("deco_gate_bwd", ptr(ds), ptr(a), ptr(gate), gate.stride(0), ptr(da), ptr(dgate), dgate.stride(0), ptr(dbias),
         ptr(ws), L, ds.shape[0], ds.shape[1], _st(ds))
    return da


def silu_add_rows_bwd(dout: torch.Tensor, x: torch.Tensor, row: torch.Tensor, drow: torch.Tensor, L: int) -> torch.Tensor:
    _cuda(dout, x, row, drow)
    assert dout.dtype == bf16 and dout.is_contiguous() and x.dtype == torch.float32 and x.is_contiguous()
    assert row.dtype == bf16 and row.is_contiguous() and drow.dtype == torch.float32 and drow.is_contiguous()
    dx = torch.empty_like(x)
    call("deco_silu_add_rows_bwd", ptr(dout), ptr(x), ptr(row), ptr(dx), ptr(drow), L, x.shape[0], x.shape[1], _st(x))
    return dx


def swiglu_fwd(y13: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(y13)
    assert y13.dtype == bf16 and y13.is_contiguous()
    M, N2 = y13.shape
    if out is None:
        out = torch.empty((M, N2 // 2), dtype=bf16, device=y13.device)
    call("deco_swiglu_fwd", ptr(y13), ptr(out), M, N2 // 2, _st(y13))
    return out


def swiglu_bwd(y13: torch.Tensor, du: torch.Tensor) -> torch.Tensor:
    _cuda(y13, du)
    assert y13.dtype == bf16 and y13.is_contiguous() and du.dtype == bf16 and du.is_contiguous()
    M, N2 = y13.shape
    assert du.shape == (M, N2 // 2)
    dy = torch.empty_like(y13)
    call("deco_swiglu_bwd", ptr(y13), ptr(du), ptr(dy), M, N2 // 2, _st(y13))
    return dy


def rmsnorm_modulate_bwd_(ds: torch.Tensor, dh: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor,
                          dweight: torch.Tensor, dshift: torch.Tensor, dscale: torch.Tensor, L: int,
                          eps: float = 1e-6) -> torch.Tensor:
    """ds += d x of h = rms(x) * w * (1 + scale) + shift; dweight / dshift / dscale accumulated (fp32)."""
    _cuda(ds, dh, x, weight, scale, dweight, dshift, dscale)
    assert ds.dtype == torch.float32 and ds.is_contiguous() and x.dtype == torch.float32 and x.is_contiguous()
    assert dh.dtype == bf16 and dh.is_contiguous() and dh.shape == x.shape == ds.shape
    assert scale.dtype == bf16 and scale.stride(1) == 1
    assert dshift.dtype == torch.float32 and dshift.stride(1) == 1 and dshift.stride(0) == dscale.stride(0)
    ws = torch.empty(2 * x.shape[0], dtype=torch.float32, device=x.device)     # per-row (rstd, k2) of the first pass
    iws = torch.zeros((x.shape[0] // L, x.shape[1]), dtype=torch.float32, device=x.device)   # per-image partial sums
    call("deco_rmsnorm_modulate_bwd", ptr(dh), ptr(x), ptr(weight), ptr(scale), scale.stride(0), ptr(ds), ptr(dweight),
         ptr(dshift), ptr(dscale), dshift.stride(0), ptr(ws), ptr(iws), L, x.shape[0], x.shape[1], float(eps), _st(x))
    return ds


def rmsnorm_modulate_bwd_gate_(ds: torch.Tensor, dh: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, scale: torch.Tensor,
                               dweight: torch.Tensor, dshift: torch.Tensor, dscale: torch.Tensor, L: int,
                               a: torch.Tensor, gate: torch.Tensor, dgate: torch.Tensor, dbias: Optional[torch.Tensor] = None,
                               eps: float = 1e-6) -> torch.Tensor:
    """rmsnorm_modulate_bwd_ followed by gate_bwd on the updated ds, in one pass over it; returns da (bf16)."""
    _cuda(ds, dh, x, weight, scale, dweight, dshift, dscale, a, gate, dgate, dbias)
    assert ds.dtype == torch.float32 and ds.is_contiguous() and x.dtype == torch.float32 and x.is_contiguous()
    assert dh.dtype == bf16 and dh.is_contiguous() and dh.shape == x.shape == ds.shape
    assert scale.dtype == bf16 and scale.stride(1) == 1
    assert dshift.dtype == torch.float32 and dshift.stride(1) == 1 and dshift.stride(0) == dscale.stride(0)
    assert a.dtype == bf16 and a.is_contiguous() and a.shape == ds.shape
    assert gate.dtype == bf16 and gate.stride(1) == 1 and dgate.dtype == torch.float32 and dgate.stride(1) == 1
    nimg = x.shape[0] // L
    ws = torch.empty(2 * x.shape[0], dtype=torch.float32, device=x.device)
    iws = torch.zeros((2 if dbias is not None else 1, nimg, x.shape[1]), dtype=torch.float32, device=x.device)
    da = torch.empty_like(a)
    call("deco_rmsnorm_modulate_bwd_gate", ptr(dh), ptr(x), ptr(weight), ptr(scale), scale.stride(0), ptr(ds), ptr(dweight),
         ptr(dshift), ptr(dscale), dshift.stride(0), ptr(ws), ptr(iws[0]), L, x.shape[0], x.shape[1], float(eps),
         ptr(a), ptr(gate), gate.stride(0), ptr(da), ptr(dgate), dgate.stride(0), ptr(dbias),
         ptr(iws[1]) if dbias is not None else None, _st(x))
    return da


def headnorm_rope_bwd_(g: torch.Tensor, raw: torch.Tensor, col: int, weight: torch.Tensor, rope: Optional[torch.Tensor],
                       dweight: torch.Tensor, heads: int, head_dim: int, L: int, eps: float = 1e-6) -> torch.Tensor:
    _cuda(g, raw, weight, rope, dweight)
    assert g.dtype == bf16 and raw.dtype == bf16 and g.stride(1) == 1 and raw.stride(1) == 1 and g.shape[0] == raw.shape[0]
    assert dweight.dtype == torch.float32 and dweight.numel() == head_dim
    call("deco_headnorm_rope_bwd", ptr(g), g.stride(0), ptr(raw), raw.stride(0), col, ptr(weight), ptr(rope), ptr(dweight),
         g.shape[0], heads, head_dim, L, float(eps), _st(g))
    return g


def cond_combine_bwd_(dc: torch.Tensor, temb: torch.Tensor, table: torch.Tensor, labels: torch.Tensor,
                      dtemb: torch.Tensor, dtable: torch.Tensor) -> None:
    _cuda(dc, temb, table, labels, dtemb, dtable)
    assert dc.dtype == torch.float32 and dc.is_contiguous() and temb.dtype == bf16 and temb.is_contiguous()
    assert dtemb.dtype == torch.float32 and dtemb.is_contiguous() and dtable.dtype == torch.float32 and dtable.is_contiguous()
    labels = labels.reshape(-1).to(torch.int64).contiguous()
    B, Hd = temb.shape
    call("deco_cond_combine_bwd", ptr(dc), ptr(temb), ptr(table), ptr(labels), ptr(dtemb), ptr(dtable), B, Hd,
         table.shape[0], _st(dc))


def silu_bwd(z: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    _cuda(z, dy)
    assert z.dtype == bf16 and dy.dtype == bf16 and z.is_contiguous() and dy.is_contiguous() and z.shape == dy.shape
    dz = torch.empty_like(z)
    call("deco_silu_bwd", ptr(z), ptr(dy), ptr(dz), z.numel(), _st(z))
    return dz


def attention_lse(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, heads: int, head_dim: int):
    """attention() for the training forward: returns (out, lse2 fp32 [B*heads*Lq]) -- the softmax statistics the backward
    would otherwise rebuild."""
    _cuda(q, k, v)
    Hd = heads * head_dim
    Lq, Lk = q.shape[0] // B, k.shape[0] // B
    assert q.dtype == bf16 and q.shape[1] == Hd and k.shape[1] == Hd and v.shape == k.shape and k.stride(0) == v.stride(0)
    out = torch.empty((q.shape[0], Hd), dtype=bf16, device=q.device)
    lse = torch.empty(B * heads * Lq, dtype=torch.float32, device=q.device)
    call("deco_attention_fwd_lse", ptr(q), q.stride(0), ptr(k), ptr(v), k.stride(0), Lk, ptr(out), out.stride(0), ptr(lse),
         B, heads, Lq, head_dim, float(head_dim) ** -0.5, _st(q))
    return out, lse


def attention_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, dout: torch.Tensor,
                  dq: torch.Tensor, dk: torch.Tensor, dv: torch.Tensor, B: int, heads: int, head_dim: int,
                  lse: Optional[torch.Tensor] = None) -> None:
    """Gradients of deco_attention_fwd (one key segment) written into the strided views dq / dk / dv.  lse = the forward's
    statistics (attention_lse) or None (rebuilt here)."""
    _cuda(q, k, v, o, dout, dq, dk, dv, lse)
    Lq, Lk = q.shape[0] // B, k.shape[0] // B
    for t in (q, k, v, o, dout, dq, dk, dv):
        assert t.dtype == bf16 and t.stride(1) == 1
    assert k.stride(0) == v.stride(0) and dk.stride(0) == dv.stride(0)
    n = B * heads * Lq
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == n
    ws = torch.empty((2, n), dtype=torch.float32, device=q.device)
    aligned = all(t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0 for t in (q, k, v, o, dout, dq, dk, dv))
    if lse is not None and ATTN_BWD == "tc" and aligned:
        # tcgen05 / TMEM kernels (csrc/attention_bwd_tc.cu); the mma.sync pair below stays as the A/B variant
        # (DECO_B200_ATTN_BWD=legacy) and for callers without the forward's statistics
        call("deco_attention_bwd_tc", ptr(q), q.stride(0), ptr(k), ptr(v), k.stride(0), ptr(o), o.stride(0), ptr(dout),
             dout.stride(0), ptr(dq), dq.stride(0), ptr(dk), ptr(dv), dk.stride(0), ptr(lse), ptr(ws[1]), B, heads, Lq, Lk,
             head_dim, float(head_dim) ** -0.5, _st(q))
        return
    call("deco_attention_bwd", ptr(q), q.stride(0), ptr(k), ptr(v), k.stride(0), ptr(o), o.stride(0), ptr(dout),
         dout.stride(0), ptr(dq), dq.stride(0), ptr(dk), ptr(dv), dk.stride(0), ptr(lse if lse is not None else ws[0]),
         ptr(ws[1]), int(lse is not None), B, heads, Lq, Lk, head_dim, float(head_dim) ** -0.5, _st(q))


def pixel_decoder_bwd(x: torch.Tensor, ycond: torch.Tensor, dout: torch.Tensor, blob_f32: torch.Tensor,
                      postab: torch.Tensor, patch: int, hidden_x: int, num_res_blocks: int):
    """Returns (dycond bf16 like ycond, grads fp32 [blob floats + p*p*32])."""
    _cuda(x, ycond, dout, blob_f32, postab)
    assert x.dtype == torch.float32 and x.is_contiguous() and dout.dtype == torch.float32 and dout.is_contiguous()
    assert ycond.dtype == bf16 and ycond.is_contiguous() and blob_f32.dtype == torch.float32 and blob_f32.is_contiguous()
    B, Cc, H, W = x.shape
    n = _lib.load().deco_decoder_train_blob_floats(num_res_blocks)
    assert blob_f32.numel() == n and Cc == 3
    dy = torch.empty_like(ycond)
    grads = torch.zeros(n + patch * patch * hidden_x, dtype=torch.float32, device=x.device)
    call("deco_pixel_decoder_bwd", ptr(x), ptr(ycond), ptr(dout), ptr(blob_f32), ptr(postab), ptr(dy), ptr(grads),
         B, H, W, patch, hidden_x, num_res_blocks, _st(x))
    return dy, grads


def pixel_decoder_bwd_tc(x: torch.Tensor, ycond: torch.Tensor, dout: torch.Tensor, fwd_blob: torch.Tensor,
                         bwd_blob: torch.Tensor, postab: torch.Tensor, patch: int, hidden_x: int, num_res_blocks: int):
    """Tensor-core version of pixel_decoder_bwd (csrc/decoder_bwd_mma.cu); same outputs."""
    _cuda(x, ycond, dout, fwd_blob, bwd_blob, postab)
    assert x.dtype == torch.float32 and x.is_contiguous() and dout.dtype == torch.float32 and dout.is_contiguous()
    assert ycond.dtype == bf16 and ycond.is_contiguous() and x.shape[1] == 3
    lib = _lib.load()
    assert fwd_blob.numel() * fwd_blob.element_size() == lib.deco_decoder_blob_bytes(num_res_blocks)
    assert bwd_blob.numel() * bwd_blob.element_size() == lib.deco_decoder_bwd_blob_bytes(num_res_blocks)
    B, _, H, W = x.shape
    n = lib.deco_decoder_train_blob_floats(num_res_blocks)
    dy = torch.empty_like(ycond)
    grads = torch.zeros(n + patch * patch * hidden_x, dtype=torch.float32, device=x.device)
    call("deco_pixel_decoder_bwd_tc", ptr(x), ptr(ycond), ptr(dout), ptr(fwd_blob), ptr(bwd_blob), ptr(postab), ptr(dy),
         ptr(grads), B, H, W, patch, hidden_x, num_res_blocks, _st(x))
    return dy, grads


def gemm_f32_splitk(a: torch.Tensor, w: torch.Tensor, split_k: int = 0) -> torch.Tensor:
    """out [M, N] fp32 = a @ w.T (a [M, K], w [N, K] bf16) with the K loop split over CTAs (0 = automatic) and reduced with
    fp32 atomics: skinny problems (M = batch) with a very long K."""
    _cuda(a, w)
    assert a.dtype == bf16 and w.dtype == bf16 and a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    probe = gemm_probe
    ev = probe.before() if probe is not None else None
    call("deco_gemm_bf16_f32_splitk", ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K, split_k, _st(a))
    if probe is not None:
        probe.after(ev, 2.0 * M * N * K)
    return out


def gemm_tn(at: torch.Tensor, wt: torch.Tensor, tile_n: int = 0, split_k: int = 0,
            out: Optional[torch.Tensor] = None, deinterleave16: bool = False) -> torch.Tensor:
    """out [M, N] fp32 = at^T @ wt for at [K, M], wt [K, N] bf16 row-major (wgrad: dW = dY^T . X, K = tokens); no
    transposed copies -- the GEMM stages both operands MN-major.  `out` lets the caller allocate the result on another
    stream than the one the GEMM is launched on.  deinterleave16: the M rows are [16 x w1 | 16 x w3] interleaved
    (the SwiGLU weight); out [M, N] is then written as two stacked plain matrices (dW1 ; dW3) (M % 32 == 0)."""
    _cuda(at, wt)
    assert at.dtype == bf16 and wt.dtype == bf16 and at.dim() == 2 and wt.dim() == 2
    assert at.stride(1) == 1 and wt.stride(1) == 1 and at.shape[0] == wt.shape[0]
    K, M = at.shape
    N = wt.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=at.device)
    assert out.dtype == torch.float32 and out.shape == (M, N) and out.stride(1) == 1
    probe = gemm_probe
    ev = probe.before() if probe is not None else None
    call("deco_gemm_bf16_tn_deint16" if deinterleave16 else "deco_gemm_bf16_tn", ptr(at), at.stride(0), ptr(wt), wt.stride(0),
         ptr(out), out.stride(0), M, N, K, tile_n, split_k, _st(at))
    if probe is not None:
        probe.after(ev, 2.0 * M * N * K)
    return out
