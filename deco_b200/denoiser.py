"""B200-native class-conditional DeCo denoiser -- drop-in for the reference module.

Mirrors `src/models/transformer/dit_c2i_DeCo.py:417-536` (class PixNerDiT): same constructor arguments, the same
`state_dict` keys and shapes (checkpoint contract), the same `forward(x, t, y, s=None, mask=None)` and
`forward_sx` signatures.  The computation itself never touches a PyTorch operator: every step is one of the
hand-written sm_100a kernels behind include/deco_b200.h.

Per forward (B' = CFG rows, L tokens, H hidden):
   patchify -> s_embedder GEMM -> [t sinusoid -> 2 GEMMs] -> cond_combine -> ONE adaLN GEMM for all blocks
   per block: rmsnorm_modulate -> QKV GEMM -> qknorm_rope (in place) -> attention (strided, no transposes)
              -> proj GEMM (+gate, +residual) -> rmsnorm_modulate -> W1|W3 GEMM (SwiGLU epilogue) -> W2 GEMM (+gate, +res)
   silu(t + s) -> cond_embed GEMM -> fused pixel decoder (NerfEmbedder + AdaLN-MLP + fold)
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

bf16 = torch.bfloat16
COMPOSITE_SHIFT = os.environ.get("DECO_B200_COMPOSITE_SHIFT", "1") != "0"
# pixel decoder: "tc" = tcgen05 kernel (csrc/decoder_tc.cu), "legacy" = register-resident mma.sync kernel (csrc/decoder.cu,
# kept for A/B measurements and as the forward of the training path, whose backward consumes the pre-activation ycond)
DECODER = os.environ.get("DECO_B200_DECODER", "tc")
QKV_PITCH80 = os.environ.get("DECO_B200_QKV_PITCH80", "1") != "0"
# training loops: re-preparation of the bf16 / packed / transposed weight copies after every optimizer step is ~300 small
# copy / gather kernels -- captured ONCE into a CUDA graph and replayed (static buffers), instead of re-launched one by one
PREP_GRAPH = os.environ.get("DECO_B200_PREP_GRAPH", "1") != "0"


# ------------------------------------------------------------------------------------------------ parameter holders
class _Weight(nn.Module):
    """RMSNorm parameter holder (dit_c2i_DeCo.py:85-92)."""

    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _Embed(nn.Module):
    def __init__(self, in_chans: int, embed_dim: int):
        super().__init__()
        self.proj = nn.Linear(in_chans, embed_dim, bias=True)


class _TimestepEmbedder(nn.Module):
    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size


class _LabelEmbedder(nn.Module):
    def __init__(self, num_classes: int, hidden_size: int):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes, hidden_size)
        self.num_classes = num_classes


class _Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.q_norm = _Weight(self.head_dim)
        self.k_norm = _Weight(self.head_dim)
        self.proj = nn.Linear(dim, dim)


class _FeedForward(nn.Module):
    def __init__(self, dim: int, hidden_dim: int):
        super().__init__()
        hidden_dim = int(2 * hidden_dim / 3)
        self.w1 = nn.Linear(dim, hidden_dim, bias=False)
        self.w3 = nn.Linear(dim, hidden_dim, bias=False)
        self.w2 = nn.Linear(hidden_dim, dim, bias=False)


class _DiTBlock(nn.Module):
    def __init__(self, hidden_size: int, groups: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.norm1 = _Weight(hidden_size)
        self.attn = _Attention(hidden_size, groups)
        self.norm2 = _Weight(hidden_size)
        self.mlp = _FeedForward(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.Linear(hidden_size, 6 * hidden_size, bias=True))


class _NerfEmbedder(nn.Module):
    def __init__(self, in_channels: int, hidden_size_input: int, max_freqs: int):
        super().__init__()
        self.max_freqs = max_freqs
        self.embedder = nn.Sequential(nn.Linear(in_channels + max_freqs ** 2, hidden_size_input, bias=True))


class _ResBlock(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.in_ln = nn.LayerNorm(channels, eps=1e-6)
        self.mlp = nn.Sequential(nn.Linear(channels, channels, bias=True), nn.SiLU(),
                                 nn.Linear(channels, channels, bias=True))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(channels, 3 * channels, bias=True))


class _DecFinal(nn.Module):
    def __init__(self, model_channels: int, out_channels: int):
        super().__init__()
        self.linear = nn.Linear(model_channels, out_channels, bias=True)


class _PixelDecoder(nn.Module):
    """Parameter layout of SimpleMLPAdaLN (dit_c2i_DeCo.py:334-393), including its init."""

    def __init__(self, in_channels, model_channels, out_channels, z_channels, num_res_blocks, patch_size):
        super().__init__()
        self.cond_embed = nn.Linear(z_channels, patch_size ** 2 * model_channels)
        self.input_proj = nn.Linear(in_channels, model_channels)
        self.res_blocks = nn.ModuleList([_ResBlock(model_channels) for _ in range(num_res_blocks)])
        self.final_layer = _DecFinal(model_channels, out_channels)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        for blk in self.res_blocks:
            nn.init.constant_(blk.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(blk.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)


# ------------------------------------------------------------------------------------------------ host-side tables
def rope_cos_sin(head_dim: int, height: int, width: int, theta: float = 10000.0, scale: float = 16.0) -> torch.Tensor:
    """[L, head_dim/2, 2] (cos, sin) of the 2-D axial RoPE angles; pair 2k <-> x, 2k+1 <-> y
    (dit_c2i_DeCo.py:116-131).  Evaluated in fp32 on the host exactly like the reference's table."""
    x_pos = torch.linspace(0, scale, width)
    y_pos = torch.linspace(0, scale, height)
    y_pos, x_pos = torch.meshgrid(y_pos, x_pos, indexing="ij")
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 4)[: head_dim // 4].float() / head_dim))
    ang = torch.stack([torch.outer(x_pos.reshape(-1), freqs), torch.outer(y_pos.reshape(-1), freqs)], dim=-1)
    ang = ang.reshape(height * width, -1).float()
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()


def nerf_pos_table(patch_size: int, max_freqs: int) -> torch.Tensor:
    """[p*p, max_freqs^2] constant of NerfEmbedder.fetch_pos (dit_c2i_DeCo.py:221-236)."""
    pos = torch.linspace(0, 1, patch_size)
    pos_y, pos_x = torch.meshgrid(pos, pos, indexing="ij")
    pos_x, pos_y = pos_x.reshape(-1, 1, 1), pos_y.reshape(-1, 1, 1)
    freqs = torch.linspace(0, max_freqs, max_freqs)
    fx, fy = freqs[None, :, None], freqs[None, None, :]
    return (torch.cos(pos_x * fx * torch.pi) * torch.cos(pos_y * fy * torch.pi) * (1 + fx * fy) ** -1
            ).view(-1, max_freqs ** 2)


def _frag_index(n_out: int, permuted: bool, device):
    """(rows, cols) gather indices of the mma.m16n8k16 B-fragment order [n_tile, k_step, lane, 4] for W [n_out, 32]:
    lane = 4*g + t holds W[8j+g][k0], W[8j+g][k0+1], W[8j+g][k1], W[8j+g][k1+1] with (k0, k1) = (16s+2t, 16s+8+2t), or --
    for the adaLN layers whose A operand is the 16-byte condition chunk -- the permuted (8t+4s, 8t+4s+2).
    See csrc/decoder.cu."""
    key = (n_out, permuted, str(device))
    if key not in _FRAG_INDEX:
        j = torch.arange(n_out // 8).view(-1, 1, 1, 1)
        s = torch.arange(2).view(1, -1, 1, 1)
        lane = torch.arange(32).view(1, 1, -1, 1)
        e = torch.arange(4).view(1, 1, 1, -1)
        g, t = lane // 4, lane % 4
        half, lo = e // 2, e % 2
        k = (8 * t + 4 * s + 2 * half + lo) if permuted else (16 * s + 8 * half + 2 * t + lo)
        shape = (n_out // 8, 2, 32, 4)
        _FRAG_INDEX[key] = ((8 * j + g).expand(shape).contiguous().to(device), k.expand(shape).contiguous().to(device))
    return _FRAG_INDEX[key]


_FRAG_INDEX: dict = {}


def _frag(w: torch.Tensor, permuted: bool) -> torch.Tensor:
    """Pack W [n_out, 32] (bf16) into B-fragment order (gather on W's own device, no host round trip)."""
    rows, k = _frag_index(w.shape[0], permuted, w.device)
    return w[rows, k].contiguous()


def pack_decoder(embed_linear: nn.Linear, dn: "_PixelDecoder", C: int, pos_table: torch.Tensor, device):
    """Weights of NerfEmbedder / input_proj / ResBlocks / final layer in the layout csrc/decoder.cu expects.
    pos_table [p*p, max_freqs^2] is the model's constant positional input of the NerfEmbedder; it is folded with the
    embedder weight into postab [p*p, 32] (fp32).  Everything stays on `device`: a training loop re-packs after every
    optimizer step and must not synchronise with the host."""
    R = len(dn.res_blocks)

    def rb(t):  # bf16-rounded copy on the device
        return t.detach().to(device=device, dtype=torch.float32).to(bf16)

    def fv(t):
        return t.detach().to(device=device, dtype=torch.float32).reshape(-1)

    wx = rb(embed_linear.weight)                                     # [32, C + 64]
    tab = pos_table.to(device=device, dtype=torch.float32).to(bf16).float()          # Linear input cast
    postab = tab @ wx[:, C:].float().t() + fv(embed_linear.bias)     # [p*p, 32] fp32
    frags = [_frag(rb(dn.input_proj.weight), False)]
    vec = [wx[:, :C].float().reshape(-1), torch.zeros(96 - 32 * C, device=device), fv(dn.input_proj.bias)]
    for blk in dn.res_blocks:
        frags += [_frag(rb(blk.adaLN_modulation[1].weight), True), _frag(rb(blk.mlp[0].weight), False),
                  _frag(rb(blk.mlp[2].weight), False)]
        vec += [fv(blk.adaLN_modulation[1].bias), fv(blk.in_ln.weight), fv(blk.in_ln.bias), fv(blk.mlp[0].bias),
                fv(blk.mlp[2].bias)]
    wf = torch.zeros(8, 32, dtype=bf16, device=device)
    wf[:C] = rb(dn.final_layer.linear.weight)
    frags.append(_frag(wf, False))
    bf = torch.zeros(8, device=device)
    bf[:C] = fv(dn.final_layer.linear.bias)
    vec.append(bf)
    frag_bytes = torch.cat([f.reshape(-1) for f in frags]).view(torch.uint8)
    vec_bytes = torch.cat(vec).view(torch.uint8)
    blob = torch.cat([frag_bytes, vec_bytes]).contiguous()
    from . import _lib
    assert blob.numel() == _lib.load().deco_decoder_blob_bytes(R), (blob.numel(), R)
    return blob, postab.contiguous()


def _sw32_tile(mat: torch.Tensor) -> torch.Tensor:
    """bf16 [rows, cols] (cols a multiple of 16, rows of 8) -> the 32-byte-swizzled K-major operand image csrc/decoder_tc.cu
    keeps in shared memory (csrc/tcgen05.cuh::sw32_offset: [16-column chunk][row][32 B], 16-byte halves swapped in rows
    4..7 of every 8-row group); returned as a flat bf16 tensor of rows * cols elements."""
    rows, cols = mat.shape
    assert cols % 16 == 0 and rows % 8 == 0
    r = torch.arange(rows, device=mat.device).view(-1, 1)
    c = torch.arange(cols, device=mat.device).view(1, -1)
    off = (c >> 4) * rows * 32 + r * 32 + ((((c >> 3) & 1) ^ ((r >> 2) & 1)) << 4) + (c & 7) * 2
    out = torch.zeros(rows * cols, dtype=bf16, device=mat.device)
    out[(off // 2).reshape(-1)] = mat.to(bf16).reshape(-1)
    return out


def _bias_tile(bias: torch.Tensor, rows: int) -> torch.Tensor:
    """fp32 [n <= rows] -> B operand [rows x 16] of the bias MMA: column 0 = bf16(bias), column 1 = bf16(bias - hi)."""
    t = torch.zeros(rows, 16, dtype=torch.float32, device=bias.device)
    hi = bias.to(bf16).float()
    t[: bias.numel(), 0] = hi
    t[: bias.numel(), 1] = (bias - hi).to(bf16).float()
    return _sw32_tile(t.to(bf16))


@torch.no_grad()
def pack_decoder_tc(embed_linear: nn.Linear, dn: "_PixelDecoder", C: int, pos_table: torch.Tensor, device) -> torch.Tensor:
    """Weight image of the tcgen05 pixel decoder (csrc/decoder_tc.cu; the algebra is spelled out in its header): per res
    block the scale|gate rows of adaLN, W0 diag(g) / 2, the composite (W0 diag(b) Wscale + W0 Wshift) / 2, W2 and the
    (hi, lo) bias tiles; the final linear; then the fp32 tables T' = Win (Wpos table + bx) + bin and W' = Win Wrgb that
    replace NerfEmbedder + input_proj.  Matrices are the bf16-rounded weights the reference multiplies with under autocast;
    composites are formed in fp32 and rounded once."""
    def rb(t):
        return t.detach().to(device=device, dtype=torch.float32).to(bf16).float()

    def fv(t):
        return t.detach().to(device=device, dtype=torch.float32)

    H = dn.input_proj.weight.shape[0]
    assert H == 32 and C == 3
    ones = torch.zeros(128, 16, dtype=bf16, device=device)
    ones[:, :2] = 1
    parts = [_sw32_tile(ones)]
    for blk in dn.res_blocks:
        wa, ba = rb(blk.adaLN_modulation[1].weight), fv(blk.adaLN_modulation[1].bias)       # rows: shift | scale | gate
        wsh, wsc, wgt = wa[:H], wa[H:2 * H], wa[2 * H:]
        bsh, bsc, bgt = ba[:H], ba[H:2 * H], ba[2 * H:]
        gm, bt = fv(blk.in_ln.weight), fv(blk.in_ln.bias)
        w0, b0 = rb(blk.mlp[0].weight), fv(blk.mlp[0].bias)
        w2, b2 = rb(blk.mlp[2].weight), fv(blk.mlp[2].bias)
        w0g = 0.5 * w0 * gm.view(1, -1)
        c0 = 0.5 * ((w0 * bt.view(1, -1)) @ wsc + w0 @ wsh)
        bias0 = 0.5 * (b0 + w0 @ (bt * (1.0 + bsc)) + w0 @ bsh)
        parts += [_sw32_tile(torch.cat([wsc, wgt], 0)), _sw32_tile(w0g), _sw32_tile(c0), _sw32_tile(w2),
                  _bias_tile(torch.cat([1.0 + bsc, bgt]), 64), _bias_tile(bias0, 32), _bias_tile(b2, 32)]
    wf = torch.zeros(16, H, device=device)
    wf[:C] = rb(dn.final_layer.linear.weight)
    parts += [_sw32_tile(wf), _bias_tile(fv(dn.final_layer.linear.bias), 16),
              torch.zeros(256, dtype=bf16, device=device)]                                   # pad the final section to 2048 B
    wx = rb(embed_linear.weight)                                                             # [32, C + 64]
    win, b_in = rb(dn.input_proj.weight), fv(dn.input_proj.bias)
    tab = pos_table.to(device=device, dtype=torch.float32).to(bf16).float()
    x_pos = tab @ wx[:, C:].t() + fv(embed_linear.bias)                                      # [256, 32]
    tprime = torch.zeros(256, 36, device=device)
    tprime[:, :H] = x_pos @ win.t() + b_in
    wrgb = torch.zeros(H, 4, device=device)
    wrgb[:, :C] = win @ wx[:, :C]
    blob = torch.cat([torch.cat(parts).view(torch.uint8), tprime.reshape(-1).view(torch.uint8),
                      wrgb.reshape(-1).view(torch.uint8)]).contiguous()
    from . import _lib
    assert blob.numel() == _lib.load().deco_decoder_tc_blob_bytes(len(dn.res_blocks)), blob.numel()
    return blob


def interleave_w13(w1: torch.Tensor, w3: torch.Tensor, Fp: int) -> torch.Tensor:
    """Rows of w1 and w3 (bf16 [F, H]) zero-padded to Fp and interleaved in groups of 16, the layout the SwiGLU GEMM
    epilogue expects (one 32-column accumulator chunk holds 16 matching (a, b) pairs)."""
    F_, H = w1.shape
    a = torch.zeros(Fp, H, device=w1.device, dtype=w1.dtype)
    b = torch.zeros(Fp, H, device=w1.device, dtype=w1.dtype)
    a[:F_], b[:F_] = w1, w3
    return torch.stack([a.view(Fp // 16, 16, H), b.view(Fp // 16, 16, H)], dim=1).reshape(2 * Fp, H).contiguous()


class StreamState:
    """Working set of the fused block path: fp32 stream s, its pre-modulated bf16 copy xg for the NEXT norm, and the
    ping-pong partial sums of squares the FE_STREAM epilogues leave for the consumer GEMMs (csrc/gemm_fused.cu)."""

    def __init__(self, M: int, H: int, ffn: int, device, heads: int = 0):
        # q / k / v of head_dim-72 heads are written at a pitch of 80 columns (zero padded): every 16-column TMA box of the
        # attention operands is then one aligned 32-byte sector (1.36 -> 1.01 GB DRAM reads per launch at 512 rows, -14 % time)
        d = H // heads if heads else 0
        self.head_pitch = 80 if (d == 72 and QKV_PITCH80) else 0
        qkv_cols = 3 * heads * self.head_pitch if self.head_pitch else 3 * H
        self.s = torch.empty((M, H), dtype=torch.float32, device=device)
        self.xg = torch.empty((M, H), dtype=bf16, device=device)
        self._ssq = torch.empty((2, (H + 127) // 128, M), dtype=torch.float32, device=device)
        self.qkv = torch.empty((M, qkv_cols), dtype=bf16, device=device)
        self.o = torch.empty((M, H), dtype=bf16, device=device)
        self.u = torch.empty((M, ffn), dtype=bf16, device=device)

    def ssq(self, which: int, N: int, K: int) -> torch.Tensor:
        """Partial-sum buffer `which` (ping-pong), sized for its producer: the [*, K] x [K, N] FE_STREAM GEMM."""
        return self._ssq[which, :ops.gemm_stream_parts(N, K)]


def fused_blocks(blocks, mod: torch.Tensor, mod0: int, st: StreamState, a0: torch.Tensor, w0: torch.Tensor,
                 b0: torch.Tensor, B: int, L: int, H: int, heads: int, pos, wp: int, ytxt=None, T: int = 0,
                 shw_all: Optional[torch.Tensor] = None):
    """s = a0 @ w0.T + b0, then every AdaLN block of `blocks` on the fp32 stream, with no stand-alone norm / modulate /
    q-k-norm / RoPE pass: see csrc/gemm_fused.cu.  mod [B, *] is the batched adaLN output; block i uses the six H-wide
    column groups starting at (mod0 + i) * 6H (shift, scale, gate) x (attention, MLP) (dit_c2i_DeCo.py:207).
    ytxt (t2i): refined text stream bf16 [B*T, H] feeding kv_y (dit_t2i_pixnerd.py:47-49)."""
    d = H // heads
    nb = len(blocks)
    hp = getattr(st, "head_pitch", 0)           # columns per head in st.qkv (0 = dense)
    Hq = heads * hp if hp else H

    def sl(i, j):
        k = (mod0 + i) * 6 + j
        return mod[:, k * H:(k + 1) * H]

    s, xg = st.s, st.xg
    ffn = st.u.shape[1]
    # shift products sh @ W^T of every block ([B, H] x [H, N], functions of the modulation only): either one column slice
    # each of shw_all (ONE GEMM against composite weights, PixNerDiT.prepare) or two tiny GEMMs per block
    if shw_all is not None:
        S_ = 3 * H + 2 * ffn
        shw_qkv = [shw_all[:, i * S_:i * S_ + 3 * H] for i in range(nb)]
        shw_13 = [shw_all[:, i * S_ + 3 * H:(i + 1) * S_] for i in range(nb)]
    else:
        shw_qkv = [ops.gemm(sl(i, 0), bp["wqkv"], None, ops.EPI_BIAS_F32) for i, bp in enumerate(blocks)]
        shw_13 = [ops.gemm(sl(i, 3), bp["w13"], None, ops.EPI_BIAS_F32) for i, bp in enumerate(blocks)]
    q0 = st.ssq(0, H, a0.shape[1])      # statistics of the stream entering block 0 (producer: the embedding GEMM)
    q_attn = st.ssq(1, H, H)            # ... after the attention branch (producer: proj, K = H)
    q_mlp = st.ssq(0, H, ffn)           # ... after the MLP branch (producer: w2, K = ffn)
    ops.gemm_stream(a0, w0, b0, s, rows_per_image=L, next_w=blocks[0]["n1"], next_scale=sl(0, 1), xg=xg, ssq=q0)
    q_in = q0
    for i, bp in enumerate(blocks):
        ops.gemm_norm_qkv(xg, bp["wqkv"], st.qkv, L, heads, d, seg_w=(bp["qn"], bp["kn"], None), rope_mask=3, rope=pos,
                          rope_tokens_per_row=wp, ssq=q_in, norm_hidden=H, shw=shw_qkv[i], out_head_pitch=hp)
        k2 = v2 = None
        if ytxt is not None:
            kvy = torch.empty((ytxt.shape[0], 2 * H), dtype=bf16, device=s.device)
            ops.gemm_norm_qkv(ytxt, bp["wkvy"], kvy, T, heads, d, seg_w=(bp["kn"], None), rope_mask=0)
            k2, v2 = kvy[:, :H], kvy[:, H:]
        ops.attention(st.qkv[:, :Hq], st.qkv[:, Hq:2 * Hq], st.qkv[:, 2 * Hq:], B, heads, d, k2=k2, v2=v2, out=st.o,
                      head_pitch=hp)
        ops.gemm_stream(st.o, bp["wproj"], bp["bproj"], s, resid=s, gate=sl(i, 2), rows_per_image=L,
                        next_w=bp["n2"], next_scale=sl(i, 4), xg=xg, ssq=q_attn)
        ops.gemm_norm_swiglu(xg, bp["w13"], st.u, L, ssq=q_attn, norm_hidden=H, shw=shw_13[i])
        if i + 1 < nb:
            ops.gemm_stream(st.u, bp["w2"], None, s, resid=s, gate=sl(i, 5), rows_per_image=L,
                            next_w=blocks[i + 1]["n1"], next_scale=sl(i + 1, 1), xg=xg, ssq=q_mlp)
            q_in = q_mlp
        else:
            ops.gemm_stream(st.u, bp["w2"], None, s, resid=s, gate=sl(i, 5), rows_per_image=L)
    return s


@torch.no_grad()
def prepare_dit_blocks(blocks, H: int, device):
    """bf16 / fp32 device copies of the AdaLN DiT blocks' weights in the layouts the fused GEMMs expect (shared by the DeCo
    and the patch-linear baseline denoisers, whose FlattenDiTBlock is the same module: dit_c2i_DeCo.py:194-210,
    dit_c2i_baseline.py:194-210).  Returns (list of per-block dicts, padded FFN width)."""
    def W(t):
        return t.detach().to(device=device, dtype=bf16).contiguous()

    def Fv(t):
        return t.detach().to(device=device, dtype=torch.float32).contiguous()

    F = blocks[0].mlp.w1.weight.shape[0] if len(blocks) else 0
    Fp = (F + 15) // 16 * 16     # pad the FFN width (L/16: 2730 -> 2736) with zero rows / columns
    out = []
    for b in blocks:
        w1 = torch.zeros(Fp, H, device=device, dtype=bf16)
        w3 = torch.zeros(Fp, H, device=device, dtype=bf16)
        w1[:F], w3[:F] = W(b.mlp.w1.weight), W(b.mlp.w3.weight)
        # interleave 16 rows of w1 with 16 rows of w3 so that one 32-column accumulator chunk holds matching pairs
        w13 = torch.stack([w1.view(Fp // 16, 16, H), w3.view(Fp // 16, 16, H)], dim=1).reshape(2 * Fp, H).contiguous()
        w2 = torch.zeros(H, Fp, device=device, dtype=bf16)
        w2[:, :F] = W(b.mlp.w2.weight)
        out.append(dict(n1=Fv(b.norm1.weight), wqkv=W(b.attn.qkv.weight), qn=Fv(b.attn.q_norm.weight),
                        kn=Fv(b.attn.k_norm.weight), wproj=W(b.attn.proj.weight), bproj=Fv(b.attn.proj.bias),
                        n2=Fv(b.norm2.weight), w13=w13, w2=w2))
    return out, Fp


@torch.no_grad()
def composite_shift_weights(blocks, prepared, H: int, device):
    """(W . Wada_shift, W . bada_shift) for W = wqkv and w13 of every block, concatenated in block order: see
    PixNerDiT._composite_shift."""
    wc, bc = [], []
    for b, bp in zip(blocks, prepared):
        wa = b.adaLN_modulation[0].weight.detach().to(device=device, dtype=torch.float32)
        ba = b.adaLN_modulation[0].bias.detach().to(device=device, dtype=torch.float32)
        for wmat, lo in ((bp["wqkv"], 0), (bp["w13"], 3 * H)):
            wf = wmat.float()
            wc.append((wf @ wa[lo:lo + H]).to(bf16))
            bc.append(wf @ ba[lo:lo + H])
    return torch.cat(wc, 0).contiguous(), torch.cat(bc, 0).contiguous()


# ------------------------------------------------------------------------------------------------ the module
class PixNerDiT(nn.Module):
    """Drop-in for src/models/transformer/dit_c2i_DeCo.py::PixNerDiT (constructor :417-433, forward :488-510)."""
    cuda_graph_safe = True   # the inference forward makes no host-device synchronisation (samplers may capture it)

    def __init__(self, in_channels=4, num_groups=12, hidden_size=1152, hidden_size_x=64, nerf_mlpratio=4,
                 num_blocks=18, num_cond_blocks=4, patch_size=2, num_classes=1000, learn_sigma=True,
                 deep_supervision=0, weight_path=None, load_ema=False):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels
        self.hidden_size = hidden_size
        self.hidden_size_x = hidden_size_x
        self.num_groups = num_groups
        self.num_blocks = num_blocks
        self.num_cond_blocks = num_cond_blocks
        self.patch_size = patch_size
        self.x_embedder = _NerfEmbedder(in_channels, hidden_size_x, max_freqs=8)
        self.s_embedder = _Embed(in_channels * patch_size ** 2, hidden_size)
        self.t_embedder = _TimestepEmbedder(hidden_size)
        self.y_embedder = _LabelEmbedder(num_classes + 1, hidden_size)
        self.weight_path = weight_path
        self.load_ema = load_ema
        self.blocks = nn.ModuleList([_DiTBlock(hidden_size, num_groups) for _ in range(num_cond_blocks)])
        self.dec_net = _PixelDecoder(hidden_size_x, hidden_size_x, in_channels, hidden_size,
                                     num_blocks - num_cond_blocks, patch_size)
        self.initialize_weights()
        self.precompute_pos: Dict[Tuple[int, int], torch.Tensor] = {}
        self._prep = None
        self._prep_key = None
        # fused block path (csrc/gemm_fused.cu); False = one kernel per reference op (kept for A/B measurements)
        self.fused = os.environ.get("DECO_B200_FUSED", "1") != "0"

    def initialize_weights(self):
        """dit_c2i_DeCo.py:475-486."""
        w = self.s_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.s_embedder.proj.bias, 0)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)

    # -------------------------------------------------------------------------------------------- weight preparation
    def _weights_key(self, device):
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    @torch.no_grad()
    def prepare(self, device) -> dict:
        """bf16 copies / packed layouts of the fp32 master parameters; cached until a parameter changes.

        In a training loop the parameters change after every optimizer step: from the second re-preparation on, the whole
        sequence (prepare + autograd.prepare_train: casts, interleaves, transposes, decoder packing) is a CUDA graph replay
        into static buffers -- same kernels, no per-kernel launch cost (DECO_B200_PREP_GRAPH=0: plain re-launch)."""
        key = self._weights_key(device)
        if self._prep is not None and self._prep_key == key:
            return self._prep
        ptrs = (str(device),) + tuple(p.data_ptr() for p in self.parameters())
        g = self.__dict__.get("_prep_graph")
        if g is not None and g["ptrs"] == ptrs and self.training:
            g["graph"].replay()                       # the static buffers of g["P"] now hold the new weights' copies
            P = g["P"]
            for k in ("wshift", "bshift"):            # lazily built inference extras: stale now, rebuilt on demand
                P.pop(k, None)
            self._prep, self._prep_key = P, key
            return P
        want_graph = (PREP_GRAPH and self.training and torch.device(device).type == "cuda"
                      and not torch.cuda.is_current_stream_capturing())
        n_prev = self.__dict__.get("_prep_count", 0)
        self.__dict__["_prep_count"] = n_prev + 1
        if want_graph and n_prev >= 1:                # second re-preparation in training mode: this is a loop, capture it
            try:
                P = self._capture_prepare(device, ptrs)
                self._prep, self._prep_key = P, key
                return P
            except Exception as e:   # noqa: BLE001 -- the graph is an optimisation; the eager path computes the same copies
                import logging
                logging.getLogger(__name__).warning("CUDA-graph capture of the weight re-preparation failed (%s)", e)
                self.__dict__["_prep_graph"] = None
        P = self._prepare_eager(device)
        self._prep, self._prep_key = P, key
        return P

    def _capture_prepare(self, device, ptrs):
        from .autograd import prepare_train
        prepare_train(self, self._prepare_eager(device), device)   # warm-up outside the capture: constant tables, index caches
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            P = self._prepare_eager(device)
            prepare_train(self, P, device)
        graph.replay()
        self.__dict__["_prep_graph"] = dict(graph=graph, P=P, ptrs=ptrs)
        return P

    def _prepare_eager(self, device) -> dict:
        H, Hx, p = self.hidden_size, self.hidden_size_x, self.patch_size
        if Hx != 32 or p != 16 or self.in_channels != 3:
            raise NotImplementedError("the fused pixel decoder is built for in_channels=3, patch_size=16, "
                                      "hidden_size_x=32 (configs_c2i/DeCo_*.yaml)")
        d = H // self.num_groups
        if d not in (64, 72):
            raise NotImplementedError(f"head_dim {d}: attention/qknorm kernels are built for 64 and 72")

        def W(t):
            return t.detach().to(device=device, dtype=bf16).contiguous()

        def Fv(t):
            return t.detach().to(device=device, dtype=torch.float32).contiguous()

        P = {}
        P["ws"], P["bs"] = W(self.s_embedder.proj.weight), Fv(self.s_embedder.proj.bias)
        P["wt0"], P["bt0"] = W(self.t_embedder.mlp[0].weight), Fv(self.t_embedder.mlp[0].bias)
        P["wt2"], P["bt2"] = W(self.t_embedder.mlp[2].weight), Fv(self.t_embedder.mlp[2].bias)
        P["ytab"] = Fv(self.y_embedder.embedding_table.weight)
        P["wada"] = W(torch.cat([b.adaLN_modulation[0].weight for b in self.blocks], 0))
        P["bada"] = Fv(torch.cat([b.adaLN_modulation[0].bias for b in self.blocks], 0))
        P["blocks"], P["ffn_pad"] = prepare_dit_blocks(self.blocks, H, device)
        P["wcond"], P["bcond"] = W(self.dec_net.cond_embed.weight), Fv(self.dec_net.cond_embed.bias)
        P["blob"], P["postab"] = self._pack_decoder(device)
        P["blob_tc"] = pack_decoder_tc(self.x_embedder.embedder[0], self.dec_net, self.in_channels,
                                       self.precompute_pos[("nerf_tab", str(device))], device)
        return P

    @torch.no_grad()
    def _composite_shift(self, P, device):
        """shift_i . W^T = (c . Wada_shift^T + bada_shift) . W^T = c . (W . Wada_shift)^T + W . bada_shift: the 2 x nb
        batch-sized shift products of a forward collapse into ONE GEMM of c against these composite weights (at 8-GPU
        sharding the 56 tiny launches were 4 % of a step).  Built on first use by the fused inference path and cached with
        the prepared weights (a training loop that changes the weights every step never pays for it)."""
        return composite_shift_weights(self.blocks, P["blocks"], self.hidden_size, device)

    def _pack_decoder(self, device):
        key = ("nerf_tab", str(device))
        if key not in self.precompute_pos:      # constant table, cached on the device (no per-step host -> device copy)
            self.precompute_pos[key] = nerf_pos_table(self.patch_size, self.x_embedder.max_freqs).to(device)
        return pack_decoder(self.x_embedder.embedder[0], self.dec_net, self.in_channels, self.precompute_pos[key], device)

    def fetch_pos(self, height, width, device):
        """RoPE table cache per (h, w) (dit_c2i_DeCo.py:467-473); here as real (cos, sin)."""
        key = (height, width)
        if key not in self.precompute_pos:
            self.precompute_pos[key] = rope_cos_sin(self.hidden_size // self.num_groups, height, width)
        tab = self.precompute_pos[key]
        if tab.device != torch.device(device):
            tab = tab.to(device)
            self.precompute_pos[key] = tab
        return tab

    # -------------------------------------------------------------------------------------------- forward
    def _encode(self, P, xp, t, y, B, L, pos, wp):
        """Patch tokens -> DiT blocks -> decoder condition s (dit_c2i_DeCo.py:492-499)."""
        H, heads = self.hidden_size, self.num_groups
        d = H // heads
        nb = len(P["blocks"])
        tfreq = ops.timestep_freq(t, self.t_embedder.frequency_embedding_size)
        h1 = ops.gemm(tfreq, P["wt0"], P["bt0"], ops.EPI_BIAS_SILU)
        temb = ops.gemm(h1, P["wt2"], P["bt2"], ops.EPI_BIAS)                       # [B, H]
        c = ops.cond_combine(temb, P["ytab"], y)
        # residual stream in fp32 (the reference keeps it in bf16; fp32 costs ~4 % more HBM traffic per block and
        # halves the distance to the fp32 reference -- DESIGN.md "precision")
        if nb and self.fused and H % 32 == 0:
            mod = ops.gemm(c, P["wada"], P["bada"], ops.EPI_BIAS)                   # [B, nb*6H]
            st = StreamState(B * L, H, P["ffn_pad"], xp.device, heads=heads)
            if COMPOSITE_SHIFT and "wshift" not in P:
                P["wshift"], P["bshift"] = self._composite_shift(P, xp.device)
            shw_all = ops.gemm(c, P["wshift"], P["bshift"], ops.EPI_BIAS_F32) if "wshift" in P else None
            s = fused_blocks(P["blocks"], mod, 0, st, xp, P["ws"], P["bs"], B, L, H, heads, pos, wp, shw_all=shw_all)
            return ops.silu_add_rows(s, temb, L, out=st.o)
        s = ops.gemm(xp, P["ws"], P["bs"], ops.EPI_BIAS_F32)                        # [B*L, H] fp32
        if nb:
            mod = ops.gemm(c, P["wada"], P["bada"], ops.EPI_BIAS)                   # [B, nb*6H]
            hbuf = torch.empty((B * L, H), dtype=bf16, device=s.device)
            qkv = torch.empty((B * L, 3 * H), dtype=bf16, device=s.device)
            obuf = torch.empty((B * L, H), dtype=bf16, device=s.device)
            ubuf = torch.empty((B * L, P["ffn_pad"]), dtype=bf16, device=s.device)
        for i, bp in enumerate(P["blocks"]):
            m = mod[:, i * 6 * H:(i + 1) * 6 * H]
            sh1, sc1, g1, sh2, sc2, g2 = (m[:, j * H:(j + 1) * H] for j in range(6))
            ops.rmsnorm_modulate(s, bp["n1"], sh1, sc1, L, out=hbuf)
            ops.gemm(hbuf, bp["wqkv"], None, ops.EPI_BIAS, out=qkv)
            ops.qknorm_rope_(qkv, bp["qn"], bp["kn"], pos, heads, d, L)
            ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d, out=obuf)
            ops.gemm(obuf, bp["wproj"], bp["bproj"], ops.EPI_GATE_RESIDUAL, out=s, resid=s, gate=g1, rows_per_gate=L)
            ops.rmsnorm_modulate(s, bp["n2"], sh2, sc2, L, out=hbuf)
            ops.gemm(hbuf, bp["w13"], None, ops.EPI_SWIGLU, out=ubuf)
            ops.gemm(ubuf, bp["w2"], None, ops.EPI_GATE_RESIDUAL, out=s, resid=s, gate=g2, rows_per_gate=L)
        return ops.silu_add_rows(s, temb, L, out=(hbuf if nb else None))

    def _forward_impl(self, x, t, y, s=None, mask=None):
        if mask is not None:
            raise NotImplementedError("attention masks are not supported (the reference always passes mask=None)")
        if not x.is_cuda:
            raise RuntimeError("deco_b200.PixNerDiT runs on CUDA (sm_100a) only; there is no CPU fallback")
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("the denoiser backward yields parameter gradients only (the reference never "
                                      "differentiates w.r.t. x_t); detach x")
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            # training step (training_repa_DeCo.py:257 + loss.backward()): one autograd node, hand-written backward
            if s is not None:
                raise NotImplementedError("training with a precomputed s is not supported")
            from .autograd import denoiser_train_apply
            return denoiser_train_apply(self, x, t, y), None
        B, Cc, Hh, Ww = x.shape
        p, H = self.patch_size, self.hidden_size
        assert Cc == self.in_channels and Hh % p == 0 and Ww % p == 0
        L = (Hh // p) * (Ww // p)
        with torch.no_grad():
            P = self.prepare(x.device)
            x32 = x.detach().to(torch.float32).contiguous()
            if s is None:
                pos = self.fetch_pos(Hh // p, Ww // p, x.device)
                xp = ops.patchify(x32, p)
                s2 = self._encode(P, xp, t.reshape(-1).to(torch.float32), y.reshape(-1), B, L, pos, Ww // p)
            else:
                s2 = s.detach().reshape(B * L, H).to(bf16).contiguous()
            out = self._decode(P, x32, s2)
        return out, s2.view(B, L, H)

    def _decode(self, P, x32, s2, out_dtype=bf16):
        """cond_embed + per-pixel AdaLN-MLP decoder + fold (dit_c2i_DeCo.py:501-509)."""
        R = self.num_blocks - self.num_cond_blocks
        if DECODER == "tc":
            ysilu = ops.gemm(s2, P["wcond"], P["bcond"], ops.EPI_BIAS_SILU)      # silu(cond_embed(s)): the adaLN input
            return ops.pixel_decoder_tc(x32, ysilu, P["blob_tc"], self.patch_size, self.hidden_size_x, R, out_dtype=out_dtype)
        ycond = ops.gemm(s2, P["wcond"], P["bcond"], ops.EPI_BIAS)
        return ops.pixel_decoder(x32, ycond, P["blob"], P["postab"], self.patch_size, self.hidden_size_x, R, out_dtype=out_dtype)

    supports_fused_cfg_step = True

    @torch.no_grad()
    def cfg_step(self, x, t2, cfg_condition, dev=None, g=1.0, dt=0.0, c0=1.0, c1=0.0, p1=None, x_out=None,
                 pred_out=None, u8_out=None, x_base=None):
        """One CFG-batched sampler step with the update fused into the decoder epilogue (north_star (4)): evaluates the
        rows [uncond || cond] = net(cat[x, x], t2, cfg_condition) (sampling.py:89-97) WITHOUT materialising cat[x, x] or
        the bf16 network output, and writes x_out = x + dt (c0 pred + c1 p1), pred = u + g (c - u) (guidance.py:3-6,
        sampling.py:100-104, adam_sampling.py:104-117).  x fp32 [B,C,H,W]; t2 [2B]; cfg_condition [2B]; dev = device
        vector {g, dt, c0, c1} (graph replays) else the host scalars.  Returns x_out."""
        if not x.is_cuda:
            raise RuntimeError("deco_b200.PixNerDiT runs on CUDA (sm_100a) only; there is no CPU fallback")
        if DECODER != "tc":
            raise NotImplementedError("the fused sampler step needs the tcgen05 decoder (DECO_B200_DECODER=tc)")
        B, Cc, Hh, Ww = x.shape
        p, H = self.patch_size, self.hidden_size
        L = (Hh // p) * (Ww // p)
        P = self.prepare(x.device)
        x32 = x.detach().to(torch.float32).contiguous()
        pos = self.fetch_pos(Hh // p, Ww // p, x.device)
        xp = torch.empty((2 * B * L, Cc * p * p), dtype=bf16, device=x.device)
        ops.patchify(x32, p, out=xp[: B * L])          # both CFG halves see the same image: patchify it into each half
        ops.patchify(x32, p, out=xp[B * L:])
        s2 = self._encode(P, xp, t2.reshape(-1).to(torch.float32), cfg_condition.reshape(-1), 2 * B, L, pos, Ww // p)
        ysilu = ops.gemm(s2, P["wcond"], P["bcond"], ops.EPI_BIAS_SILU)
        return ops.pixel_decoder_tc_step(x32, ysilu, P["blob_tc"], p, self.hidden_size_x, self.num_blocks - self.num_cond_blocks,
                                         dev=dev, g=g, dt=dt, c0=c0, c1=c1, p1=p1, x_out=x_out, pred_out=pred_out,
                                         u8_out=u8_out, x_base=x_base)

    def forward(self, x, t, y, s=None, mask=None):
        """x [B,C,H,W], t [B] in [0,1], y [B] int64 (num_classes = null) -> velocity [B,C,H,W] (bf16, as the
        reference under autocast)."""
        return self._forward_impl(x, t, y, s, mask)[0]

    def forward_sx(self, x, t, y, s=None, mask=None):
        """dit_c2i_DeCo.py:512-536: also returns s as [B, H, sqrt(L), sqrt(L)]."""
        out, s2 = self._forward_impl(x, t, y, s, mask)
        if s2 is None:
            raise NotImplementedError("forward_sx is inference-only")
        B, L, H = s2.shape
        r = int(math.sqrt(L))
        return out, s2.reshape(B, r, r, H).permute(0, 3, 1, 2)
