"""B200-native text-to-image DeCo denoiser (DeCo-XXL, configs_t2i/sft_res512.yaml:45-56).

Drop-in for the ORIGINAL `src/models/transformer/dit_t2i_DeCo.py::PixNerDiT` (the checkout's live file of that name is
the UniFlow fork's reconstruction model; the original survives as CPython-3.10 bytecode).  Its encoder is
the logic of `src/models/transformer/dit_t2i_pixnerd.py` -- Attention :16-63 (image queries over [image || text]
keys, shared k_norm, RoPE on the image q/k only), FlattenDiTBlock :65-81, TextRefineAttention/Block :144-198,
forward :276-297 -- and its decoder is SimpleMLPAdaLN (`dit_c2i_DeCo.py:288-415`).  Same constructor arguments and
`state_dict` keys (members s_embedder, x_embedder, t_embedder, y_embedder, y_pos_embedding, blocks, dec_net,
text_refine_blocks), `forward(x, t, y)` with y = text-encoder states [B, T, txt_embed_dim].

Per forward:  t sinusoid -> 2 GEMMs -> silu -> ONE adaLN GEMM for all text + image blocks
  text:   y_embedder GEMM -> rmsnorm + y_pos_embedding (fp32 stream) -> num_text_blocks x [norm/mod -> QKV GEMM ->
          head-norm -> attention -> proj GEMM (+gate,+res) -> norm/mod -> W12 GEMM (SwiGLU) -> W3 GEMM (+gate,+res)]
  image:  patchify -> s_embedder GEMM -> num_encoder_blocks x [same, with kv_y GEMM on the text stream + k-norm and
          the attention kernel's second key/value segment]
  silu(t + s) -> cond_embed GEMM -> fused pixel decoder
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import ops
from .denoiser import (StreamState, _Embed, _NerfEmbedder, _PixelDecoder, _TimestepEmbedder, _Weight, fused_blocks,
                       interleave_w13, pack_decoder, pack_decoder_tc)
from . import denoiser as denoiser_mod

bf16 = torch.bfloat16


class _TextEmbed(nn.Module):
    """layers/patch_embed.py:6-22 with norm_layer=RMSNorm."""

    def __init__(self, in_chans: int, embed_dim: int):
        super().__init__()
        self.proj = nn.Linear(in_chans, embed_dim, bias=True)
        self.norm = _Weight(embed_dim)


class _SwiGLU12(nn.Module):
    """layers/swiglu.py:4-17 (fused w12, un-scaled hidden width)."""

    def __init__(self, dim: int, hidden_dim: int):
        super().__init__()
        self.w12 = nn.Linear(dim, hidden_dim * 2, bias=False)
        self.w3 = nn.Linear(hidden_dim, dim, bias=False)


class _JointAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.qkv_x = nn.Linear(dim, dim * 3, bias=False)
        self.kv_y = nn.Linear(dim, dim * 2, bias=False)
        self.q_norm = _Weight(dim // num_heads)
        self.k_norm = _Weight(dim // num_heads)
        self.proj = nn.Linear(dim, dim)


class _TextAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.q_norm = _Weight(dim // num_heads)
        self.k_norm = _Weight(dim // num_heads)
        self.proj = nn.Linear(dim, dim)


class _Block(nn.Module):
    def __init__(self, hidden_size: int, groups: int, joint: bool, mlp_ratio: int = 4):
        super().__init__()
        self.norm1 = _Weight(hidden_size)
        self.attn = _JointAttention(hidden_size, groups) if joint else _TextAttention(hidden_size, groups)
        self.norm2 = _Weight(hidden_size)
        self.mlp = _SwiGLU12(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.Linear(hidden_size, 6 * hidden_size, bias=True))


def rope_cos_sin_ex2d(head_dim: int, height: int, width: int, theta: float = 10000.0, scale: float = 1.0) -> torch.Tensor:
    """[L, head_dim/2, 2] (cos, sin) of precompute_freqs_cis_ex2d (layers/rope.py:22-37): x positions
    linspace(0, height*scale, width), y positions linspace(0, width*scale, height); pair 2k <-> x, 2k+1 <-> y."""
    x_pos = torch.linspace(0, height * scale, width)
    y_pos = torch.linspace(0, width * scale, height)
    y_pos, x_pos = torch.meshgrid(y_pos, x_pos, indexing="ij")
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 4)[: head_dim // 4].float() / head_dim))
    ang = torch.stack([torch.outer(x_pos.reshape(-1), freqs), torch.outer(y_pos.reshape(-1), freqs)], dim=-1)
    ang = ang.reshape(height * width, -1).float()
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()


class PixNerDiT(nn.Module):
    """Original text-to-image denoiser (constructor per the bytecode / dit_t2i_pixnerd.py:202-216)."""
    cuda_graph_safe = True   # no host-device synchronisation in the forward (samplers may capture it)

    def __init__(self, in_channels=4, num_groups=12, hidden_size=1152, decoder_hidden_size=64, num_encoder_blocks=18,
                 num_decoder_blocks=4, num_text_blocks=4, patch_size=2, txt_embed_dim=1024, txt_max_length=100,
                 weight_path=None, load_ema=False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = in_channels
        self.hidden_size = hidden_size
        self.num_groups = num_groups
        self.decoder_hidden_size = decoder_hidden_size
        self.num_encoder_blocks = num_encoder_blocks
        self.num_decoder_blocks = num_decoder_blocks
        self.num_blocks = num_encoder_blocks + num_decoder_blocks
        self.num_text_blocks = num_text_blocks
        self.patch_size = patch_size
        self.txt_embed_dim = txt_embed_dim
        self.txt_max_length = txt_max_length
        self.s_embedder = _Embed(in_channels * patch_size ** 2, hidden_size)
        self.x_embedder = _NerfEmbedder(in_channels, decoder_hidden_size, max_freqs=8)
        self.t_embedder = _TimestepEmbedder(hidden_size)
        self.y_embedder = _TextEmbed(txt_embed_dim, hidden_size)
        self.y_pos_embedding = nn.Parameter(torch.randn(1, txt_max_length, hidden_size), requires_grad=True)
        self.blocks = nn.ModuleList([_Block(hidden_size, num_groups, joint=True) for _ in range(num_encoder_blocks)])
        self.dec_net = _PixelDecoder(decoder_hidden_size, decoder_hidden_size, in_channels, hidden_size,
                                     num_decoder_blocks, patch_size)
        self.text_refine_blocks = nn.ModuleList([_Block(hidden_size, num_groups, joint=False)
                                                 for _ in range(num_text_blocks)])
        self.initialize_weights()
        self.precompute_pos: Dict[Tuple[int, int], torch.Tensor] = {}
        self.weight_path = weight_path
        self.load_ema = load_ema
        self._prep = None
        self._prep_key = None
        self.fused = os.environ.get("DECO_B200_FUSED", "1") != "0"   # image blocks on csrc/gemm_fused.cu

    def initialize_weights(self):
        """dit_t2i_pixnerd.py:258-270 (the decoder's zero-init lives in _PixelDecoder)."""
        w = self.s_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.s_embedder.proj.bias, 0)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)

    # -------------------------------------------------------------------------------------------- weight preparation
    def _weights_key(self, device):
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    @torch.no_grad()
    def prepare(self, device) -> dict:
        key = self._weights_key(device)
        if self._prep is not None and self._prep_key == key:
            return self._prep
        H, Hx, p = self.hidden_size, self.decoder_hidden_size, self.patch_size
        if Hx != 32 or p != 16 or self.in_channels != 3:
            raise NotImplementedError("the fused pixel decoder is built for in_channels=3, patch_size=16, "
                                      "decoder_hidden_size=32 (configs_t2i/sft_res512.yaml)")
        d = H // self.num_groups
        if d not in (64, 72):
            raise NotImplementedError(f"head_dim {d}: attention/head-norm kernels are built for 64 and 72")
        if H % 8 or H > 2048 or self.txt_embed_dim % 8:
            raise NotImplementedError("hidden_size must be a multiple of 8 and <= 2048, txt_embed_dim a multiple of 8")

        def W(t):
            return t.detach().to(device=device, dtype=bf16).contiguous()

        def Fv(t):
            return t.detach().to(device=device, dtype=torch.float32).contiguous()

        P = {}
        P["ws"], P["bs"] = W(self.s_embedder.proj.weight), Fv(self.s_embedder.proj.bias)
        P["wt0"], P["bt0"] = W(self.t_embedder.mlp[0].weight), Fv(self.t_embedder.mlp[0].bias)
        P["wt2"], P["bt2"] = W(self.t_embedder.mlp[2].weight), Fv(self.t_embedder.mlp[2].bias)
        P["wy"], P["by"], P["yn"] = W(self.y_embedder.proj.weight), Fv(self.y_embedder.proj.bias), Fv(self.y_embedder.norm.weight)
        P["ypos"] = Fv(self.y_pos_embedding[0])
        allb = list(self.text_refine_blocks) + list(self.blocks)
        P["wada"] = W(torch.cat([b.adaLN_modulation[0].weight for b in allb], 0))
        P["bada"] = Fv(torch.cat([b.adaLN_modulation[0].bias for b in allb], 0))
        P["zero_row"] = torch.zeros(1, H, dtype=torch.float32, device=device)

        def pack(b, joint):
            F_ = b.mlp.w3.weight.shape[1]
            w12 = W(b.mlp.w12.weight)
            d_ = dict(n1=Fv(b.norm1.weight), qn=Fv(b.attn.q_norm.weight), kn=Fv(b.attn.k_norm.weight),
                      wproj=W(b.attn.proj.weight), bproj=Fv(b.attn.proj.bias), n2=Fv(b.norm2.weight),
                      w13=interleave_w13(w12[:F_], w12[F_:], (F_ + 15) // 16 * 16), w2=W(b.mlp.w3.weight))
            assert F_ % 16 == 0, "SwiGLU width must be a multiple of 16"
            if joint:
                d_["wqkv"], d_["wkvy"] = W(b.attn.qkv_x.weight), W(b.attn.kv_y.weight)
            else:
                d_["wqkv"] = W(b.attn.qkv.weight)
            return d_

        P["text"] = [pack(b, False) for b in self.text_refine_blocks]
        P["blocks"] = [pack(b, True) for b in self.blocks]
        P["ffn"] = self.blocks[0].mlp.w3.weight.shape[1] if len(self.blocks) else 0
        P["wcond"], P["bcond"] = W(self.dec_net.cond_embed.weight), Fv(self.dec_net.cond_embed.bias)
        # NerfEmbedder.fetch_pos of the t2i model (dit_t2i_pixnerd.py:92-96): real part of the complex ex2d table
        tab = rope_cos_sin_ex2d(self.x_embedder.max_freqs ** 2 * 2, p, p)[..., 0]
        P["blob"], P["postab"] = pack_decoder(self.x_embedder.embedder[0], self.dec_net, self.in_channels, tab, device)
        P["blob_tc"] = pack_decoder_tc(self.x_embedder.embedder[0], self.dec_net, self.in_channels, tab, device)
        self._prep, self._prep_key = P, key
        return P

    def fetch_pos(self, height, width, device):
        key = (height, width)
        if key not in self.precompute_pos:
            self.precompute_pos[key] = rope_cos_sin_ex2d(self.hidden_size // self.num_groups, height, width)
        tab = self.precompute_pos[key]
        if tab.device != torch.device(device):
            tab = tab.to(device)
            self.precompute_pos[key] = tab
        return tab

    # -------------------------------------------------------------------------------------------- forward
    def _block(self, bp, m, s, rows, B, bufs, pos=None, ytxt=None, T=0):
        """One AdaLN block on the fp32 stream s [B*rows, H]; m = this block's [B, 6H] modulation."""
        H, heads = self.hidden_size, self.num_groups
        d = H // heads
        hbuf, qkv, obuf, ubuf = bufs
        sh1, sc1, g1, sh2, sc2, g2 = (m[:, j * H:(j + 1) * H] for j in range(6))
        ops.rmsnorm_modulate(s, bp["n1"], sh1, sc1, rows, out=hbuf)
        ops.gemm(hbuf, bp["wqkv"], None, ops.EPI_BIAS, out=qkv)
        ops.headnorm_rope_(qkv, 0, bp["qn"], heads, d, rows, col1=H, w1=bp["kn"], rope=pos)
        k2 = v2 = None
        if ytxt is not None:
            kvy = ops.gemm(ytxt, bp["wkvy"], None, ops.EPI_BIAS)                      # [B*T, 2H]
            ops.headnorm_rope_(kvy, 0, bp["kn"], heads, d, T)
            k2, v2 = kvy[:, :H], kvy[:, H:]
        ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d, k2=k2, v2=v2, out=obuf)
        ops.gemm(obuf, bp["wproj"], bp["bproj"], ops.EPI_GATE_RESIDUAL, out=s, resid=s, gate=g1, rows_per_gate=rows)
        ops.rmsnorm_modulate(s, bp["n2"], sh2, sc2, rows, out=hbuf)
        ops.gemm(hbuf, bp["w13"], None, ops.EPI_SWIGLU, out=ubuf)
        ops.gemm(ubuf, bp["w2"], None, ops.EPI_GATE_RESIDUAL, out=s, resid=s, gate=g2, rows_per_gate=rows)

    def _bufs(self, rows, P, dev):
        H = self.hidden_size
        return (torch.empty((rows, H), dtype=bf16, device=dev), torch.empty((rows, 3 * H), dtype=bf16, device=dev),
                torch.empty((rows, H), dtype=bf16, device=dev), torch.empty((rows, P["ffn"]), dtype=bf16, device=dev))

    def forward(self, x, t, y):
        """x [B,C,H,W], t [B] in [0,1], y [B, T, txt_embed_dim] text-encoder states -> velocity [B,C,H,W] (bf16)."""
        if not x.is_cuda:
            raise RuntimeError("deco_b200 t2i PixNerDiT runs on CUDA (sm_100a) only; there is no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
            raise NotImplementedError("the denoiser backward yields parameter gradients only (the reference never "
                                      "differentiates w.r.t. x_t or the frozen text encoder's states); detach x and y")
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            # training step: one autograd node with a hand-written backward (deco_b200/autograd.py::t2i_train_forward /
            # _backward: the DeCo denoiser's block and decoder backward, joint attention over one [image || text] key segment)
            from .autograd import denoiser_train_apply
            return denoiser_train_apply(self, x, t, y)
        B, Cc, Hh, Ww = x.shape
        p = self.patch_size
        assert Cc == self.in_channels and Hh % p == 0 and Ww % p == 0
        with torch.no_grad():
            P = self.prepare(x.device)
            x32 = x.detach().to(torch.float32).contiguous()
            s2 = self._encode(P, ops.patchify(x32, p), t, y, B, Hh, Ww)
            R = self.num_decoder_blocks
            if denoiser_mod.DECODER == "tc":
                ysilu = ops.gemm(s2, P["wcond"], P["bcond"], ops.EPI_BIAS_SILU)
                return ops.pixel_decoder_tc(x32, ysilu, P["blob_tc"], p, self.decoder_hidden_size, R)
            ycond = ops.gemm(s2, P["wcond"], P["bcond"], ops.EPI_BIAS)
            return ops.pixel_decoder(x32, ycond, P["blob"], P["postab"], p, self.decoder_hidden_size, R)

    supports_fused_cfg_step = True

    @torch.no_grad()
    def cfg_step(self, x, t2, cfg_condition, dev=None, g=1.0, dt=0.0, c0=1.0, c1=0.0, p1=None, x_out=None,
                 pred_out=None, u8_out=None, x_base=None):
        """CFG-batched sampler step with guidance + multistep update fused into the decoder epilogue (see
        deco_b200.denoiser.PixNerDiT.cfg_step): x fp32 [B,C,H,W], t2 [2B], cfg_condition [2B, T, txt_embed_dim] stacked
        [uncond || cond] (adam_sampling.py:101-117)."""
        if not x.is_cuda:
            raise RuntimeError("deco_b200 t2i PixNerDiT runs on CUDA (sm_100a) only; there is no CPU fallback")
        if denoiser_mod.DECODER != "tc":
            raise NotImplementedError("the fused sampler step needs the tcgen05 decoder (DECO_B200_DECODER=tc)")
        B, Cc, Hh, Ww = x.shape
        p = self.patch_size
        L = (Hh // p) * (Ww // p)
        P = self.prepare(x.device)
        x32 = x.detach().to(torch.float32).contiguous()
        xp = torch.empty((2 * B * L, Cc * p * p), dtype=bf16, device=x.device)
        ops.patchify(x32, p, out=xp[: B * L])
        ops.patchify(x32, p, out=xp[B * L:])
        s2 = self._encode(P, xp, t2, cfg_condition, 2 * B, Hh, Ww)
        ysilu = ops.gemm(s2, P["wcond"], P["bcond"], ops.EPI_BIAS_SILU)
        return ops.pixel_decoder_tc_step(x32, ysilu, P["blob_tc"], p, self.decoder_hidden_size, self.num_decoder_blocks,
                                         dev=dev, g=g, dt=dt, c0=c0, c1=c1, p1=p1, x_out=x_out, pred_out=pred_out,
                                         u8_out=u8_out, x_base=x_base)

    def _encode(self, P, xp, t, y, B, Hh, Ww):
        """Patch tokens xp [B*L, C*p*p] + text states -> decoder condition s2 [B*L, H] (dit_t2i_pixnerd.py:276-297)."""
        p, H = self.patch_size, self.hidden_size
        assert y.dim() == 3 and y.shape[0] == B and y.shape[2] == self.txt_embed_dim
        T = y.shape[1]
        assert T == self.txt_max_length, "y_pos_embedding is added without slicing: T must equal txt_max_length"
        L = (Hh // p) * (Ww // p)
        dev = xp.device
        pos = self.fetch_pos(Hh // p, Ww // p, dev)
        tfreq = ops.timestep_freq(t.reshape(-1).to(torch.float32), self.t_embedder.frequency_embedding_size)
        h1 = ops.gemm(tfreq, P["wt0"], P["bt0"], ops.EPI_BIAS_SILU)
        temb = ops.gemm(h1, P["wt2"], P["bt2"], ops.EPI_BIAS)                                   # [B, H]
        c = ops.cond_combine(temb, P["zero_row"], torch.zeros(B, dtype=torch.int64, device=dev))  # silu(t)
        nt, ni = len(P["text"]), len(P["blocks"])
        mod = ops.gemm(c, P["wada"], P["bada"], ops.EPI_BIAS) if nt + ni else None              # [B, (nt+ni)*6H]
        # ---- text path
        y16 = y.detach().reshape(B * T, self.txt_embed_dim).to(bf16).contiguous()
        yraw = ops.gemm(y16, P["wy"], P["by"], ops.EPI_BIAS_F32)
        ys = ops.rmsnorm_addpos(yraw, P["yn"], P["ypos"])                                       # fp32 [B*T, H]
        if nt:
            bufs = self._bufs(B * T, P, dev)
            for i, bp in enumerate(P["text"]):
                self._block(bp, mod[:, i * 6 * H:(i + 1) * 6 * H], ys, T, B, bufs)
        ytxt = ops.cast_bf16(ys)
        # ---- image path
        if ni and self.fused and H % 32 == 0:
            st = StreamState(B * L, H, P["ffn"], dev, heads=self.num_groups)
            s = fused_blocks(P["blocks"], mod, nt, st, xp, P["ws"], P["bs"], B, L, H, self.num_groups, pos,
                             Ww // p, ytxt=ytxt, T=T)
            ni = 0
        else:
            s = ops.gemm(xp, P["ws"], P["bs"], ops.EPI_BIAS_F32)
        if ni:
            bufs = self._bufs(B * L, P, dev)
            for i, bp in enumerate(P["blocks"]):
                self._block(bp, mod[:, (nt + i) * 6 * H:(nt + i + 1) * 6 * H], s, L, B, bufs, pos=pos, ytxt=ytxt, T=T)
        return ops.silu_add_rows(s, temb, L)
