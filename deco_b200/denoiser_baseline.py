"""B200-native patch-linear baseline denoiser -- drop-in for the reference's `FlattenDiT`.

Mirrors `src/models/transformer/dit_c2i_baseline.py:289-401` (class FlattenDiT; configs_c2i/Baseline_DiT.yaml,
Baseline_DiT_JiT.yaml): same constructor arguments, the same `state_dict` keys and shapes, `forward(x, t, y, masks=None)`
and `forward_sx`.  The AdaLN DiT blocks are the very module of the DeCo denoiser (dit_c2i_baseline.py:194-210 ==
dit_c2i_DeCo.py:194-210), so they run on the same fused tcgen05 GEMM / attention kernels (`denoiser.fused_blocks`); only the
head differs: instead of the per-pixel decoder, an AdaLN FinalLayer (LayerNorm without affine + modulate, :70-83) and a
Linear(H -> C p^2) per patch, folded back to the image (:376-378).

Per forward:  patchify -> [x_embedder GEMM -> blocks] (fused stream) ; t sinusoid -> 2 GEMMs -> cond_combine -> ONE adaLN
GEMM for all blocks + the final layer -> layernorm_modulate -> final GEMM (+bias) -> unpatchify.
With `EulerSamplerJiT` (sampling.py:109-188) the output is an x-prediction that the sampler's update kernel turns into a
velocity.  In `.train()` mode with grad enabled the forward is one autograd node with a hand-written backward
(`deco_b200.autograd.baseline_train_forward / _backward`: the DeCo denoiser's block backward between the patch-embedding
head and the FinalLayer / fold tail).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import ops
from .denoiser import (COMPOSITE_SHIFT, StreamState, _DiTBlock, _Embed, _LabelEmbedder, _TimestepEmbedder, bf16,
                       composite_shift_weights, fused_blocks, prepare_dit_blocks, rope_cos_sin)


class _FinalLayer(nn.Module):
    """Parameter layout of FinalLayer (dit_c2i_baseline.py:70-77); norm_final has no parameters."""

    def __init__(self, hidden_size: int, out_channels: int):
        super().__init__()
        self.linear = nn.Linear(hidden_size, out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.Linear(hidden_size, 2 * hidden_size, bias=True))


class FlattenDiT(nn.Module):
    """Drop-in for src/models/transformer/dit_c2i_baseline.py::FlattenDiT (constructor :290-303, forward :357-379)."""
    cuda_graph_safe = True

    def __init__(self, in_channels=4, num_groups=12, hidden_size=1152, num_blocks=18, patch_size=2, num_classes=1000,
                 learn_sigma=True, deep_supervision=0, weight_path=None, load_ema=False):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels
        self.hidden_size = hidden_size
        self.num_groups = num_groups
        self.num_blocks = num_blocks
        self.patch_size = patch_size
        self.x_embedder = _Embed(in_channels * patch_size ** 2, hidden_size)
        self.t_embedder = _TimestepEmbedder(hidden_size)
        self.y_embedder = _LabelEmbedder(num_classes + 1, hidden_size)
        self.final_layer = _FinalLayer(hidden_size, in_channels * patch_size ** 2)
        self.weight_path = weight_path
        self.load_ema = load_ema
        self.blocks = nn.ModuleList([_DiTBlock(hidden_size, num_groups) for _ in range(num_blocks)])
        self.initialize_weights()
        self.precompute_pos: Dict[Tuple[int, int], torch.Tensor] = {}
        self._prep = None
        self._prep_key = None

    def initialize_weights(self):
        """dit_c2i_baseline.py:332-355 (the output layers start at zero)."""
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.x_embedder.proj.bias, 0)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].weight, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)

    # -------------------------------------------------------------------------------------------- weight preparation
    @torch.no_grad()
    def prepare(self, device) -> dict:
        """bf16 copies / packed layouts of the fp32 master parameters; cached until a parameter changes."""
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._prep is not None and self._prep_key == key:
            return self._prep
        H, p = self.hidden_size, self.patch_size
        d = H // self.num_groups
        if d not in (64, 72):
            raise NotImplementedError(f"head_dim {d}: attention/qknorm kernels are built for 64 and 72")
        if p % 8 or H % 32 or not len(self.blocks):
            raise NotImplementedError("FlattenDiT kernels need patch_size % 8 == 0, hidden_size % 32 == 0, >= 1 block")

        def W(t):
            return t.detach().to(device=device, dtype=bf16).contiguous()

        def Fv(t):
            return t.detach().to(device=device, dtype=torch.float32).contiguous()

        P = {}
        P["wx"], P["bx"] = W(self.x_embedder.proj.weight), Fv(self.x_embedder.proj.bias)
        P["wt0"], P["bt0"] = W(self.t_embedder.mlp[0].weight), Fv(self.t_embedder.mlp[0].bias)
        P["wt2"], P["bt2"] = W(self.t_embedder.mlp[2].weight), Fv(self.t_embedder.mlp[2].bias)
        P["ytab"] = Fv(self.y_embedder.embedding_table.weight)
        # adaLN of every block and of the final layer (2H more rows) as ONE weight: one GEMM per forward
        fin = self.final_layer.adaLN_modulation[0]
        P["wada"] = W(torch.cat([b.adaLN_modulation[0].weight for b in self.blocks] + [fin.weight], 0))
        P["bada"] = Fv(torch.cat([b.adaLN_modulation[0].bias for b in self.blocks] + [fin.bias], 0))
        P["blocks"], P["ffn_pad"] = prepare_dit_blocks(self.blocks, H, device)
        P["wfin"], P["bfin"] = W(self.final_layer.linear.weight), Fv(self.final_layer.linear.bias)
        self._prep, self._prep_key = P, key
        return P

    def fetch_pos(self, height, width, device):
        """RoPE table cache per (h, w) (dit_c2i_baseline.py:324-330); here as real (cos, sin)."""
        key = (height, width)
        if key not in self.precompute_pos:
            self.precompute_pos[key] = rope_cos_sin(self.hidden_size // self.num_groups, height, width)
        tab = self.precompute_pos[key]
        if tab.device != torch.device(device):
            tab = tab.to(device)
            self.precompute_pos[key] = tab
        return tab

    # -------------------------------------------------------------------------------------------- forward
    def _forward_impl(self, x, t, y, masks=None):
        if masks is not None and not (isinstance(masks, (list, tuple)) and all(m is None for m in masks)):
            raise NotImplementedError("attention masks are not supported (every config passes masks=None)")
        if not x.is_cuda:
            raise RuntimeError("deco_b200.FlattenDiT runs on CUDA (sm_100a) only; there is no CPU fallback")
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("the denoiser backward yields parameter gradients only (the reference never "
                                      "differentiates w.r.t. x_t); detach x")
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            # training step (net(x_t, t, y) + loss.backward()): one autograd node with a hand-written backward -- the DiT
            # blocks of deco_b200.autograd, the FinalLayer / fold tail next to them
            from .autograd import denoiser_train_apply
            return denoiser_train_apply(self, x, t, y), None
        B, Cc, Hh, Ww = x.shape
        p, H, heads = self.patch_size, self.hidden_size, self.num_groups
        assert Cc == self.in_channels and Hh % p == 0 and Ww % p == 0
        L = (Hh // p) * (Ww // p)
        nb = len(self.blocks)
        with torch.no_grad():
            P = self.prepare(x.device)
            x32 = x.detach().to(torch.float32).contiguous()
            pos = self.fetch_pos(Hh // p, Ww // p, x.device)
            xp = ops.patchify(x32, p)                                                     # [B*L, C*p*p] bf16
            tfreq = ops.timestep_freq(t.reshape(-1).to(torch.float32), self.t_embedder.frequency_embedding_size)
            h1 = ops.gemm(tfreq, P["wt0"], P["bt0"], ops.EPI_BIAS_SILU)
            temb = ops.gemm(h1, P["wt2"], P["bt2"], ops.EPI_BIAS)                         # [B, H]
            c = ops.cond_combine(temb, P["ytab"], y.reshape(-1))                          # silu(t + y) (:368)
            mod = ops.gemm(c, P["wada"], P["bada"], ops.EPI_BIAS)                         # [B, nb*6H + 2H]
            st = StreamState(B * L, H, P["ffn_pad"], x.device, heads=heads)
            if COMPOSITE_SHIFT and "wshift" not in P:
                P["wshift"], P["bshift"] = composite_shift_weights(self.blocks, P["blocks"], H, x.device)
            shw_all = ops.gemm(c, P["wshift"], P["bshift"], ops.EPI_BIAS_F32) if "wshift" in P else None
            s = fused_blocks(P["blocks"], mod, 0, st, xp, P["wx"], P["bx"], B, L, H, heads, pos, Ww // p, shw_all=shw_all)
            shift, scale = mod[:, nb * 6 * H:nb * 6 * H + H], mod[:, nb * 6 * H + H:]     # chunk(2) (:80)
            h = ops.layernorm_modulate(s, shift, scale, L, out=st.o)
            tok = ops.gemm(h, P["wfin"], P["bfin"], ops.EPI_BIAS)                         # [B*L, C*p*p] bf16
            out = ops.unpatchify(tok, B, Cc, Hh, Ww, p)
        return out, s.view(B, L, H)

    def forward(self, x, t, y, masks=None):
        """x [B,C,H,W], t [B] in [0,1], y [B] int64 (num_classes = null) -> [B,C,H,W] (bf16, as the reference under
        autocast)."""
        return self._forward_impl(x, t, y, masks)[0]

    def forward_sx(self, x, t, y, masks=None):
        """dit_c2i_baseline.py:381-401: also returns the last block's stream as [B, H, sqrt(L), sqrt(L)]."""
        out, s = self._forward_impl(x, t, y, masks)
        if s is None:
            raise NotImplementedError("forward_sx is an inference entry point; the training path returns the output only")
        B, L, H = s.shape
        r = int(math.sqrt(L))
        return out, s.to(bf16).reshape(B, r, r, H).permute(0, 3, 1, 2)
