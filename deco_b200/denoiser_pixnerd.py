"""B200-native PixNerd baseline denoiser (hyper-network pixel decoder) -- drop-in for the reference module.

Mirrors `src/models/transformer/dit_c2i_pixnerd.py:288-381` (class PixNerDiT; `configs_c2i/Baseline_PixNerd.yaml`): the same
constructor arguments and `state_dict` keys (DiT blocks and NerfBlocks share ONE `blocks` ModuleList, :325-330), the same
`forward(x, t, y, s=None, mask=None)`.  The DiT part is the fused block stream of the DeCo denoiser (same FlattenDiTBlock);
per NerfBlock the parameter generator is one tcgen05 GEMM ([tokens, H] x [H, 2 * 64 * 128]) whose bf16 output rows the
decoder kernel (csrc/nerf_decoder.cu) consumes through TMA as per-patch MLP weights.  Inference only.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from .denoiser import (PixNerDiT as _DeCoDiT, _DiTBlock, _Embed, _LabelEmbedder, _NerfEmbedder, _TimestepEmbedder, _Weight,
                       _sw32_tile, nerf_pos_table, prepare_dit_blocks, rope_cos_sin)

bf16 = torch.bfloat16


class _NerfBlock(nn.Module):
    """Parameter layout of NerfBlock (dit_c2i_pixnerd.py:250-258)."""

    def __init__(self, hidden_size_s: int, hidden_size_x: int, mlp_ratio: int):
        super().__init__()
        self.param_generator1 = nn.Sequential(nn.Linear(hidden_size_s, 2 * hidden_size_x ** 2 * mlp_ratio, bias=True))
        self.norm = _Weight(hidden_size_x)
        self.mlp_ratio = mlp_ratio


class _NerfFinalLayer(nn.Module):
    def __init__(self, hidden_size: int, out_channels: int):
        super().__init__()
        self.norm = _Weight(hidden_size)
        self.linear = nn.Linear(hidden_size, out_channels, bias=True)


class PixNerDiT(nn.Module):
    """Drop-in for src/models/transformer/dit_c2i_pixnerd.py::PixNerDiT."""
    cuda_graph_safe = True

    def __init__(self, in_channels=4, num_groups=12, hidden_size=1152, hidden_size_x=64, nerf_mlpratio=4, num_blocks=18,
                 num_cond_blocks=4, patch_size=2, num_classes=1000, learn_sigma=True, deep_supervision=0, weight_path=None,
                 load_ema=False):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels
        self.hidden_size = hidden_size
        self.hidden_size_x = hidden_size_x
        self.nerf_mlpratio = nerf_mlpratio
        self.num_groups = num_groups
        self.num_blocks = num_blocks
        self.num_cond_blocks = num_cond_blocks
        self.patch_size = patch_size
        self.x_embedder = _NerfEmbedder(in_channels, hidden_size_x, max_freqs=8)
        self.s_embedder = _Embed(in_channels * patch_size ** 2, hidden_size)
        self.t_embedder = _TimestepEmbedder(hidden_size)
        self.y_embedder = _LabelEmbedder(num_classes + 1, hidden_size)
        self.final_layer = _NerfFinalLayer(hidden_size_x, self.out_channels)
        self.weight_path = weight_path
        self.load_ema = load_ema
        self.blocks = nn.ModuleList([_DiTBlock(hidden_size, num_groups) for _ in range(num_cond_blocks)])
        self.blocks.extend([_NerfBlock(hidden_size, hidden_size_x, nerf_mlpratio) for _ in range(num_cond_blocks, num_blocks)])
        self.initialize_weights()
        self.precompute_pos: Dict[Tuple[int, int], torch.Tensor] = {}
        self._prep = None
        self._prep_key = None
        self.fused = True

    def initialize_weights(self):
        """dit_c2i_pixnerd.py:343-358."""
        w = self.s_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.s_embedder.proj.bias, 0)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        nn.init.zeros_(self.final_layer.linear.weight)
        nn.init.zeros_(self.final_layer.linear.bias)

    # -------------------------------------------------------------------------------------------- weight preparation
    def _weights_key(self, device):
        return (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    @torch.no_grad()
    def prepare(self, device) -> dict:
        key = self._weights_key(device)
        if self._prep is not None and self._prep_key == key:
            return self._prep
        H, Hx, p, C = self.hidden_size, self.hidden_size_x, self.patch_size, self.in_channels
        if Hx != 64 or self.nerf_mlpratio != 2 or p != 16 or C != 3:
            raise NotImplementedError("the hyper-network decoder kernel is built for in_channels=3, patch_size=16, "
                                      "hidden_size_x=64, nerf_mlpratio=2 (configs_c2i/Baseline_PixNerd.yaml)")
        d = H // self.num_groups
        if d not in (64, 72):
            raise NotImplementedError(f"head_dim {d}: attention/qknorm kernels are built for 64 and 72")

        def W(t):
            return t.detach().to(device=device, dtype=bf16).contiguous()

        def Fv(t):
            return t.detach().to(device=device, dtype=torch.float32).contiguous()

        nc = self.num_cond_blocks
        dit, nerf = list(self.blocks[:nc]), list(self.blocks[nc:])
        P = {}
        P["ws"], P["bs"] = W(self.s_embedder.proj.weight), Fv(self.s_embedder.proj.bias)
        P["wt0"], P["bt0"] = W(self.t_embedder.mlp[0].weight), Fv(self.t_embedder.mlp[0].bias)
        P["wt2"], P["bt2"] = W(self.t_embedder.mlp[2].weight), Fv(self.t_embedder.mlp[2].bias)
        P["ytab"] = Fv(self.y_embedder.embedding_table.weight)
        if nc:
            P["wada"] = W(torch.cat([b.adaLN_modulation[0].weight for b in dit], 0))
            P["bada"] = Fv(torch.cat([b.adaLN_modulation[0].bias for b in dit], 0))
        P["blocks"], P["ffn_pad"] = prepare_dit_blocks(dit, H, device)
        P["wgen"] = [W(b.param_generator1[0].weight) for b in nerf]
        P["bgen"] = [Fv(b.param_generator1[0].bias) for b in nerf]
        # constant blob of csrc/nerf_decoder.cu: Wf tile | norm weights | final norm | final bias | Wrgb | T
        wx = W(self.x_embedder.embedder[0].weight).float()                              # [64, C + 64]
        tab = nerf_pos_table(p, self.x_embedder.max_freqs).to(device).to(bf16).float()  # Linear input cast
        T = torch.zeros(p * p, 68, device=device)
        T[:, :Hx] = tab @ wx[:, C:].t() + Fv(self.x_embedder.embedder[0].bias)
        wrgb = torch.zeros(Hx, 4, device=device)
        wrgb[:, :C] = wx[:, :C]
        wf = torch.zeros(16, Hx, device=device)
        wf[:C] = W(self.final_layer.linear.weight).float()
        bias_f = torch.zeros(4, device=device)
        bias_f[:C] = Fv(self.final_layer.linear.bias)
        floats = torch.cat([Fv(b.norm.weight) for b in nerf] + [Fv(self.final_layer.norm.weight), bias_f, wrgb.reshape(-1),
                                                               T.reshape(-1)])
        blob = torch.cat([_sw32_tile(wf.to(bf16)).view(torch.uint8), floats.view(torch.uint8)]).contiguous()
        assert blob.numel() == _lib.load().deco_nerf_decoder_blob_bytes(len(nerf)), blob.numel()
        P["blob"] = blob
        self._prep, self._prep_key = P, key
        return P

    def _composite_shift(self, P, device):
        from .denoiser import composite_shift_weights
        return composite_shift_weights(list(self.blocks[:self.num_cond_blocks]), P["blocks"], self.hidden_size, device)

    def fetch_pos(self, height, width, device):
        key = (height, width)
        if key not in self.precompute_pos:
            self.precompute_pos[key] = rope_cos_sin(self.hidden_size // self.num_groups, height, width)
        tab = self.precompute_pos[key]
        if tab.device != torch.device(device):
            tab = tab.to(device)
            self.precompute_pos[key] = tab
        return tab

    # -------------------------------------------------------------------------------------------- forward
    def forward(self, x, t, y, s=None, mask=None):
        """x [B,C,H,W], t [B] in [0,1], y [B] int64 -> velocity [B,C,H,W] (bf16, as the reference under autocast)."""
        if mask is not None:
            raise NotImplementedError("attention masks are not supported (the reference always passes mask=None)")
        if not x.is_cuda:
            raise RuntimeError("deco_b200 PixNerd PixNerDiT runs on CUDA (sm_100a) only; there is no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or (self.training and any(p.requires_grad for p in self.parameters()))):
            raise NotImplementedError("the PixNerd baseline denoiser is inference-only here; call under torch.no_grad() / .eval()")
        B, Cc, Hh, Ww = x.shape
        p, H = self.patch_size, self.hidden_size
        assert Cc == self.in_channels and Hh % p == 0 and Ww % p == 0
        L = (Hh // p) * (Ww // p)
        with torch.no_grad():
            P = self.prepare(x.device)
            x32 = x.detach().to(torch.float32).contiguous()
            if s is None:
                pos = self.fetch_pos(Hh // p, Ww // p, x.device)
                s2 = _DeCoDiT._encode(self, P, ops.patchify(x32, p), t.reshape(-1).to(torch.float32), y.reshape(-1), B, L, pos,
                                      Ww // p)
            else:
                s2 = s.detach().reshape(B * L, H).to(bf16).contiguous()
            # per NerfBlock: the generated fc1 | fc2 of every patch (param_generator1, dit_c2i_pixnerd.py:260-263)
            params = [ops.gemm(s2, w, b, ops.EPI_BIAS) for w, b in zip(P["wgen"], P["bgen"])]
            return ops.nerf_decoder(x32, params, P["blob"], p, self.hidden_size_x, self.nerf_mlpratio)
