"""GPU parity of the training-step backward (BASELINE configs[3]): every backward kernel against PyTorch autograd of the
same op, then the whole denoiser + DCT/FM loss against autograd over the fp32 oracle.

Tolerances: the kernels take / emit bf16 GEMM operands like the forward (reference: bf16 autocast), so element-wise
outputs are compared at rel-L2 <= 1e-2 (bf16 rounding is 4e-3 worst case per element, ~2e-3 in L2) and fp32 reductions at
<= 2e-3; whole-model parameter gradients at <= 3e-2 per tensor against the fp32 oracle (the reference's own bf16-autocast
gradients sit at 1-2e-2 from its fp32 self on these sizes)."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import build_module, rel_l2
from oracle import deco_oracle as O

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _g(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


def test_transpose_cast_and_colsum(dev):
    from deco_b200 import ops
    for dt in (torch.float32, bf16):
        for R, C in [(70, 130), (256, 64), (5, 34)]:
            x = torch.randn(R, C + 6, device=dev, generator=_g(R)).to(dt)[:, :C]
            t = ops.transpose_cast(x)
            Rp = (R + 7) // 8 * 8
            assert t.shape == (C, Rp)
            assert torch.equal(t[:, :R], x.to(bf16).t())
            assert float(t[:, R:].abs().sum()) == 0.0
            out = ops.colsum_(torch.zeros(C, device=dev), x)
            assert rel_l2(out, x.float().sum(0)) < 1e-5


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (520, 144, 264), (3456, 1152, 2048), (96, 1152, 8), (72, 72, 1000)])
def test_gemm_tn_mn_major_operands(dev, M, N, K):
    """out = At^T . Wt with both operands MN-major in shared memory (wgrad without transposed copies)."""
    from deco_b200 import ops
    at = torch.randn(K, M, device=dev, generator=_g(M)).to(bf16)
    wt = torch.randn(K, N, device=dev, generator=_g(N + 1)).to(bf16)
    ref = at.float().t() @ wt.float()
    for tile_n in (0, 128, 256):
        out = ops.gemm_tn(at, wt, tile_n, split_k=1)
        assert rel_l2(out, ref) < 2e-3, (tile_n, rel_l2(out, ref))
    for split_k in (0, 2, 5):
        if split_k <= max(1, (K + 63) // 64):
            out = ops.gemm_tn(at, wt, 0, split_k=split_k)
            assert rel_l2(out, ref) < 2e-3, (split_k, rel_l2(out, ref))
    # strided operands (column slices of a wider buffer)
    big = torch.randn(K, M + N + 16, device=dev, generator=_g(7)).to(bf16)
    a2, w2 = big[:, 8:8 + M], big[:, 8 + M:8 + M + N]
    assert rel_l2(ops.gemm_tn(a2, w2), a2.float().t() @ w2.float()) < 2e-3


def test_gemm_split_k_skinny(dev):
    """d c = d mod . Wada: M = batch rows, K = 6 x hidden x blocks; K slices reduced with fp32 atomics."""
    from deco_b200 import ops
    M, N, K = 32, 1152, 6912 * 4
    a = torch.randn(M, K, device=dev, generator=_g(1)).to(bf16)
    w = (torch.randn(N, K, device=dev, generator=_g(2)) * K ** -0.5).to(bf16)
    ref = a.float() @ w.float().t()
    for split_k in (0, 1, 7):
        assert rel_l2(ops.gemm_f32_splitk(a, w, split_k), ref) < 2e-3


def test_wgrad_dgrad_through_the_gemm(dev):
    from deco_b200 import ops
    from deco_b200.autograd import _wgrad
    M, N, K = 520, 144, 264
    x = torch.randn(M, K, device=dev, generator=_g(1)).to(bf16)
    w = (torch.randn(N, K, device=dev, generator=_g(2)) * K ** -0.5).to(bf16)
    dy = torch.randn(M, N, device=dev, generator=_g(3)).to(bf16)
    dw = _wgrad(dy, x)
    assert rel_l2(dw, dy.float().t() @ x.float()) < 2e-3
    dx = ops.gemm(dy, ops.transpose_cast(w, rows_pad=N), None, ops.EPI_BIAS)
    assert rel_l2(dx.float(), dy.float() @ w.float()) < 5e-3
    # tiny reduction dimension (batch-sized wgrad: adaLN, t_embedder)
    dy2 = torch.randn(4, N, device=dev, generator=_g(4))
    x2 = torch.randn(4, K, device=dev, generator=_g(5)).to(bf16)
    assert rel_l2(_wgrad(dy2, x2), dy2.to(bf16).float().t() @ x2.float()) < 2e-3


@pytest.mark.parametrize("B,L,H", [(2, 16, 576), (3, 64, 1152), (2, 5, 1536), (1, 3, 136)])
def test_gate_residual_norm_equals_the_two_kernels(dev, B, L, H):
    """The fused residual add + next norm of the training forward is bit-identical to gate_residual + rmsnorm_modulate."""
    from deco_b200 import ops
    M = B * L
    s = torch.randn(M, H, device=dev, generator=_g(1))
    a = torch.randn(M, H, device=dev, generator=_g(2)).to(bf16)
    mod = torch.randn(B, 6 * H, device=dev, generator=_g(3)).to(bf16)
    gate, shift, scale = mod[:, 2 * H:3 * H], mod[:, 3 * H:4 * H], mod[:, 4 * H:5 * H]
    w = torch.randn(H, device=dev, generator=_g(4))
    s_ref = ops.gate_residual(s, a, gate, L)
    h_ref = ops.rmsnorm_modulate(s_ref, w, shift, scale, L)
    s_out, h = ops.gate_residual_norm(s, a, gate, L, w, shift, scale)
    assert torch.equal(s_out, s_ref) and torch.equal(h, h_ref)
    x = s + gate.float().repeat_interleave(L, 0) * a.float()
    ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * w * (1 + scale.float().repeat_interleave(L, 0)) \
        + shift.float().repeat_interleave(L, 0)
    assert rel_l2(h.float(), ref) < 5e-3


@pytest.mark.parametrize("B,L,H,with_bias", [(2, 16, 576, True), (3, 64, 1152, False), (2, 4, 144, True)])
def test_rmsnorm_bwd_gate_equals_the_two_kernels(dev, B, L, H, with_bias):
    """norm backward + the next gate backward fused == rmsnorm_modulate_bwd_ followed by gate_bwd."""
    from deco_b200 import ops
    M = B * L
    g = lambda k: _g(100 + k)   # noqa: E731
    ds0 = torch.randn(M, H, device=dev, generator=g(1))
    dh = torch.randn(M, H, device=dev, generator=g(2)).to(bf16)
    x = torch.randn(M, H, device=dev, generator=g(3))
    w = torch.randn(H, device=dev, generator=g(4))
    mod = torch.randn(B, 6 * H, device=dev, generator=g(5)).to(bf16)
    scale, gate = mod[:, H:2 * H], mod[:, 5 * H:6 * H]
    a = torch.randn(M, H, device=dev, generator=g(6)).to(bf16)

    def run(fused):
        ds, dw, dmod = ds0.clone(), torch.zeros(H, device=dev), torch.zeros(B, 6 * H, device=dev)
        dbias = torch.zeros(H, device=dev) if with_bias else None
        if fused:
            da = ops.rmsnorm_modulate_bwd_gate_(ds, dh, x, w, scale, dw, dmod[:, :H], dmod[:, H:2 * H], L, a, gate,
                                                dmod[:, 5 * H:6 * H], dbias=dbias)
        else:
            ops.rmsnorm_modulate_bwd_(ds, dh, x, w, scale, dw, dmod[:, :H], dmod[:, H:2 * H], L)
            da = ops.gate_bwd(ds, a, gate, dmod[:, 5 * H:6 * H], L, dbias=dbias)
        return ds, dw, dmod, da, dbias

    r0, r1 = run(False), run(True)
    assert torch.equal(r0[0], r1[0]) and torch.equal(r0[3], r1[3])         # ds and da: same arithmetic per element
    for u, v in zip(r0[1:3], r1[1:3]):                                       # atomically reduced sums: order differs
        assert rel_l2(v, u) < 1e-5
    if with_bias:
        assert rel_l2(r1[4], r0[4]) < 1e-5


@pytest.mark.parametrize("B,L,H", [(2, 16, 576), (3, 64, 1152), (2, 4, 144)])
def test_gate_and_silu_rows_backward(dev, B, L, H):
    from deco_b200 import ops
    M = B * L
    ds = torch.randn(M, H, device=dev, generator=_g(1))
    a = torch.randn(M, H, device=dev, generator=_g(2)).to(bf16)
    mod = torch.randn(B, 3 * H, device=dev, generator=_g(3)).to(bf16)
    gate = mod[:, H:2 * H]
    s = torch.randn(M, H, device=dev, generator=_g(4))
    # forward
    out = ops.gate_residual(s, a, gate, L)
    ref = s + gate.float().repeat_interleave(L, 0) * a.float()
    assert rel_l2(out, ref) < 1e-6
    # backward
    dmod = torch.zeros(B, 3 * H, device=dev)
    dbias = torch.zeros(H, device=dev)
    da = ops.gate_bwd(ds, a, gate, dmod[:, H:2 * H], L, dbias=dbias)
    da_ref = gate.float().repeat_interleave(L, 0) * ds
    assert rel_l2(da.float(), da_ref) < 5e-3
    assert rel_l2(dmod[:, H:2 * H], (ds * a.float()).view(B, L, H).sum(1)) < 1e-4
    assert float(dmod[:, :H].abs().sum()) == 0.0 and float(dmod[:, 2 * H:].abs().sum()) == 0.0
    assert rel_l2(dbias, da_ref.sum(0)) < 1e-4
    # silu(x + row)
    row = torch.randn(B, H, device=dev, generator=_g(5)).to(bf16)
    dout = torch.randn(M, H, device=dev, generator=_g(6)).to(bf16)
    x = s.clone().requires_grad_(True)
    r = row.float().clone().requires_grad_(True)
    F.silu(x + r.repeat_interleave(L, 0)).backward(dout.float())
    drow = torch.zeros(B, H, device=dev)
    dx = ops.silu_add_rows_bwd(dout, s, row, drow, L)
    assert rel_l2(dx, x.grad) < 1e-4 and rel_l2(drow, r.grad) < 1e-4


def test_swiglu_forward_backward(dev):
    from deco_b200 import ops
    M, Fp = 130, 96
    y13 = torch.randn(M, 2 * Fp, device=dev, generator=_g(1)).to(bf16)
    du = torch.randn(M, Fp, device=dev, generator=_g(2)).to(bf16)
    yy = y13.float().view(M, Fp // 16, 2, 16).clone().requires_grad_(True)
    u_ref = (F.silu(yy[:, :, 0]) * yy[:, :, 1]).reshape(M, Fp)
    u_ref.backward(du.float())
    u = ops.swiglu_fwd(y13)
    assert rel_l2(u.float(), u_ref) < 5e-3
    dy = ops.swiglu_bwd(y13, du)
    assert rel_l2(dy.float(), yy.grad.reshape(M, 2 * Fp)) < 5e-3


@pytest.mark.parametrize("Fp,H,T", [(48, 64, 200), (688, 144, 1024), (3072, 1152, 2048)])
def test_wgrad_deinterleaved_rows(dev, Fp, H, T):
    """The SwiGLU weight gradient: dY's 2 Fp columns are the interleaved [16 x w1 | 16 x w3] columns; the TN GEMM with
    de-interleaving leaves (dW1 ; dW3) stacked (deco_gemm_bf16_tn_deint16), equal to pulling the plain result apart."""
    from deco_b200 import ops
    dy = torch.randn(T, 2 * Fp, device=dev, generator=_g(1)).to(bf16)
    x = torch.randn(T, H, device=dev, generator=_g(2)).to(bf16)
    plain = ops.gemm_tn(dy, x, split_k=1).view(Fp // 16, 2, 16, H)
    out = torch.full((2 * Fp + 8, H), 7.0, device=dev)
    ops.gemm_tn(dy, x, out=out[:2 * Fp], deinterleave16=True, split_k=1)
    assert torch.equal(out[:Fp], plain[:, 0].reshape(Fp, H)) and torch.equal(out[Fp:2 * Fp], plain[:, 1].reshape(Fp, H))
    assert bool((out[2 * Fp:] == 7.0).all())
    ref = dy.float().t() @ x.float()
    assert rel_l2(plain.reshape(2 * Fp, H), ref) < 2e-3
    auto = ops.gemm_tn(dy, x, deinterleave16=True)       # automatic split-K (atomics into the zeroed stacked output)
    assert rel_l2(auto, out[:2 * Fp]) < 1e-5


@pytest.mark.parametrize("M,K,Fp", [(300, 144, 688), (1024, 1152, 3072), (77, 64, 48)])
def test_swiglu_gemm_epilogues_vs_autograd(dev, M, K, Fp):
    """The training GEMMs that carry the SwiGLU passes (csrc/gemm_tcgen05.cu EPI_SWIGLU_DUAL / EPI_SWIGLU_BWD) against
    autograd over the same bf16 operands: y13 = h W13^T, u = silu(a) b (dit_c2i_DeCo.py:113), dy13 from du = da W2."""
    from deco_b200 import ops
    h = torch.randn(M, K, device=dev, generator=_g(1)).to(bf16)
    w13 = (torch.randn(2 * Fp, K, device=dev, generator=_g(2)) * K ** -0.5).to(bf16)     # rows interleaved [16 w1 | 16 w3]
    if Fp % 16 == 0 and (2 * Fp) % 32 == 0:
        y13 = torch.full((M + 1, 2 * Fp), 7.0, device=dev).to(bf16)      # one guard row
        u = ops.gemm(h, w13, None, ops.EPI_SWIGLU_DUAL, aux=y13[:M])
        y_ref = h.float() @ w13.float().t()
        yy = y_ref.view(M, Fp // 16, 2, 16)
        u_ref = (F.silu(yy[:, :, 0]) * yy[:, :, 1]).reshape(M, Fp)
        assert rel_l2(y13[:M].float(), y_ref) < 5e-3 and rel_l2(u.float(), u_ref) < 5e-3
        assert bool((y13[M] == 7.0).all())
        assert torch.equal(y13[:M], ops.gemm(h, w13, None, ops.EPI_BIAS))       # the very bf16 the plain GEMM writes
    # backward: du = da . W2 (the dgrad GEMM takes W2^T [Fp, H] as its weight operand), dy13 = SwiGLU'(y13) du
    Hh = 96
    da = torch.randn(M, Hh, device=dev, generator=_g(3)).to(bf16)
    w2T = (torch.randn(Fp, Hh, device=dev, generator=_g(4)) * Hh ** -0.5).to(bf16)
    y13 = torch.randn(M, 2 * Fp, device=dev, generator=_g(5)).to(bf16)
    yy = y13.float().view(M, Fp // 16, 2, 16).clone().requires_grad_(True)
    (F.silu(yy[:, :, 0]) * yy[:, :, 1]).reshape(M, Fp).backward(da.float() @ w2T.float().t())
    dy = torch.full((M + 1, 2 * Fp), 7.0, device=dev).to(bf16)
    ops.gemm(da, w2T, None, ops.EPI_SWIGLU_BWD, aux=y13, out=dy[:M])
    assert rel_l2(dy[:M].float(), yy.grad.reshape(M, 2 * Fp)) < 5e-3
    assert bool((dy[M] == 7.0).all())
    two = ops.swiglu_bwd(y13, ops.gemm(da, w2T, None, ops.EPI_BIAS))            # the two-kernel form it replaces
    assert rel_l2(dy[:M].float(), two.float()) < 6e-3


@pytest.mark.parametrize("B,L,H", [(2, 16, 576), (2, 64, 1152), (1, 4, 1024)])
def test_rmsnorm_modulate_backward(dev, B, L, H):
    from deco_b200 import ops
    M = B * L
    x = torch.randn(M, H, device=dev, generator=_g(1)) * 1.7
    w = (1 + 0.2 * torch.randn(H, device=dev, generator=_g(2)))
    mod = (0.3 * torch.randn(B, 2 * H, device=dev, generator=_g(3))).to(bf16)
    shift, scale = mod[:, :H], mod[:, H:]
    dh = torch.randn(M, H, device=dev, generator=_g(4)).to(bf16)
    xa, wa = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    sh, sc = shift.float().clone().requires_grad_(True), scale.float().clone().requires_grad_(True)
    h = O.modulate(O.rmsnorm(xa.view(B, L, H), wa), sh.view(B, 1, H), sc.view(B, 1, H))
    h.backward(dh.float().view(B, L, H))
    ds0 = torch.randn(M, H, device=dev, generator=_g(5))
    ds = ds0.clone()
    dmod = torch.zeros(B, 2 * H, device=dev)
    dw = torch.zeros(H, device=dev)
    ops.rmsnorm_modulate_bwd_(ds, dh, x, w, scale, dw, dmod[:, :H], dmod[:, H:], L)
    assert rel_l2(ds - ds0, xa.grad) < 1e-4
    assert rel_l2(dw, wa.grad) < 1e-4
    assert rel_l2(dmod[:, :H], sh.grad) < 1e-4 and rel_l2(dmod[:, H:], sc.grad) < 1e-4


@pytest.mark.parametrize("heads,d,hw", [(8, 72, 4), (4, 64, 8)])
def test_headnorm_rope_backward(dev, heads, d, hw):
    from deco_b200 import ops
    from deco_b200.denoiser import rope_cos_sin
    B, L, H = 2, hw * hw, heads * d
    M = B * L
    raw = torch.randn(M, 3 * H, device=dev, generator=_g(1)).to(bf16)
    qw = 1 + 0.2 * torch.randn(d, device=dev, generator=_g(2))
    kw = 1 + 0.2 * torch.randn(d, device=dev, generator=_g(3))
    g = torch.randn(M, 3 * H, device=dev, generator=_g(4)).to(bf16)
    pos = rope_cos_sin(d, hw, hw).to(dev)
    ang = O.rope_table_2d(d, hw, hw).to(dev)
    ra = raw.float().clone().requires_grad_(True)
    qa, ka = qw.clone().requires_grad_(True), kw.clone().requires_grad_(True)
    q, k, v = ra.view(B, L, 3, heads, d).unbind(2)
    q = O.apply_rope(O.rmsnorm(q, qa), ang)
    k = O.apply_rope(O.rmsnorm(k, ka), ang)
    out = torch.stack([q, k, v], 2).reshape(M, 3 * H)
    out.backward(g.float())
    # forward consistency of the convention
    fw = raw.clone()
    ops.qknorm_rope_(fw, qw, kw, pos, heads, d, L)
    assert rel_l2(fw.float(), out) < 1e-2
    gg = g.clone()
    dq, dk = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    ops.headnorm_rope_bwd_(gg, raw, 0, qw, pos, dq, heads, d, L)
    ops.headnorm_rope_bwd_(gg, raw, H, kw, pos, dk, heads, d, L)
    assert rel_l2(gg.float(), ra.grad) < 1e-2
    assert rel_l2(dq, qa.grad) < 2e-3 and rel_l2(dk, ka.grad) < 2e-3


def test_cond_combine_and_silu_backward(dev):
    from deco_b200 import ops
    B, H, ncls = 5, 576, 11
    temb = torch.randn(B, H, device=dev, generator=_g(1)).to(bf16)
    table = torch.randn(ncls, H, device=dev, generator=_g(2))
    y = torch.tensor([1, 3, 3, 10, 0], device=dev)
    dc = torch.randn(B, H, device=dev, generator=_g(3))
    ta, tb = temb.float().clone().requires_grad_(True), table.clone().requires_grad_(True)
    F.silu(ta + F.embedding(y, tb)).backward(dc)
    dtemb0 = torch.randn(B, H, device=dev, generator=_g(4))
    dtemb, dtab = dtemb0.clone(), torch.zeros_like(table)
    ops.cond_combine_bwd_(dc, temb, table, y, dtemb, dtab)
    assert rel_l2(dtemb - dtemb0, ta.grad) < 1e-4 and rel_l2(dtab, tb.grad) < 1e-4
    z = torch.randn(B, H, device=dev, generator=_g(5)).to(bf16)
    dy = torch.randn(B, H, device=dev, generator=_g(6)).to(bf16)
    za = z.float().clone().requires_grad_(True)
    F.silu(za).backward(dy.float())
    assert rel_l2(ops.silu_bwd(z, dy).float(), za.grad) < 5e-3


@pytest.mark.parametrize("B,heads,d,L", [(2, 4, 72, 16), (2, 2, 64, 256), (1, 3, 72, 80), (1, 2, 72, 1024)])
def test_attention_backward(dev, B, heads, d, L):
    from deco_b200 import ops
    H = heads * d
    M = B * L
    qkv = torch.randn(M, 3 * H, device=dev, generator=_g(L)).to(bf16)
    do = torch.randn(M, H, device=dev, generator=_g(L + 1)).to(bf16)
    qa = qkv.float().clone().requires_grad_(True)
    q, k, v = (t.transpose(1, 2) for t in qa.view(B, L, 3, heads, d).unbind(2))     # [B, heads, L, d]
    o_ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(M, H)
    o_ref.backward(do.float())
    o = ops.attention(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d)
    assert rel_l2(o.float(), o_ref) < 1e-2
    dqkv = torch.zeros_like(qkv)
    ops.attention_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, do,
                      dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], B, heads, d)
    g = qa.grad
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        e = rel_l2(dqkv[:, sl].float(), g[:, sl])
        assert e < 1.5e-2, (name, e)
    # with the forward's softmax statistics (the training path): same gradients, no statistics pass in the dQ kernel
    o2, lse = ops.attention_lse(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, heads, d)
    assert torch.equal(o2, o)
    s_ref = (q.detach() @ k.detach().transpose(-1, -2)) * d ** -0.5                    # [B, heads, L, L]
    lse_ref = torch.logsumexp(s_ref, -1) / math.log(2.0)
    assert rel_l2(lse.view(B, heads, L), lse_ref) < 1e-3
    dqkv2 = torch.zeros_like(qkv)
    ops.attention_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, do,
                      dqkv2[:, :H], dqkv2[:, H:2 * H], dqkv2[:, 2 * H:], B, heads, d, lse=lse)
    assert rel_l2(dqkv2.float(), g) < 1.5e-2


def _decoder_ref(P, cfg, x, y):
    """oracle pixel path with the per-pixel condition y [BL, p*p, 32] given (O.denoiser_forward + O.pixel_decoder minus
    cond_embed), so that autograd yields d y."""
    B, _, Hh, Ww = x.shape
    p, Hx = cfg.patch_size, cfg.hidden_size_x
    xp = F.unfold(x, kernel_size=p, stride=p).transpose(1, 2)
    L = xp.shape[1]
    px = xp.reshape(B * L, cfg.in_channels, p * p).transpose(1, 2)
    tab = O.nerf_pos_table(p, cfg.max_freqs).to(x.device)
    h = F.linear(torch.cat([px, tab[None].expand(B * L, -1, -1)], -1), P["x_embedder.embedder.0.weight"],
                 P["x_embedder.embedder.0.bias"])
    h = F.linear(h, P["dec_net.input_proj.weight"], P["dec_net.input_proj.bias"])
    for j in range(cfg.num_res_blocks):
        b = f"dec_net.res_blocks.{j}."
        mod = F.linear(F.silu(y), P[b + "adaLN_modulation.1.weight"], P[b + "adaLN_modulation.1.bias"])
        sh, sc, g = mod.chunk(3, dim=-1)
        t = O.modulate(F.layer_norm(h, (Hx,), P[b + "in_ln.weight"], P[b + "in_ln.bias"], 1e-6), sh, sc)
        t = F.linear(F.silu(F.linear(t, P[b + "mlp.0.weight"], P[b + "mlp.0.bias"])), P[b + "mlp.2.weight"], P[b + "mlp.2.bias"])
        h = h + g * t
    h = F.layer_norm(h, (Hx,), None, None, 1e-6)
    out = F.linear(h, P["dec_net.final_layer.linear.weight"], P["dec_net.final_layer.linear.bias"])
    out = out.transpose(1, 2).reshape(B, L, -1)
    return F.fold(out.transpose(1, 2).contiguous(), (Hh, Ww), kernel_size=p, stride=p)


@pytest.mark.parametrize("variant", ["mma", "scalar"])
@pytest.mark.parametrize("B,res", [(2, 64), (3, 32)])
def test_pixel_decoder_backward(dev, B, res, variant):
    from deco_b200 import ops
    from deco_b200.autograd import prepare_train
    cfg = O.DenoiserCfg(num_groups=2, hidden_size=144, num_blocks=4, num_cond_blocks=1, num_classes=10)
    m, P = build_module(cfg, dev)
    prep = m.prepare(dev)
    T = prepare_train(m, prep, dev)
    L = (res // 16) ** 2
    x = torch.randn(B, 3, res, res, device=dev, generator=_g(1))
    ycond = torch.randn(B * L, 256 * 32, device=dev, generator=_g(2)).to(bf16)
    dout = torch.randn(B, 3, res, res, device=dev, generator=_g(3))
    # reference: autograd over the oracle with the same bf16-rounded weights / inputs the kernels see
    names = [k for k in P if k.startswith("dec_net.") and "cond_embed" not in k] + ["x_embedder.embedder.0.weight", "x_embedder.embedder.0.bias"]
    Pr = {k: P[k].to(dev).clone().requires_grad_(True) for k in names}
    ya = ycond.float().view(B * L, 256, 32).clone().requires_grad_(True)
    ref = _decoder_ref(Pr, cfg, x.to(bf16).float(), ya)
    ref.backward(dout)
    out = ops.pixel_decoder(x, ycond, prep["blob"], prep["postab"], 16, 32, 3, out_dtype=torch.float32)
    assert rel_l2(out, ref) < 1e-2
    if variant == "scalar":
        dy, gdec = ops.pixel_decoder_bwd(x, ycond, dout, T["dec_blob"], prep["postab"], 16, 32, 3)
    else:
        dy, gdec = ops.pixel_decoder_bwd_tc(x, ycond, dout, prep["blob"], T["dec_bwd_blob"], prep["postab"], 16, 32, 3)
    e = rel_l2(dy.float().view(B * L, 256, 32), ya.grad)
    assert e < 1.5e-2, e
    # parameter gradients, read out of the blob layout (csrc/decoder_bwd.cu) the way deco_b200/autograd.py does

    def chk(name, got, tol=1.5e-2):
        err = rel_l2(got, Pr[name].grad)
        assert err < tol, (name, err)

    chk("dec_net.input_proj.weight", gdec[96:1120].view(32, 32))
    chk("dec_net.input_proj.bias", gdec[1120:1152])
    for j in range(3):
        o0 = 1152 + j * 5344
        pre = f"dec_net.res_blocks.{j}."
        chk(pre + "adaLN_modulation.1.weight", gdec[o0:o0 + 3072].view(96, 32))
        chk(pre + "adaLN_modulation.1.bias", gdec[o0 + 3072:o0 + 3168])
        chk(pre + "in_ln.weight", gdec[o0 + 3168:o0 + 3200])
        chk(pre + "in_ln.bias", gdec[o0 + 3200:o0 + 3232])
        chk(pre + "mlp.0.weight", gdec[o0 + 3232:o0 + 4256].view(32, 32))
        chk(pre + "mlp.0.bias", gdec[o0 + 4256:o0 + 4288])
        chk(pre + "mlp.2.weight", gdec[o0 + 4288:o0 + 5312].view(32, 32))
        chk(pre + "mlp.2.bias", gdec[o0 + 5312:o0 + 5344])
    of = 1152 + 3 * 5344
    chk("dec_net.final_layer.linear.weight", gdec[of:of + 128].view(4, 32)[:3])
    chk("dec_net.final_layer.linear.bias", gdec[of + 128:of + 131])
    chk("x_embedder.embedder.0.weight", torch.cat([gdec[0:96].view(32, 3),
        ops.gemm(ops.transpose_cast(gdec[T["dec_blob"].numel():].view(256, 32)), T["tabT"], None, ops.EPI_BIAS_F32)], 1), 2e-2)
    chk("x_embedder.embedder.0.bias", gdec[T["dec_blob"].numel():].view(256, 32).sum(0))


def _grad_check(dev, cfg, B, res, tol):
    from deco_b200 import LinearScheduler, REPATrainer
    m, P = build_module(cfg, dev)
    m.train()
    x = torch.tanh(torch.randn(B, 3, res, res, device=dev, generator=_g(11)))
    noise = torch.randn(B, 3, res, res, device=dev, generator=_g(12))
    t = torch.rand(B, device=dev, generator=_g(13))
    y = torch.randint(0, cfg.num_classes + 1, (B,), device=dev, generator=_g(14))
    x_t, v_t = O.make_xt_vt(x, noise, t)
    tr = REPATrainer(scheduler=LinearScheduler()).to(dev)
    out = m(x_t, t, y)
    assert out.requires_grad and out.dtype == torch.float32
    d = tr.loss(out, v_t)
    d["loss"].backward()
    # fp32 oracle under autograd
    Pr = {k: v.to(dev).clone().requires_grad_(True) for k, v in P.items()}
    ref = O.denoiser_forward(Pr, cfg, x_t, t, y)
    dr = O.dct_fm_loss(ref, v_t)
    dr["loss"].backward()
    assert rel_l2(out, ref) < 1e-2
    assert abs(float(d["loss"]) - float(dr["loss"])) <= 2e-2 * float(dr["loss"])
    worst = []
    num = den = 0.0
    for name, prm in m.named_parameters():
        assert prm.grad is not None, name
        g, gr = prm.grad.double(), Pr[name].grad.double()
        num += float((g - gr).pow(2).sum())
        den += float(gr.pow(2).sum())
        worst.append((rel_l2(g, gr), name))
    worst.sort(reverse=True)
    print("global grad rel-L2 %.3e; worst tensors: %s" % (math.sqrt(num / den), worst[:6]))
    assert math.sqrt(num / den) < tol, worst[:6]
    bad = [(e, n) for e, n in worst if e > 3 * tol]
    assert not bad, bad


def test_training_step_gradients_toy(dev):
    cfg = O.DenoiserCfg(num_groups=2, hidden_size=144, num_blocks=5, num_cond_blocks=2, num_classes=10)
    _grad_check(dev, cfg, B=3, res=64, tol=2e-2)


def test_training_step_gradients_d64_ragged_ffn(dev):
    # head_dim 64, hidden 256 -> FFN width int(2*1024/3) = 682 (padded to 688), 128 px -> 64 tokens
    cfg = O.DenoiserCfg(num_groups=4, hidden_size=256, num_blocks=6, num_cond_blocks=3, num_classes=10)
    _grad_check(dev, cfg, B=2, res=128, tol=2e-2)


def test_wgrad_side_stream_matches_single_stream(dev, monkeypatch):
    """The weight-gradient GEMMs on the second stream (autograd.WgradLane) leave the same gradients as the single-stream
    backward -- eagerly, twice in a row (allocator reuse across the two streams), and captured into a CUDA graph."""
    from deco_b200 import autograd as A
    cfg = O.DenoiserCfg(num_groups=4, hidden_size=288, num_blocks=6, num_cond_blocks=4, num_classes=10)
    m, _ = build_module(cfg, dev)
    m.train()
    x = torch.randn(4, 3, 128, 128, device=dev, generator=_g(21))
    t = torch.rand(4, device=dev, generator=_g(22))
    y = torch.tensor([1, 10, 3, 7], device=dev)
    w = torch.randn(4, 3, 128, 128, device=dev, generator=_g(23))

    def grads():
        for p in m.parameters():
            p.grad = None
        (m(x, t, y) * w).sum().backward()
        return {n: p.grad.clone() for n, p in m.named_parameters()}

    monkeypatch.setattr(A, "WGRAD_STREAM", False)
    g0 = grads()
    monkeypatch.setattr(A, "WGRAD_STREAM", True)
    g1, g2 = grads(), grads()
    def same(a, b):     # fp32 atomics (split-K, per-image sums, decoder) make two runs equal up to summation order only
        for n in b:
            lane_out = n.startswith("blocks.") and n.endswith(".weight") and b[n].dim() == 2 and "adaLN" not in n
            assert rel_l2(a[n], b[n]) < (1e-5 if lane_out else 2e-3), n

    same(g1, g0)
    same(g2, g0)
    # graph capture: fork / join of the side stream inside the capture
    static = {n: torch.zeros_like(p) for n, p in m.named_parameters()}
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        grads()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for p in m.parameters():
            p.grad = None
        (m(x, t, y) * w).sum().backward()
        for n, p in m.named_parameters():
            static[n].copy_(p.grad)
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    same(static, g0)


@pytest.mark.parametrize("hidden,groups,blocks,res,B", [(288, 4, 4, 64, 3), (256, 4, 3, 128, 2)])
def test_baseline_training_step_gradients(dev, hidden, groups, blocks, res, B):
    """Patch-linear baseline (FlattenDiT, dit_c2i_baseline.py:357-379) in .train() mode: parameter gradients of the
    hand-written backward (shared DiT blocks + FinalLayer / fold tail) against autograd over the fp32 oracle."""
    from helpers import build_baseline_module
    cfg = O.BaselineCfg(in_channels=3, num_groups=groups, hidden_size=hidden, num_blocks=blocks, patch_size=16, num_classes=10)
    m, P = build_baseline_module(cfg, dev)
    m.train()
    x = torch.tanh(torch.randn(B, 3, res, res, device=dev, generator=_g(31)))
    t = torch.rand(B, device=dev, generator=_g(32))
    y = torch.randint(0, cfg.num_classes + 1, (B,), device=dev, generator=_g(33))
    w = torch.randn(B, 3, res, res, device=dev, generator=_g(34))
    out = m(x, t, y)
    assert out.requires_grad
    (out.float() * w).sum().backward()
    Pr = {k: v.to(dev).clone().requires_grad_(True) for k, v in P.items()}
    ref = O.baseline_forward(Pr, cfg, x, t, y)
    (ref * w).sum().backward()
    assert rel_l2(out.float(), ref) < 1e-2
    worst, num, den = [], 0.0, 0.0
    for name, prm in m.named_parameters():
        assert prm.grad is not None, name
        g, gr = prm.grad.double(), Pr[name].grad.double()
        num += float((g - gr).pow(2).sum())
        den += float(gr.pow(2).sum())
        worst.append((rel_l2(g, gr), name))
    worst.sort(reverse=True)
    print("baseline global grad rel-L2 %.3e; worst tensors: %s" % (math.sqrt(num / den), worst[:6]))
    assert math.sqrt(num / den) < 2e-2, worst[:6]
    assert not [(e, n) for e, n in worst if e > 6e-2], worst[:6]


def test_t2i_training_step_gradients_xxl(dev):
    """The configs_t2i/sft_res512.yaml architecture (DeCo-XXL: H 1536, 24 heads of 64, 4 text + 16 joint blocks, 128 text
    tokens of 2048) at 256 px, 2 images."""
    _t2i_grad_check(dev, O.CFG_XXL_T2I, 256, 2)


@pytest.mark.parametrize("hidden,groups,res,B", [(256, 4, 64, 2), (288, 4, 128, 2)])
def test_t2i_training_step_gradients(dev, hidden, groups, res, B):
    """Text-to-image denoiser (dit_t2i_pixnerd.py:276-297 + SimpleMLPAdaLN) in .train() mode: parameter gradients of the
    hand-written backward (text embedding, text-refine blocks, joint-attention image blocks with the kv_y branch into the
    text stream, pixel decoder) against autograd over the fp32 oracle."""
    cfg = O.T2ICfg(in_channels=3, num_groups=groups, hidden_size=hidden, decoder_hidden_size=32, num_encoder_blocks=3,
                   num_decoder_blocks=2, num_text_blocks=2, patch_size=16, txt_embed_dim=64, txt_max_length=24)
    _t2i_grad_check(dev, cfg, res, B)


def _t2i_grad_check(dev, cfg, res, B):
    from helpers import build_t2i_module
    m, P = build_t2i_module(cfg, dev)
    m.train()
    x = torch.tanh(torch.randn(B, 3, res, res, device=dev, generator=_g(41)))
    t = torch.rand(B, device=dev, generator=_g(42))
    y = torch.randn(B, cfg.txt_max_length, cfg.txt_embed_dim, device=dev, generator=_g(43))
    w = torch.randn(B, 3, res, res, device=dev, generator=_g(44))
    out = m(x, t, y)
    assert out.requires_grad and out.dtype == torch.float32
    (out * w).sum().backward()
    Pr = {k: v.to(dev).clone().requires_grad_(True) for k, v in P.items()}
    ref = O.t2i_forward(Pr, cfg, x, t, y)
    (ref * w).sum().backward()
    assert rel_l2(out, ref) < 1e-2
    worst, num, den = [], 0.0, 0.0
    for name, prm in m.named_parameters():
        assert prm.grad is not None, name
        g, gr = prm.grad.double(), Pr[name].grad.double()
        assert g.shape == gr.shape, name
        num += float((g - gr).pow(2).sum())
        den += float(gr.pow(2).sum())
        worst.append((rel_l2(g, gr), name))
    worst.sort(reverse=True)
    print("t2i global grad rel-L2 %.3e; worst tensors: %s" % (math.sqrt(num / den), worst[:8]))
    assert math.sqrt(num / den) < 2e-2, worst[:8]
    assert not [(e, n) for e, n in worst if e > 6e-2], worst[:8]


def test_center_rows_kernel(dev):
    from deco_b200 import ops
    for M, H in ((70, 144), (513, 1152), (9, 2048)):
        x = torch.randn(M, H, device=dev, generator=_g(M)) + 3.0
        ref = x - x.mean(1, keepdim=True)
        assert (ops.center_rows(x) - ref).abs().max() < 1e-5
        y = x.clone()
        ops.center_rows(y, out=y)
        assert (y - ref).abs().max() < 1e-5


def test_optimizer_step_invalidates_weight_cache(dev):
    """prepare() keys on parameter versions: after an optimizer step the next forward must see the new weights."""
    cfg = O.DenoiserCfg(num_groups=2, hidden_size=144, num_blocks=4, num_cond_blocks=1, num_classes=10)
    m, _ = build_module(cfg, dev)
    m.train()
    opt = torch.optim.SGD(m.parameters(), lr=0.5)
    x = torch.randn(2, 3, 64, 64, device=dev, generator=_g(1))
    t = torch.tensor([0.3, 0.8], device=dev)
    y = torch.tensor([1, 10], device=dev)
    out0 = m(x, t, y)
    out0.square().mean().backward()
    opt.step()
    with torch.no_grad():
        out1 = m(x, t, y).float()
    assert rel_l2(out1, out0) > 1e-3


def test_training_step_gradients_xl16(dev):
    """DeCo-XL/16 at 256 px (configs_c2i/DeCo_XL.yaml architecture, BASELINE configs[3] shape at batch 2)."""
    _grad_check(dev, O.CFG_XL, B=2, res=256, tol=2e-2)


def test_fused_adamw_ema_matches_torch(dev):
    """csrc/optimizer.cu against torch.optim.AdamW (configs_c2i/DeCo_XL.yaml:89-93) + the SimpleEMA update
    (src/callbacks/simple_ema.py:27-33), three steps, tensors with unaligned tails and sizes across the chunk boundary."""
    from deco_b200 import FusedAdamWEMA
    shapes = [(7,), (33, 5), (4096 * 4 + 3,), (129, 130), (1,)]
    g = _g(3)
    ps = [torch.nn.Parameter(torch.randn(s, device=dev, generator=g)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ema = [p.detach().clone() for p in ps]
    ema_ref = [p.detach().clone() for p in ps]
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    opt = FusedAdamWEMA(ps, ema, ema_decay=0.99, **kw)
    topt = torch.optim.AdamW(ref, **kw)
    for it in range(3):
        v0 = ps[0]._version
        for p, r in zip(ps, ref):
            gr = torch.randn(p.shape, device=dev, generator=g)
            p.grad, r.grad = gr.clone(), gr.clone()
        opt.step()
        topt.step()
        with torch.no_grad():
            torch._foreach_mul_(ema_ref, 0.99)
            torch._foreach_add_(ema_ref, ref, alpha=0.01)
        assert ps[0]._version > v0
        for p, r, e, er in zip(ps, ref, ema, ema_ref):
            assert torch.allclose(p, r, rtol=2e-6, atol=2e-7), (it, p.shape, float((p - r).abs().max()))
            assert torch.allclose(e, er, rtol=2e-6, atol=2e-7)


def test_fused_adamw_unsynced_steps_with_moving_grad_buffers(dev):
    """the pointer table is re-uploaded through ONE pinned staging buffer whenever a .grad tensor moved.  With
    the host running ahead of the GPU (a long kernel queued first, no synchronisation between steps) every step must still
    see ITS gradients: compare with torch.optim.AdamW fed the same gradients."""
    from deco_b200 import FusedAdamWEMA
    g = _g(11)
    shapes = [(1 << 16,), (257, 129), (5,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev, generator=g)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    kw = dict(lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01)
    opt, topt = FusedAdamWEMA(ps, None, **kw), torch.optim.AdamW(ref, **kw)
    nstep = 6
    grads = [[torch.randn(s, device=dev, generator=g) for s in shapes] for _ in range(nstep)]
    keep = [[gr.clone() for gr in row] for row in grads]          # distinct buffers per step: the pointers move every step
    big = torch.randn(8192, 8192, device=dev)
    torch.cuda.synchronize()
    for _ in range(6):
        big = big @ big * 1e-4          # ~tens of ms of queued GPU work: the host loop below runs ahead of the device
    for it in range(nstep):
        for p, gr in zip(ps, keep[it]):
            p.grad = gr
        opt.step()
    torch.cuda.synchronize()
    for it in range(nstep):
        for r, gr in zip(ref, grads[it]):
            r.grad = gr
        topt.step()
    for p, r in zip(ps, ref):
        assert torch.allclose(p, r, rtol=5e-6, atol=5e-7), float((p - r).abs().max())


def test_fused_adamw_state_dict_roundtrip_with_torch(dev):
    """state_dict() has torch.optim.AdamW's layout (what Lightning stores in `optimizer_states`): load it into torch's
    AdamW and back, continue both, same parameters."""
    from deco_b200 import FusedAdamWEMA
    g = _g(5)
    shapes = [(33, 5), (4097,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev, generator=g)) for s in shapes]
    kw = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    opt = FusedAdamWEMA(ps, None, **kw)
    for _ in range(2):
        for p in ps:
            p.grad = torch.randn(p.shape, device=dev, generator=g)
        opt.step()
    import copy
    sd = copy.deepcopy(opt.state_dict())     # torch's load_state_dict keeps same-device tensors by reference: no aliasing
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    topt = torch.optim.AdamW(ref, **kw)
    topt.load_state_dict(sd)
    ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt2 = FusedAdamWEMA(ps2, None, lr=1.0)
    opt2.load_state_dict(topt.state_dict())
    assert opt2.step_count == 2 and opt2.lr == kw["lr"] and opt2.betas == kw["betas"]
    for p, r, q in zip(ps, ref, ps2):
        gr = torch.randn(p.shape, device=dev, generator=g)
        p.grad, r.grad, q.grad = gr.clone(), gr.clone(), gr.clone()
    opt.step(); topt.step(); opt2.step()
    for p, r, q in zip(ps, ref, ps2):
        assert torch.allclose(p, r, rtol=2e-6, atol=2e-7) and torch.equal(p, q)


def test_train_input_kernels_vs_reference_expressions(dev):
    """csrc/train_inputs.cu against the reference's tensor expressions (training_repa_DeCo.py:222-237,
    base/training.py:14-20) on the same draws: t within 1 ulp-level tolerance, x_t / v_t bit-exact given t, labels exact."""
    from deco_b200 import GVPScheduler, LinearScheduler, ops
    g = _g(21)
    B = 37
    nt = torch.randn(B, device=dev, generator=g)
    uu = torch.rand(B, device=dev, generator=g)
    us = torch.rand(B, device=dev, generator=g)
    us[:4] = torch.tensor([0.9, 0.90001, 0.0, 1.0], device=dev)            # the <= 0.9 boundary
    for shift in (1.0, 2.5):
        t, coef = ops.train_timesteps(nt, uu, us, shift, True)
        ref = O.time_shift(torch.where(us <= 0.9, torch.sigmoid(nt), uu), shift)
        assert torch.allclose(t, ref, rtol=3e-7, atol=1e-7), float((t - ref).abs().max())
        assert torch.equal(coef[:, 0], t) and torch.equal(coef[:, 1], 1 - t)
        assert float((coef[:, 2] - 1).abs().max()) == 0 and float((coef[:, 3] + 1).abs().max()) == 0
        t2, c2 = ops.train_timesteps(nt, uu, us, shift, False)
        assert c2 is None and torch.equal(t2, t)
    x = torch.tanh(torch.randn(B, 3, 24, 40, device=dev, generator=g))
    eps = torch.randn(x.shape, device=dev, generator=g)
    for sch in (LinearScheduler(), GVPScheduler()):
        coef = torch.stack([f(t).reshape(-1) for f in (sch.alpha, sch.sigma, sch.dalpha, sch.dsigma)], 1).contiguous()
        x_t, v_t = ops.flow_pair(x, eps, coef)
        assert torch.equal(x_t, sch.alpha(t) * x + eps * sch.sigma(t))
        assert torch.equal(v_t, sch.dalpha(t) * x + sch.dsigma(t) * eps)
    cond = torch.arange(B, device=dev) % 10
    unc = torch.full((B,), 10, device=dev)
    u = torch.rand(B, device=dev, generator=g)
    got = ops.label_dropout(cond, unc, u, 0.3)
    mask = (u < 0.3).to(cond.dtype)
    assert torch.equal(got, cond * (1 - mask) + unc * mask)
    with pytest.raises(Exception):
        ops.flow_pair(x[:, :, :3, :3].contiguous(), eps[:, :, :3, :3].contiguous(), coef)     # 27 elements per image: % 4 != 0


@pytest.mark.parametrize("flw", [0.0, 1.0])
def test_impl_trainstep_vs_oracle_under_fixed_generator(dev, flw):
    """REPATrainer.__call__ (label dropout -> t mixture -> x_t / v_t -> denoiser -> loss) against the oracle restatement
    of base/training.py:25-28 + training_repa_DeCo.py:216-288 (pinned to the reference by tests/golden/trainstep.npz) under
    the same CUDA generator seed: same labels, same t, same x_t handed to the net, loss within the forward tolerance.
    flw = 0 is the fork's objective (fm only, two dict keys), flw = 1 adds the block-DCT term."""
    from deco_b200 import LinearScheduler, REPATrainer
    cfg = O.DenoiserCfg(num_groups=2, hidden_size=144, num_blocks=4, num_cond_blocks=2, num_classes=10)
    m, P = build_module(cfg, dev)
    Pd = {k: v.to(dev) for k, v in P.items()}
    B = 6
    x = torch.tanh(torch.randn(B, 3, 64, 64, device=dev, generator=_g(31)))
    cond = torch.arange(B, device=dev) % 10
    unc = torch.full((B,), 10, device=dev)
    tr = REPATrainer(scheduler=LinearScheduler(), null_condition_p=0.5, timeshift=1.7, freq_loss_weight=flw).to(dev)
    seen = {}

    def rec_net(x_t, t, y):
        seen.update(x_t=x_t.clone(), t=t.clone(), y=y.clone())
        return m(x_t, t, y)
    torch.manual_seed(99)
    with torch.no_grad():
        d = tr(rec_net, None, None, x, cond, unc)
    seen_o = {}

    def onet(x_t, t, y):
        seen_o.update(x_t=x_t.clone(), t=t.clone(), y=y.clone())
        return O.denoiser_forward(Pd, cfg, x_t, t, y)
    torch.manual_seed(99)
    do = O.trainstep(onet, x, cond, unc, null_condition_p=0.5, timeshift=1.7, freq_loss_weight=flw)
    assert torch.equal(seen["y"], seen_o["y"]) and 0 < int((seen["y"] == 10).sum()) < B
    assert torch.allclose(seen["t"], seen_o["t"], rtol=3e-7, atol=1e-7)
    assert rel_l2(seen["x_t"], seen_o["x_t"]) < 1e-6
    assert set(d) == ({"fm_loss", "loss"} if flw == 0 else {"fm_loss", "fm_loss_freq", "loss"})
    for k in d:
        assert abs(float(d[k]) - float(do[k])) <= 2e-2 * abs(float(do[k])), (k, float(d[k]), float(do[k]))
    print(f"trainstep flw={flw}: loss {float(d['loss']):.6f} vs oracle {float(do['loss']):.6f}")


def test_training_loop_with_fused_optimizer_reduces_loss(dev):
    """forward + backward + FusedAdamWEMA for a few steps on one fixed batch: the weight caches follow the parameters and
    the loss goes down."""
    from deco_b200 import FusedAdamWEMA, LinearScheduler, REPATrainer
    cfg = O.DenoiserCfg(num_groups=2, hidden_size=144, num_blocks=4, num_cond_blocks=2, num_classes=10)
    m, _ = build_module(cfg, dev)
    m.train()
    tr = REPATrainer(scheduler=LinearScheduler()).to(dev)
    opt = FusedAdamWEMA(m.parameters(), lr=2e-3)
    x = torch.tanh(torch.randn(4, 3, 64, 64, device=dev, generator=_g(1)))
    noise = torch.randn(4, 3, 64, 64, device=dev, generator=_g(2))
    t = torch.tensor([0.2, 0.4, 0.6, 0.8], device=dev)
    y = torch.tensor([1, 2, 3, 10], device=dev)
    x_t, v_t = O.make_xt_vt(x, noise, t)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        d = tr.loss(m(x_t, t, y), v_t)
        d["loss"].backward()
        opt.step()
        losses.append(float(d["loss"]))
    print("losses", [round(v, 4) for v in losses])
    assert losses[-1] < 0.9 * losses[0]


def test_training_pipeline_step_and_checkpoint_resume(dev, tmp_path):
    """The original LightningModel.training_step wiring (pyc L105-113) without Lightning: conditioner -> trainer (label
    dropout, t, x_t / v_t, denoiser, loss) -> backward -> fused AdamW + EMA; a checkpoint written in the reference layout
    (src/lightning_model.py:333-350) resumes bit-identically, optimizer moments and EMA included."""
    from deco_b200 import LinearScheduler, PixNerDiT, REPATrainer
    from deco_b200.data import LabelConditioner, PixelAE
    from deco_b200.pipeline import TrainingPipeline
    from deco_b200.utils import randomize_
    kw = dict(in_channels=3, num_groups=2, hidden_size=144, hidden_size_x=32, num_blocks=4, num_cond_blocks=2, patch_size=16,
              num_classes=10)

    def make():
        net = randomize_(PixNerDiT(**kw).to(dev), seed=3)
        return TrainingPipeline(net, REPATrainer(scheduler=LinearScheduler(), null_condition_p=0.2), None,
                                LabelConditioner(10), PixelAE(), lr=2e-3, ema_decay=0.9, device=dev)
    g = torch.Generator().manual_seed(5)
    batches = [(torch.tanh(torch.randn(4, 3, 64, 64, generator=g)), [1, 2, 3, 4], {}) for _ in range(5)]
    a = make()
    w0 = a.denoiser.blocks[0].attn.qkv.weight.detach().clone()
    torch.manual_seed(11)
    losses = [float(a.training_step(b)["loss"]) for b in batches[:3]]
    assert a.global_step == 3 and all(math.isfinite(v) for v in losses)
    w3 = a.denoiser.blocks[0].attn.qkv.weight.detach()
    e3 = a.ema_denoiser.blocks[0].attn.qkv.weight.detach()
    assert not torch.equal(w3, w0)
    assert not any(p.requires_grad for p in a.ema_denoiser.parameters())
    # EMA after 3 steps of decay 0.9 lies strictly between the initial and the current weights
    assert float((e3 - w0).abs().max()) > 0 and float((e3 - w3).abs().max()) > 0
    path = a.save_checkpoint(str(tmp_path / "step3.ckpt"))
    keys = list(torch.load(path, map_location="cpu")["state_dict"])
    assert keys[0].startswith("denoiser.") and keys[-1].startswith("ema_denoiser.") and len(keys) == 2 * len(a.denoiser.state_dict())
    b = make()
    b.load_checkpoint(path)
    assert b.global_step == 3 and b.optimizer.step_count == 3 and b.ema_decay == 0.9
    torch.manual_seed(12)
    la = [float(a.training_step(x)["loss"]) for x in batches[3:]]
    torch.manual_seed(12)
    lb = [float(b.training_step(x)["loss"]) for x in batches[3:]]
    assert la[0] == lb[0], (la, lb)                       # same weights, same draws: the first resumed loss is identical
    assert abs(la[1] - lb[1]) <= 1e-3 * abs(la[1])        # then up to the fp32 atomics of the weight gradients
    for (n, p), (_, q) in zip(a.ema_denoiser.state_dict().items(), b.ema_denoiser.state_dict().items()):
        assert rel_l2(q, p) < 1e-4, n


def test_weight_reprep_graph_replay_matches_eager(dev):
    """Training loop: from the second re-preparation on, prepare() + prepare_train() (bf16 casts, w1|w3 interleave, dgrad
    transposes, decoder fragment / tcgen05 packing) are ONE CUDA-graph replay into static buffers; every prepared tensor
    must equal what a fresh eager preparation of the updated weights produces."""
    from deco_b200 import FusedAdamWEMA, LinearScheduler, REPATrainer
    from deco_b200.autograd import prepare_train
    cfg = O.DenoiserCfg(num_groups=2, hidden_size=144, num_blocks=4, num_cond_blocks=2, num_classes=10)
    m, _ = build_module(cfg, dev)
    m.train()
    tr = REPATrainer(scheduler=LinearScheduler()).to(dev)
    opt = FusedAdamWEMA(m.parameters(), lr=2e-3)
    x = torch.tanh(torch.randn(4, 3, 64, 64, device=dev, generator=_g(1)))
    y, unc = torch.tensor([1, 2, 3, 4], device=dev), torch.full((4,), 10, device=dev)
    for _ in range(4):
        opt.zero_grad()
        tr(m, None, None, x, y, unc)["loss"].backward()
        opt.step()
    P = m.prepare(dev)
    assert m.__dict__.get("_prep_graph") is not None and P is m._prep_graph["P"], "the re-preparation was not graph-replayed"
    with torch.no_grad():
        E = m._prepare_eager(dev)
        prepare_train(m, E, dev)

    def same(a, b, path):
        if torch.is_tensor(a):
            assert torch.equal(a, b), path
        elif isinstance(a, dict):
            assert set(a) == set(b), path
            for k in a:
                same(a[k], b[k], f"{path}.{k}")
        elif isinstance(a, (list, tuple)):
            assert len(a) == len(b), path
            for i, (u, v) in enumerate(zip(a, b)):
                same(u, v, f"{path}[{i}]")
        else:
            assert a == b, path
    same(P, E, "P")
