"""Per-kernel parity: every C-ABI entry point against a plain PyTorch fp32 evaluation of the same operands."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2
from oracle import deco_oracle as O

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def _rand(shape, dev, seed, scale=1.0, dtype=bf16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(device=dev, dtype=dtype)


@pytest.mark.parametrize("M,N,K,tile_n", [
    (256, 256, 128, 128), (256, 256, 128, 256), (300, 1152, 1152, 0), (128, 3456, 1152, 192),
    (520, 576, 576, 128), (8, 1024, 256, 0), (1024, 8192, 1152, 256), (640, 2736, 1024, 0), (384, 1152, 2736, 0),
])
def test_gemm_bias(cuda_dev, M, N, K, tile_n):
    from deco_b200 import ops
    a, w = _rand((M, K), cuda_dev, 1), _rand((N, K), cuda_dev, 2, K ** -0.5)
    bias = _rand((N,), cuda_dev, 3, 0.1, torch.float32)
    ref = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, bias, ops.EPI_BIAS, tile_n=tile_n)
    torch.cuda.synchronize()
    assert rel_l2(out.float(), ref) < 4e-3
    out = ops.gemm(a, w, None, ops.EPI_BIAS, tile_n=tile_n)
    assert rel_l2(out.float(), a.float() @ w.float().t()) < 4e-3
    out = ops.gemm(a, w, bias, ops.EPI_BIAS_SILU, tile_n=tile_n)
    assert rel_l2(out.float(), F.silu(ref)) < 4e-3


@pytest.mark.parametrize("cg,staged", [(1, 0), (1, 1), (2, 0), (2, 1)])
def test_gemm_variants_agree(cuda_dev, lib, cg, staged):
    """Every (cta_group, epilogue style) build of the kernel on a ragged shape (M, N not multiples of the tile)."""
    from deco_b200 import ops
    M, N, K, L = 1000, 1160, 1152, 250
    a, w = _rand((M, K), cuda_dev, 1), _rand((N, K), cuda_dev, 2, K ** -0.5)
    bias = _rand((N,), cuda_dev, 3, 0.1, torch.float32)
    resid = _rand((M, N), cuda_dev, 4, dtype=torch.float32)
    gate = _rand((M // L, N), cuda_dev, 5)
    ref = a.float() @ w.float().t() + bias
    lib.deco_gemm_set_tuning(cg, staged)
    try:
        for tn in (128, 192, 256):
            assert rel_l2(ops.gemm(a, w, bias, ops.EPI_BIAS, tile_n=tn).float(), ref) < 4e-3
            o = ops.gemm(a, w, bias, ops.EPI_GATE_RESIDUAL, resid=resid, gate=gate, rows_per_gate=L, tile_n=tn)
            assert rel_l2(o, resid + gate.float().repeat_interleave(L, 0) * ref) < 1e-5
        w13 = _rand((1184, K), cuda_dev, 6, K ** -0.5)       # 37 groups of [16 | 16]
        g = w13.view(37, 2, 16, K)
        sw = F.silu(a.float() @ g[:, 0].reshape(-1, K).float().t()) * (a.float() @ g[:, 1].reshape(-1, K).float().t())
        assert rel_l2(ops.gemm(a, w13, None, ops.EPI_SWIGLU).float(), sw) < 5e-3
    finally:
        lib.deco_gemm_set_tuning(-1, -1)


def test_gemm_strided_a(cuda_dev):
    """A operand as a column slice of a wider matrix (how attention output / qkv views are consumed)."""
    from deco_b200 import ops
    big = _rand((384, 3 * 256), cuda_dev, 5)
    a = big[:, 256:512]
    w = _rand((128, 256), cuda_dev, 6, 1 / 16)
    out = ops.gemm(a, w)
    assert rel_l2(out.float(), a.float() @ w.float().t()) < 4e-3


@pytest.mark.parametrize("M,N,K,L,tile_n", [(512, 1152, 1152, 256, 0), (200, 576, 3072, 100, 128), (512, 1024, 2736, 64, 256)])
def test_gemm_gate_residual(cuda_dev, M, N, K, L, tile_n):
    from deco_b200 import ops
    a, w = _rand((M, K), cuda_dev, 1), _rand((N, K), cuda_dev, 2, K ** -0.5)
    bias = _rand((N,), cuda_dev, 3, 0.1, torch.float32)
    resid = _rand((M, N), cuda_dev, 4, dtype=torch.float32)
    nb = (M + L - 1) // L
    mod = _rand((nb, 6 * N), cuda_dev, 5)
    gate = mod[:, 2 * N:3 * N]
    ref = resid.float() + gate.float().repeat_interleave(L, 0)[:M] * (a.float() @ w.float().t() + bias)
    out = resid.clone()
    ops.gemm(a, w, bias, ops.EPI_GATE_RESIDUAL, out=out, resid=out, gate=gate, rows_per_gate=L, tile_n=tile_n)
    assert out.dtype == torch.float32 and rel_l2(out, ref) < 1e-5        # fp32 accumulate, fp32 residual stream
    o32 = ops.gemm(a, w, bias, ops.EPI_BIAS_F32, tile_n=tile_n)
    assert o32.dtype == torch.float32 and rel_l2(o32, a.float() @ w.float().t() + bias) < 1e-5


@pytest.mark.parametrize("M,F_,K,tile_n", [(256, 3072, 1152, 0), (130, 2736, 1024, 0), (256, 512, 256, 128)])
def test_gemm_swiglu(cuda_dev, M, F_, K, tile_n):
    from deco_b200 import ops
    a = _rand((M, K), cuda_dev, 1)
    w1, w3 = _rand((F_, K), cuda_dev, 2, K ** -0.5), _rand((F_, K), cuda_dev, 3, K ** -0.5)
    w13 = torch.stack([w1.view(F_ // 16, 16, K), w3.view(F_ // 16, 16, K)], 1).reshape(2 * F_, K).contiguous()
    ref = F.silu(a.float() @ w1.float().t()) * (a.float() @ w3.float().t())
    out = ops.gemm(a, w13, None, ops.EPI_SWIGLU, tile_n=tile_n)
    assert out.shape == (M, F_)
    assert rel_l2(out.float(), ref) < 5e-3


def test_patchify_exact(cuda_dev):
    from deco_b200 import ops
    x = _rand((3, 3, 64, 96), cuda_dev, 7, dtype=torch.float32)
    ref = F.unfold(x, kernel_size=16, stride=16).transpose(1, 2).reshape(-1, 768).to(bf16)
    assert torch.equal(ops.patchify(x, 16), ref)


def test_timestep_and_cond(cuda_dev):
    from deco_b200 import ops
    t = torch.tensor([0.0, 0.5, 0.0999, 1.0], device=cuda_dev)
    ref = O.timestep_embedding(t)
    got = ops.timestep_freq(t).float()
    assert (got - ref).abs().max() < 5e-3          # bf16 storage
    assert abs(float(ref[1, 0]) - 0.8775826) < 1e-6 and abs(float(ref[1, 128 + 127]) - 0.0508856) < 1e-6
    temb = _rand((4, 128), cuda_dev, 1)
    table = _rand((11, 128), cuda_dev, 2, dtype=torch.float32)
    y = torch.tensor([0, 10, 3, 7], device=cuda_dev)
    ref = F.silu(temb.float() + table[y])
    assert rel_l2(ops.cond_combine(temb, table, y).float(), ref) < 4e-3


@pytest.mark.parametrize("xdt", [torch.float32, bf16])
@pytest.mark.parametrize("H", [1152, 1024, 576, 1536])
def test_rmsnorm_modulate(cuda_dev, H, xdt):
    from deco_b200 import ops
    M, L = 96, 16
    x = _rand((M, H), cuda_dev, 1, 2.0, dtype=xdt)
    w = 1 + 0.1 * _rand((H,), cuda_dev, 2, dtype=torch.float32)
    mod = _rand((M // L, 6 * H), cuda_dev, 3, 0.5)
    sh, sc = mod[:, :H], mod[:, H:2 * H]
    ref = O.modulate(O.rmsnorm(x, w), sh.float().repeat_interleave(L, 0), sc.float().repeat_interleave(L, 0))
    got = ops.rmsnorm_modulate(x, w, sh, sc, L)
    assert rel_l2(got.float(), ref) < 4e-3


@pytest.mark.parametrize("heads,d,hw", [(16, 72, (4, 4)), (8, 64, (3, 5))])
def test_qknorm_rope(cuda_dev, heads, d, hw):
    from deco_b200 import ops
    from deco_b200.denoiser import rope_cos_sin
    L = hw[0] * hw[1]
    B = 3
    qkv = _rand((B * L, 3 * heads * d), cuda_dev, 1)
    qw = 1 + 0.1 * _rand((d,), cuda_dev, 2, dtype=torch.float32)
    kw = 1 + 0.1 * _rand((d,), cuda_dev, 3, dtype=torch.float32)
    ang = O.rope_table_2d(d, hw[0], hw[1]).to(cuda_dev)
    r = qkv.view(B, L, 3, heads, d)
    q_ref = O.apply_rope(O.rmsnorm(r[:, :, 0], qw), ang)
    k_ref = O.apply_rope(O.rmsnorm(r[:, :, 1], kw), ang)
    v_ref = r[:, :, 2].clone()
    got = ops.qknorm_rope_(qkv.clone(), qw, kw, rope_cos_sin(d, hw[0], hw[1]).to(cuda_dev), heads, d, L)
    g = got.view(B, L, 3, heads, d)
    assert rel_l2(g[:, :, 0].float(), q_ref) < 4e-3
    assert rel_l2(g[:, :, 1].float(), k_ref) < 4e-3
    assert torch.equal(g[:, :, 2], v_ref)


@pytest.mark.parametrize("heads,d,Lq,Lk2", [(16, 72, 256, 0), (4, 64, 100, 0), (2, 72, 1024, 0), (3, 64, 200, 77), (2, 72, 64, 128)])
def test_attention(cuda_dev, heads, d, Lq, Lk2):
    from deco_b200 import ops
    B, H = 2, heads * d
    qkv = _rand((B * Lq, 3 * H), cuda_dev, 1)
    q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
    k2 = v2 = None
    kk = k.reshape(B, Lq, heads, d).transpose(1, 2).float()
    vv = v.reshape(B, Lq, heads, d).transpose(1, 2).float()
    if Lk2:
        kv2 = _rand((B * Lk2, 2 * H), cuda_dev, 2)
        k2, v2 = kv2[:, :H], kv2[:, H:]
        kk = torch.cat([kk, k2.reshape(B, Lk2, heads, d).transpose(1, 2).float()], 2)
        vv = torch.cat([vv, v2.reshape(B, Lk2, heads, d).transpose(1, 2).float()], 2)
    qq = q.reshape(B, Lq, heads, d).transpose(1, 2).float()
    ref = F.scaled_dot_product_attention(qq, kk, vv).transpose(1, 2).reshape(B * Lq, H)
    got = ops.attention(q, k, v, B, heads, d, k2=k2, v2=v2)
    assert rel_l2(got.float(), ref) < 6e-3


@pytest.mark.parametrize("heads,d,hw,Lk2", [(16, 72, (16, 16), 0), (4, 64, (10, 10), 0), (2, 72, (32, 32), 0),
                                            (3, 64, (12, 17), 77), (2, 64, (32, 32), 128)])
def test_attention_after_qknorm_rope(cuda_dev, heads, d, hw, Lk2):
    """q_norm / k_norm / RoPE followed by attention (dit_c2i_DeCo.py:176-187; t2i: text keys get k_norm but no RoPE,
    dit_t2i_pixnerd.py:46-59) against the oracle's rmsnorm / apply_rope / fp32 SDPA."""
    from deco_b200 import ops
    from deco_b200.denoiser import rope_cos_sin
    L = hw[0] * hw[1]
    B, H = 2, heads * d
    qkv = _rand((B * L, 3 * H), cuda_dev, 1)
    qw = 1 + 0.1 * _rand((d,), cuda_dev, 2, dtype=torch.float32)
    kw = 1 + 0.1 * _rand((d,), cuda_dev, 3, dtype=torch.float32)
    ang = O.rope_table_2d(d, hw[0], hw[1]).to(cuda_dev)
    r = qkv.view(B, L, 3, heads, d)
    qq = O.apply_rope(O.rmsnorm(r[:, :, 0], qw), ang).to(bf16).float().transpose(1, 2)
    kk = O.apply_rope(O.rmsnorm(r[:, :, 1], kw), ang).to(bf16).float().transpose(1, 2)
    vv = r[:, :, 2].float().transpose(1, 2)
    k2 = v2 = None
    if Lk2:
        kv2 = _rand((B * Lk2, 2 * H), cuda_dev, 4)
        k2, v2 = kv2[:, :H], kv2[:, H:]
        kk = torch.cat([kk, O.rmsnorm(k2.reshape(B, Lk2, heads, d), kw).to(bf16).float().transpose(1, 2)], 2)
        vv = torch.cat([vv, v2.reshape(B, Lk2, heads, d).transpose(1, 2).float()], 2)
    ref = F.scaled_dot_product_attention(qq, kk, vv).transpose(1, 2).reshape(B * L, H)
    rope = rope_cos_sin(d, hw[0], hw[1]).to(cuda_dev)
    pre = ops.qknorm_rope_(qkv.clone(), qw, kw, rope, heads, d, L)
    if Lk2:
        # text keys: k_norm without RoPE = the same kernel with an identity rotation table
        ident = torch.stack([torch.ones(Lk2, d // 2), torch.zeros(Lk2, d // 2)], -1).contiguous().to(cuda_dev)
        kvq = torch.cat([kv2[:, :H], kv2], 1).contiguous()          # [q-slot (unused) | k | v]
        pre2 = ops.qknorm_rope_(kvq, kw, kw, ident, heads, d, Lk2)
        k2, v2 = pre2[:, H:2 * H], pre2[:, 2 * H:]
    got = ops.attention(pre[:, :H], pre[:, H:2 * H], pre[:, 2 * H:], B, heads, d, k2=k2, v2=v2)
    assert rel_l2(got.float(), ref) < 6e-3


def test_attention_running_max_rescale(cuda_dev):
    """Keys whose scores grow block by block force the lazy exponent-reference update (and the O rescale through
    tensor memory) in every 128-key block; one huge outlier key in the last block dominates the softmax."""
    from deco_b200 import ops
    B, heads, d, L = 1, 2, 64, 640
    H = heads * d
    g = torch.Generator().manual_seed(7)
    q = torch.randn((B * L, H), generator=g)
    k = torch.randn((B * L, H), generator=g) * (1 + torch.arange(L).view(-1, 1) // 128 * 3.0)   # block j: scale 1 + 3j
    v = torch.randn((B * L, H), generator=g)
    k[L - 5] *= 4.0
    q, k, v = (t.to(device=cuda_dev, dtype=bf16) for t in (q, k, v))
    sp = lambda t: t.reshape(B, L, heads, d).transpose(1, 2).float()
    ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B * L, H)
    got = ops.attention(q, k, v, B, heads, d)
    assert torch.isfinite(got.float()).all()
    assert rel_l2(got.float(), ref) < 6e-3


def test_silu_add_rows(cuda_dev):
    from deco_b200 import ops
    x, row = _rand((64, 256), cuda_dev, 1), _rand((4, 256), cuda_dev, 2)
    ref = F.silu((x + row.repeat_interleave(16, 0)).float())
    assert rel_l2(ops.silu_add_rows(x, row, 16).float(), ref) < 4e-3
    assert rel_l2(ops.silu_add_rows(x.float(), row, 16).float(), ref) < 4e-3   # fp32 residual stream input
    assert rel_l2(ops.silu_add_rows(x, row, 16, out=x).float(), ref) < 4e-3   # in place


@pytest.mark.parametrize("net_dtype", [bf16, torch.float32])
def test_cfg_step(cuda_dev, net_dtype):
    from deco_b200 import ops
    x = _rand((3, 3, 16, 16), cuda_dev, 1, dtype=torch.float32)
    out = _rand((6, 3, 16, 16), cuda_dev, 2, dtype=net_dtype)
    p1 = _rand((3, 3, 16, 16), cuda_dev, 3, dtype=torch.float32)
    p2 = _rand((3, 3, 16, 16), cuda_dev, 4, dtype=torch.float32)
    u, c = out.float().chunk(2)
    pred = u + 3.2 * (c - u)
    x1, pr, v, u8 = ops.cfg_step(x, out, 3.2, 0.01, want_pred=True, want_v=True, want_u8=True)
    assert torch.allclose(pr, pred, atol=1e-6) and torch.allclose(x1, x + 0.01 * pred, atol=1e-6)
    assert torch.equal(u8, O.fp2uint8(x1))
    x2, _, v2, _ = ops.cfg_step(x, out, 1.0, 0.5, c0=0.25, prev=(p1, p2), coeffs=(-0.5, 1.25), want_v=True)
    vr = 0.25 * c + -0.5 * p1 + 1.25 * p2
    assert torch.allclose(v2, vr, atol=1e-5) and torch.allclose(x2, x + 0.5 * vr, atol=1e-5)
    z = torch.tensor([-1.5, -1.0, -0.999, 0.0, 0.5, 0.996, 1.0, 3.0], device=cuda_dev)
    assert torch.equal(ops.fp2uint8(z), O.fp2uint8(z))


# ------------------------------------------------------------------------------------------------ fused-epilogue GEMMs
def _fp32(fn):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return fn()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("M,N,K,L,mode", [
    (512, 1152, 1152, 256, "full"), (1000, 1152, 3072, 250, "full"), (96, 576, 576, 16, "full"),
    (300, 256, 256, 100, "full"), (512, 1536, 6144, 128, "full"), (640, 1024, 2736, 64, "full"),
    (512, 1152, 768, 256, "embed"), (512, 1152, 1152, 256, "last"),
])
def test_gemm_stream(cuda_dev, M, N, K, L, mode):
    """FE_STREAM: out = [resid + gate *](A.W^T + bias) in place, row sums of squares, pre-modulated bf16 copy."""
    from deco_b200 import ops
    nimg = (M + L - 1) // L
    a, w = _rand((M, K), cuda_dev, 1), _rand((N, K), cuda_dev, 2, K ** -0.5)
    bias = _rand((N,), cuda_dev, 3, 0.1, torch.float32)
    resid = _rand((M, N), cuda_dev, 4, 1.5, dtype=torch.float32)
    gate = _rand((nimg, 3 * N), cuda_dev, 5)[:, N:2 * N]
    nscale = _rand((nimg, 3 * N), cuda_dev, 6, 0.5)[:, :N]
    nw = 1 + 0.1 * _rand((N,), cuda_dev, 7, dtype=torch.float32)
    rep = lambda t: t.float().repeat_interleave(L, 0)[:M]   # noqa: E731
    lin = _fp32(lambda: a.float() @ w.float().t() + bias)
    parts = ops.gemm_stream_parts(N, K)
    ssq = torch.full((parts, M), -1.0, device=cuda_dev)
    if mode == "embed":
        out = torch.empty((M, N), device=cuda_dev)
        xg = torch.empty((M, N), device=cuda_dev, dtype=bf16)
        ops.gemm_stream(a, w, bias, out, rows_per_image=L, next_w=nw, next_scale=nscale, xg=xg, ssq=ssq)
        ref = lin
    elif mode == "last":
        out = resid.clone()
        xg = None
        ops.gemm_stream(a, w, bias, out, resid=out, gate=gate, rows_per_image=L, ssq=None)
        ref = resid + rep(gate) * lin
    else:
        out = resid.clone()
        xg = torch.empty((M, N), device=cuda_dev, dtype=bf16)
        ops.gemm_stream(a, w, bias, out, resid=out, gate=gate, rows_per_image=L, next_w=nw, next_scale=nscale, xg=xg, ssq=ssq)
        ref = resid + rep(gate) * lin
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 3e-3
    if mode != "last":
        assert rel_l2(ssq.sum(0), (out.double() ** 2).sum(1).float()) < 1e-5
        assert rel_l2(xg.float(), out * nw * (1 + rep(nscale))) < 4e-3


@pytest.mark.parametrize("axial", [True, False])
@pytest.mark.parametrize("heads,d,hw,B,K,nseg,normed", [
    (16, 72, (16, 16), 2, 1152, 3, True), (8, 72, (4, 4), 3, 576, 3, True), (4, 64, (10, 10), 2, 256, 3, True),
    (24, 64, (8, 8), 2, 1536, 3, True), (4, 64, (3, 8), 2, 256, 2, False), (16, 64, (4, 4), 5, 1024, 3, True),
    (4, 72, (32, 32), 1, 576, 3, True), (2, 64, (8, 16), 2, 256, 3, True), (3, 64, (4, 8), 2, 192, 3, True),
    (5, 72, (4, 4), 2, 360, 3, True),
])
def test_gemm_norm_qkv(cuda_dev, heads, d, hw, B, K, nseg, normed, axial):
    """FE_NORM_QKV against Linear(modulate(RMSNorm(x))) -> q_norm/k_norm -> RoPE evaluated in fp32 torch.
    nseg=2, normed=False is the t2i kv_y projection: k-norm only on segment 0, no RoPE, no input normalisation."""
    from deco_b200 import ops
    from deco_b200.denoiser import rope_cos_sin
    L = hw[0] * hw[1]
    M, H = B * L, heads * d
    N = nseg * H
    w = _rand((N, K), cuda_dev, 2, K ** -0.5)
    qw = 1 + 0.1 * _rand((d,), cuda_dev, 3, dtype=torch.float32)
    kw = 1 + 0.1 * _rand((d,), cuda_dev, 4, dtype=torch.float32)
    out = torch.empty((M, N), device=cuda_dev, dtype=bf16)
    if normed:
        x = _rand((M, K), cuda_dev, 1, 3.0, dtype=torch.float32)
        nw = 1 + 0.1 * _rand((K,), cuda_dev, 5, dtype=torch.float32)
        mod = _rand((B, 2 * K), cuda_dev, 6, 0.5)
        sh, sc = mod[:, :K], mod[:, K:]
        xg = (x * nw * (1 + sc.float().repeat_interleave(L, 0))).to(bf16)
        # partial sums of squares in 3 parts, as the producing GEMM would leave them
        ssq = torch.stack([(x[:, i::3] ** 2).sum(1) for i in range(3)]).contiguous()
        shw = _fp32(lambda: sh.float() @ w.float().t()).contiguous()
        rope = rope_cos_sin(d, hw[0], hw[1]).to(cuda_dev)
        ops.gemm_norm_qkv(xg, w, out, L, heads, d, seg_w=(qw, kw, None), rope_mask=3, rope=rope,
                          rope_tokens_per_row=(hw[1] if axial else 0), ssq=ssq, norm_hidden=K, shw=shw)
        h = O.modulate(O.rmsnorm(x, nw), sh.float().repeat_interleave(L, 0), sc.float().repeat_interleave(L, 0))
        y = _fp32(lambda: h @ w.float().t()).view(B, L, 3, heads, d)
        ang = O.rope_table_2d(d, hw[0], hw[1]).to(cuda_dev)
        ref = torch.stack([O.apply_rope(O.rmsnorm(y[:, :, 0], qw), ang), O.apply_rope(O.rmsnorm(y[:, :, 1], kw), ang),
                           y[:, :, 2]], dim=2)
    else:
        a = _rand((M, K), cuda_dev, 1)
        ops.gemm_norm_qkv(a, w, out, L, heads, d, seg_w=(kw, None), rope_mask=0)
        y = _fp32(lambda: a.float() @ w.float().t()).view(B, L, 2, heads, d)
        ref = torch.stack([O.rmsnorm(y[:, :, 0], kw), y[:, :, 1]], dim=2)
    torch.cuda.synchronize()
    g = out.view(ref.shape).float()
    for s in range(nseg):
        assert rel_l2(g[:, :, s], ref[:, :, s]) < 6e-3, s
    if normed and d == 72:
        # the padded layout the sampling path uses: heads at a pitch of 80 columns, columns 72..79 zero, values identical;
        # and the attention kernel reads it through its pitched tensor maps with the same result as from the dense layout
        outp = torch.full((M, nseg * heads * 80), 7.0, device=cuda_dev, dtype=bf16)
        ops.gemm_norm_qkv(xg, w, outp, L, heads, d, seg_w=(qw, kw, None), rope_mask=3, rope=rope,
                          rope_tokens_per_row=(hw[1] if axial else 0), ssq=ssq, norm_hidden=K, shw=shw, out_head_pitch=80)
        gp = outp.view(M, nseg, heads, 80)
        assert torch.equal(gp[..., :72].reshape(M, N), out) and float(gp[..., 72:].abs().max()) == 0.0
        Hd, Hp = heads * d, heads * 80
        o_dense = ops.attention(out[:, :Hd], out[:, Hd:2 * Hd], out[:, 2 * Hd:], B, heads, d)
        o_pitch = ops.attention(outp[:, :Hp], outp[:, Hp:2 * Hp], outp[:, 2 * Hp:], B, heads, d, head_pitch=80)
        assert torch.equal(o_dense, o_pitch)


@pytest.mark.parametrize("M,F_,K,L", [(512, 3072, 1152, 256), (300, 592, 576, 100), (640, 2736, 1024, 64), (256, 6144, 1536, 128)])
def test_gemm_norm_swiglu(cuda_dev, M, F_, K, L):
    from deco_b200 import ops
    from deco_b200.denoiser import interleave_w13
    B = (M + L - 1) // L
    x = _rand((M, K), cuda_dev, 1, 3.0, dtype=torch.float32)
    nw = 1 + 0.1 * _rand((K,), cuda_dev, 5, dtype=torch.float32)
    mod = _rand((B, 2 * K), cuda_dev, 6, 0.5)
    sh, sc = mod[:, :K], mod[:, K:]
    rep = lambda t: t.float().repeat_interleave(L, 0)[:M]   # noqa: E731
    xg = (x * nw * (1 + rep(sc))).to(bf16)
    ssq = torch.stack([(x[:, i::2] ** 2).sum(1) for i in range(2)]).contiguous()
    w1, w3 = _rand((F_, K), cuda_dev, 2, K ** -0.5), _rand((F_, K), cuda_dev, 3, K ** -0.5)
    Fp = (F_ + 15) // 16 * 16
    w13 = interleave_w13(w1, w3, Fp)
    shw = _fp32(lambda: sh.float() @ w13.float().t()).contiguous()
    out = torch.empty((M, Fp), device=cuda_dev, dtype=bf16)
    ops.gemm_norm_swiglu(xg, w13, out, L, ssq=ssq, norm_hidden=K, shw=shw)
    h = O.modulate(O.rmsnorm(x, nw), rep(sh), rep(sc))
    ref = _fp32(lambda: F.silu(h @ w1.float().t()) * (h @ w3.float().t()))
    torch.cuda.synchronize()
    assert rel_l2(out[:, :F_].float(), ref) < 6e-3
    if Fp > F_:
        assert float(out[:, F_:].float().abs().max()) == 0.0


def _decoder_module(cuda_dev, R=3, H=256):
    from helpers import build_module
    cfg = O.DenoiserCfg(num_groups=4, hidden_size=H, num_blocks=R + 1, num_cond_blocks=1, num_classes=10)
    m, P = build_module(cfg, cuda_dev)
    return cfg, m, {k: v.to(cuda_dev) for k, v in P.items()}


@pytest.mark.parametrize("B,res,R", [(1, 16, 3), (3, 48, 3), (2, 64, 1), (8, 256, 3), (5, 128, 6)])
def test_pixel_decoder_tc_vs_oracle_and_legacy(cuda_dev, B, res, R):
    """csrc/decoder_tc.cu (tcgen05 / TMEM, folded weights, SiLU via tanh) against the oracle's pixel decoder
    (dit_c2i_DeCo.py:212-248, :313-332, :395-415) on the same condition s, and against the register-resident mma.sync kernel
    it replaces.  Sizes: one token (2 tiles), ragged token counts, 1 / 3 / 6 res blocks, 2048 tokens = 28 tiles per CTA (every
    slot pipelines several tiles, both condition buffers recycle)."""
    from deco_b200 import ops
    cfg, m, Pd = _decoder_module(cuda_dev, R)
    P = m.prepare(cuda_dev)
    g = torch.Generator().manual_seed(B * 1000 + res)
    L = (res // 16) ** 2
    x = torch.randn(B, 3, res, res, generator=g).to(cuda_dev)
    s = (torch.randn(B * L, cfg.hidden_size, generator=g) * 0.7).to(cuda_dev).to(torch.bfloat16)
    ref = O.denoiser_forward(Pd, cfg, x, torch.zeros(B, device=cuda_dev), torch.zeros(B, dtype=torch.long, device=cuda_dev),
                             s=s.float().view(B, L, -1))
    ysilu = ops.gemm(s, P["wcond"], P["bcond"], ops.EPI_BIAS_SILU)
    for odt in (torch.float32, torch.bfloat16):
        got = ops.pixel_decoder_tc(x, ysilu, P["blob_tc"], 16, 32, R, out_dtype=odt)
        e = rel_l2(got.float(), ref)
        assert e < 6e-3, (str(odt), e)
    ycond = ops.gemm(s, P["wcond"], P["bcond"], ops.EPI_BIAS)
    old = ops.pixel_decoder(x, ycond, P["blob"], P["postab"], 16, 32, R, out_dtype=torch.float32)
    got = ops.pixel_decoder_tc(x, ysilu, P["blob_tc"], 16, 32, R, out_dtype=torch.float32)
    print(f"decoder_tc B={B} res={res} R={R}: rel-L2 vs oracle {rel_l2(got, ref):.3e} (legacy kernel {rel_l2(old, ref):.3e}), "
          f"tc vs legacy {rel_l2(got, old):.3e}")
    assert rel_l2(got, old) < 8e-3
    # guard band: nothing outside the output tensor was written
    big = torch.full((B * 3 * res * res + 512,), 7.0, device=cuda_dev)
    from deco_b200._lib import call, ptr
    call("deco_pixel_decoder_tc", ptr(x), ptr(ysilu), ptr(P["blob_tc"]), big[256:].data_ptr(), 0, B, res, res, 16, 32, R, 0,
         None, 0.0, 0.0, 0.0, 0.0, None, None, None, None, None, torch.cuda.current_stream().cuda_stream)
    assert torch.equal(big[256:-256].view(B, 3, res, res), got)
    assert float((big[:256] - 7).abs().max()) == 0 and float((big[-256:] - 7).abs().max()) == 0


@pytest.mark.parametrize("B,res,order2", [(2, 32, False), (3, 64, True), (16, 128, True)])
def test_pixel_decoder_tc_fused_sampler_step(cuda_dev, B, res, order2):
    """pair mode of csrc/decoder_tc.cu: decoder of the rows [uncond || cond] + guidance (base/guidance.py:3-6) + the
    Euler / Adams update (sampling.py:100-104, adam_sampling.py:109-117) + fp2uint8, against the same decoder's fp32 output
    pushed through the reference formulas; host scalars and the device-table form agree bit for bit; in-place update."""
    from deco_b200 import ops
    cfg, m, _ = _decoder_module(cuda_dev, 3)
    P = m.prepare(cuda_dev)
    g = torch.Generator().manual_seed(res + B)
    L = (res // 16) ** 2
    x = torch.randn(B, 3, res, res, generator=g).to(cuda_dev)
    s = (torch.randn(2 * B * L, cfg.hidden_size, generator=g) * 0.7).to(cuda_dev).to(torch.bfloat16)
    ysilu = ops.gemm(s, P["wcond"], P["bcond"], ops.EPI_BIAS_SILU)
    out = ops.pixel_decoder_tc(torch.cat([x, x]), ysilu, P["blob_tc"], 16, 32, 3, out_dtype=torch.float32)
    gd, dt, c0, c1 = 3.2, 0.013, (1.5 if order2 else 1.0), (-0.5 if order2 else 0.0)
    p1 = torch.randn(x.shape, generator=g).to(cuda_dev) if order2 else None
    u, c = out[:B], out[B:]
    pred = u + gd * (c - u)
    v = c0 * pred + (c1 * p1 if order2 else 0)
    x_ref = x + dt * v
    pred_o = torch.empty_like(x)
    u8_o = torch.empty(x.shape, dtype=torch.uint8, device=cuda_dev)
    xo = ops.pixel_decoder_tc_step(x, ysilu, P["blob_tc"], 16, 32, 3, g=gd, dt=dt, c0=c0, c1=c1, p1=p1, pred_out=pred_o, u8_out=u8_o)
    assert rel_l2(xo, x_ref) < 2e-6 and rel_l2(pred_o, pred) < 2e-6
    assert int((u8_o.int() - O.fp2uint8(xo).int()).abs().max()) == 0
    dev = torch.tensor([gd, dt, c0, c1, 0, 0, 0.5, 0], dtype=torch.float32, device=cuda_dev)
    x2 = x.clone()
    p2 = p1.clone() if order2 else None
    ops.pixel_decoder_tc_step(x2, ysilu, P["blob_tc"], 16, 32, 3, dev=dev, p1=p2, x_out=x2, pred_out=p2)     # in place
    assert torch.equal(x2, xo)
    if order2:
        assert torch.equal(p2, pred_o)
    # Heun corrector form: the network saw x (= x_hat), the update starts from another state x_base
    xb = torch.randn(x.shape, generator=g).to(cuda_dev)
    xh = ops.pixel_decoder_tc_step(x, ysilu, P["blob_tc"], 16, 32, 3, g=gd, dt=dt, c0=c0, c1=c1, p1=p1, x_base=xb)
    assert rel_l2(xh, xb + dt * v) < 2e-6
