// Backward of the memory-bound pieces of the DiT block (training step, SURVEY.md 8(f) rank 1 / BASELINE configs[3]).
// The dense contractions of the backward pass (dgrad, wgrad) run on the tcgen05 GEMM of gemm_tcgen05.cu; this file holds
// the glue around them: operand transposes for wgrad, gate / residual, SwiGLU, RMSNorm + modulate, per-head q/k RMSNorm +
// RoPE, silu(t + s), embedding and bias reductions.  All HBM-bound, one pass each, fp32 math, bf16 operands for the GEMMs.
//
// Differentiates (reference, /root/reference/src/models/transformer/dit_c2i_DeCo.py):
//   :11-12 modulate, :94-99 RMSNorm.forward, :112-114 FeedForward.forward, :134-145 apply_rotary_emb,
//   :178-180 q_norm / k_norm, :206-210 FlattenDiTBlock.forward (gated residuals), :494 silu(t + y), :499 silu(t + s).
#include "common.cuh"
#include <cstdlib>

namespace deco {

__device__ __forceinline__ float dsilu_f(float x) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-x));
    return s * fmaf(x, 1.0f - s, 1.0f);
}

template <typename T> struct Pair;
template <> struct Pair<float> {
    static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
};
template <> struct Pair<__nv_bfloat16> {
    static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) { return unpack_bf2(*reinterpret_cast<const uint32_t*>(p)); }
};

// ---------------------------------------------------------------- dst[c][r] = bf16(src[r][c]), rows r >= R zero-filled up to Rp
// 64 x 64 tiles; 32-bit global accesses on both sides (two bf16 / one float2 per lane).  C even.
template <typename TIn>
__global__ void __launch_bounds__(256) transpose_cast_kernel(const TIn* __restrict__ src, long long lds,
                                                             __nv_bfloat16* __restrict__ dst, long long ldd,
                                                             int R, int C, int Rp)
{
    __shared__ __nv_bfloat16 tile[64][66];
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 64; i += 8) {
        const int r = r0 + i, c = c0 + 2 * tx;
        float2 v = make_float2(0.f, 0.f);
        if (r < R && c < C) v = Pair<TIn>::ld(src + (long long)r * lds + c);
        tile[i][2 * tx] = f2bf(v.x);
        tile[i][2 * tx + 1] = f2bf(v.y);
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 8) {
        const int c = c0 + i, r = r0 + 2 * tx;
        if (c < C && r < Rp) {
            __nv_bfloat162 o;
            o.x = tile[2 * tx][i];
            o.y = tile[2 * tx + 1][i];
            *reinterpret_cast<__nv_bfloat162*>(dst + (long long)c * ldd + r) = o;
        }
    }
}

// ---------------------------------------------------------------- out[c] += sum_r x[r][c]   (bias gradients)
template <typename TIn>
__global__ void __launch_bounds__(256) colsum_kernel(const TIn* __restrict__ x, long long ldx, float* __restrict__ out,
                                                     long long M, int N, int rows_per_block)
{
    __shared__ float2 part[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 64 + 2 * tx;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float2 acc = make_float2(0.f, 0.f);
    if (c < N)
        for (long long r = r0 + ty; r < r1; r += 8) {
            const float2 v = Pair<TIn>::ld(x + r * ldx + c);
            acc.x += v.x; acc.y += v.y;
        }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < N) {
        for (int j = 1; j < 8; ++j) { acc.x += part[j][tx].x; acc.y += part[j][tx].y; }
        atomicAdd(out + c, acc.x);
        atomicAdd(out + c + 1, acc.y);
    }
}

// ---------------------------------------------------------------- out[r][c] = s[r][c] + gate[r / L][c] * a[r][c]
__global__ void __launch_bounds__(256) gate_residual_kernel(const float* s, const __nv_bfloat16* __restrict__ a,
                                                            const __nv_bfloat16* __restrict__ gate, long long gate_stride,
                                                            float* out, int L, long long M, int Hd)
{
    const int npair = Hd >> 1;
    const long long total = M * npair;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / npair;
        const int c = 2 * (int)(i % npair);
        const float2 av = Pair<__nv_bfloat16>::ld(a + r * Hd + c);
        const float2 g = Pair<__nv_bfloat16>::ld(gate + (r / L) * gate_stride + c);
        float2 v = *reinterpret_cast<const float2*>(s + r * Hd + c);
        v.x = fmaf(g.x, av.x, v.x); v.y = fmaf(g.y, av.y, v.y);
        *reinterpret_cast<float2*>(out + r * Hd + c) = v;
    }
}

// ---------------------------------------------------------------- backward of s_out = s_in + gate[b] * a
//   da = gate * ds (bf16), dgate[b] += sum_rows ds * a, dbias += sum_rows da (optional: a = x.W^T + bias)
// Block = RB rows of one image (RB divides L); thread owns column pairs tid + 256 k.
constexpr int kMaxPairIters = 4;   // hidden <= 2048
// thread owns 4 consecutive columns (blockDim = ceil32(hidden / 4)): 16-byte stream loads, 8-byte bf16 loads / stores
__global__ void __launch_bounds__(512) gate_bwd_kernel(const float* __restrict__ ds, const __nv_bfloat16* __restrict__ a,
                                                       const __nv_bfloat16* __restrict__ gate, long long gate_stride,
                                                       __nv_bfloat16* __restrict__ da, float* __restrict__ dgate,
                                                       long long dgate_stride, float* __restrict__ usum,
                                                       int L, int RB, int Hd)
{
    // usum [B, Hd] (optional): per-image column sums of ds, from which gate_bias_finalize_kernel forms the bias gradient
    // sum_b gate[b] * usum[b] -- adding every block's partial straight into dbias[Hd] put 1024 atomics on each of 1152
    // addresses and tripled the kernel time (19.9 -> 54.5 us)
    const int c = 4 * threadIdx.x;
    if (c >= Hd) return;
    const long long r0 = (long long)blockIdx.x * RB;
    const long long b = r0 / L;
    float g[4];
    {
        const uint2 q = *reinterpret_cast<const uint2*>(gate + b * gate_stride + c);
        const float2 g0 = unpack_bf2(q.x), g1 = unpack_bf2(q.y);
        g[0] = g0.x; g[1] = g0.y; g[2] = g1.x; g[3] = g1.y;
    }
    float accg[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int i = 0; i < RB; ++i) {
        const long long r = r0 + i;
        const float4 d = *reinterpret_cast<const float4*>(ds + r * Hd + c);
        const uint2 q = *reinterpret_cast<const uint2*>(a + r * Hd + c);
        const float2 a0 = unpack_bf2(q.x), a1 = unpack_bf2(q.y);
        const float o0 = g[0] * d.x, o1 = g[1] * d.y, o2 = g[2] * d.z, o3 = g[3] * d.w;
        *reinterpret_cast<uint2*>(da + r * Hd + c) = make_uint2(pack_bf2(o0, o1), pack_bf2(o2, o3));
        accg[0] = fmaf(d.x, a0.x, accg[0]); accg[1] = fmaf(d.y, a0.y, accg[1]);
        accg[2] = fmaf(d.z, a1.x, accg[2]); accg[3] = fmaf(d.w, a1.y, accg[3]);
        accb[0] += d.x; accb[1] += d.y; accb[2] += d.z; accb[3] += d.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        atomicAdd(dgate + b * dgate_stride + c + e, accg[e]);
        if (usum) atomicAdd(usum + b * Hd + c + e, accb[e]);
    }
}

// dbias[c] += sum_b gate[b][c] * usum[b][c];  block = 32 columns x 32 image lanes
__global__ void __launch_bounds__(1024) gate_bias_finalize_kernel(const __nv_bfloat16* __restrict__ gate, long long gate_stride,
                                                                  const float* __restrict__ usum, float* __restrict__ dbias,
                                                                  int B, int Hd)
{
    __shared__ float red[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < Hd)
        for (int b = threadIdx.y; b < B; b += 32) acc = fmaf(bf2f(gate[b * gate_stride + c]), usum[(size_t)b * Hd + c], acc);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < Hd) {
        for (int j = 1; j < 32; ++j) acc += red[j][threadIdx.x];
        dbias[c] += acc;
    }
}

// ---------------------------------------------------------------- backward of out = silu(x + row[b])  (dit_c2i_DeCo.py:499)
//   dx = dout * silu'(x + row) (fp32), drow[b] += sum_rows dx
__global__ void __launch_bounds__(256) silu_add_rows_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                                const float* __restrict__ x,
                                                                const __nv_bfloat16* __restrict__ rowv,
                                                                float* __restrict__ dx, float* __restrict__ drow,
                                                                int L, int RB, int Hd)
{
    const long long r0 = (long long)blockIdx.x * RB;
    const long long b = r0 / L;
    const int npair = Hd >> 1;
    float2 rv[kMaxPairIters], acc[kMaxPairIters];
#pragma unroll
    for (int k = 0; k < kMaxPairIters; ++k) {
        const int p = threadIdx.x + k * 256;
        rv[k] = p < npair ? Pair<__nv_bfloat16>::ld(rowv + b * Hd + 2 * p) : make_float2(0.f, 0.f);
        acc[k] = make_float2(0.f, 0.f);
    }
#pragma unroll 4
    for (int i = 0; i < RB; ++i) {
        const long long r = r0 + i;
#pragma unroll
        for (int k = 0; k < kMaxPairIters; ++k) {
            const int p = threadIdx.x + k * 256;
            if (p < npair) {
                const float2 d = Pair<__nv_bfloat16>::ld(dout + r * Hd + 2 * p);
                const float2 xv = *reinterpret_cast<const float2*>(x + r * Hd + 2 * p);
                const float2 o = make_float2(d.x * dsilu_f(xv.x + rv[k].x), d.y * dsilu_f(xv.y + rv[k].y));
                *reinterpret_cast<float2*>(dx + r * Hd + 2 * p) = o;
                acc[k].x += o.x; acc[k].y += o.y;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxPairIters; ++k) {
        const int p = threadIdx.x + k * 256;
        if (p < npair) { atomicAdd(drow + b * Hd + 2 * p, acc[k].x); atomicAdd(drow + b * Hd + 2 * p + 1, acc[k].y); }
    }
}

// ---------------------------------------------------------------- SwiGLU on the interleaved [16 x w1 | 16 x w3] columns
// forward:  u[r][16 g + i] = silu(y[r][32 g + i]) * y[r][32 g + 16 + i]
// backward: dy[r][32 g + i] = du * b * silu'(a),  dy[r][32 g + 16 + i] = du * silu(a)
// One thread per 8 output columns.
template <bool BWD>
__global__ void __launch_bounds__(256) swiglu_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ du,
                                                     __nv_bfloat16* __restrict__ out, long long M, int Fp)
{
    const int nch = Fp >> 3;
    const long long total = M * nch;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long r = i / nch;
        const int ch = (int)(i % nch);           // 8 columns [8 ch, 8 ch + 8) of u
        const int g = ch >> 1, half = ch & 1;
        const __nv_bfloat16* ya = y + r * 2 * Fp + 32 * g + 8 * half;
        const uint4 qa = *reinterpret_cast<const uint4*>(ya), qb = *reinterpret_cast<const uint4*>(ya + 16);
        const uint32_t wa[4] = {qa.x, qa.y, qa.z, qa.w}, wb[4] = {qb.x, qb.y, qb.z, qb.w};
        if (!BWD) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 av = unpack_bf2(wa[e]), bv = unpack_bf2(wb[e]);
                o[e] = pack_bf2(silu_f(av.x) * bv.x, silu_f(av.y) * bv.y);
            }
            *reinterpret_cast<uint4*>(out + r * Fp + 8 * ch) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
            const uint4 qd = *reinterpret_cast<const uint4*>(du + r * Fp + 8 * ch);
            const uint32_t wd[4] = {qd.x, qd.y, qd.z, qd.w};
            uint32_t oa[4], ob[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 av = unpack_bf2(wa[e]), bv = unpack_bf2(wb[e]), d = unpack_bf2(wd[e]);
                oa[e] = pack_bf2(d.x * bv.x * dsilu_f(av.x), d.y * bv.y * dsilu_f(av.y));
                ob[e] = pack_bf2(d.x * silu_f(av.x), d.y * silu_f(av.y));
            }
            __nv_bfloat16* oy = out + r * 2 * Fp + 32 * g + 8 * half;
            *reinterpret_cast<uint4*>(oy) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
            *reinterpret_cast<uint4*>(oy + 16) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
        }
    }
}

// ---------------------------------------------------------------- backward of h = rms(x) * w * (1 + scale[b]) + shift[b]
//   dxh = dh * w * (1 + scale);  dx = r * dxh - x * r^3 * mean(dxh * x);  ds += dx   (the residual stream gradient)
//   dshift[b] += sum_rows dh;  dscale[b] += sum_rows dh * xhat * w;  dw += sum_rows dh * xhat * (1 + scale)
// Two streaming passes instead of one latency-bound one: (1) a warp per row reduces the two row scalars r and
// k2 = r^3 mean(dxh x) (low register count, full occupancy); (2) a thread per 4 columns applies them, updates ds in place and
// keeps the per-image column sums in registers (no row reduction left, so no barrier or shuffle in the loop).
template <int NV>
__global__ void __launch_bounds__(256) rmsnorm_bwd_rowstats_kernel(
    const __nv_bfloat16* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ w,
    const __nv_bfloat16* __restrict__ scale, long long mod_stride, float2* __restrict__ stats, int L, long long M, int Hd, float eps)
{
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M) return;
    const __nv_bfloat16* srow = scale + (r / L) * mod_stride;
    const int nch = Hd >> 2;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int c = lane + 32 * k;
        if (c < nch) {
            const float4 xv = *reinterpret_cast<const float4*>(x + r * Hd + 4 * c);
            const uint2 d2 = *reinterpret_cast<const uint2*>(dh + r * Hd + 4 * c);
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + 4 * c));
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(srow + 4 * c));
            const float2 d0 = unpack_bf2(d2.x), d1 = unpack_bf2(d2.y), q0 = unpack_bf2(q.x), q1 = unpack_bf2(q.y);
            s1 = fmaf(xv.x, xv.x, fmaf(xv.y, xv.y, fmaf(xv.z, xv.z, fmaf(xv.w, xv.w, s1))));
            s2 = fmaf(d0.x * w4.x * (1.0f + q0.x), xv.x, fmaf(d0.y * w4.y * (1.0f + q0.y), xv.y,
                 fmaf(d1.x * w4.z * (1.0f + q1.x), xv.z, fmaf(d1.y * w4.w * (1.0f + q1.y), xv.w, s2))));
        }
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) {
        const float rs = rsqrtf(s1 / (float)Hd + eps);
        stats[r] = make_float2(rs, rs * rs * rs * s2 / (float)Hd);
    }
}

// kGate: the gate_bwd_kernel work of the branch that is differentiated NEXT (its residual add consumed the same stream
// gradient) rides along: da = gate2 * ds_new, dgate2 += sum_rows ds_new * a, usum2 += sum_rows ds_new -- the updated
// stream gradient is used while it is still in registers instead of being re-read by a second kernel.
struct GateBwdArgs {
    const __nv_bfloat16* a;
    const __nv_bfloat16* gate;
    long long gate_stride;
    __nv_bfloat16* da;
    float* dgate;
    long long dgate_stride;
    float* usum;
};

// Grid: ONE resident wave (the launcher asks the occupancy calculator), block i owns the contiguous rows
// [M i / grid, M (i + 1) / grid) -- equal shares to within a row -- and walks them image by image, so the per-image column
// sums stay in registers until the image changes (at most a couple of atomic flushes per block).  The first version gave
// every block a fixed 8 / 16 rows: 512 blocks on 444 resident slots = 1.15 waves, the tail wave running a sixth full
// (ncu: 36 % of DRAM throughput, 47 us for 171 MB).
template <bool kGate>
__global__ void __launch_bounds__(512) rmsnorm_modulate_bwd_kernel(
    const __nv_bfloat16* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ w,
    const __nv_bfloat16* __restrict__ scale, long long mod_stride, const float2* __restrict__ stats, float* __restrict__ ds,
    float* __restrict__ tsum, float* __restrict__ dshift, long long dmod_stride,
    int L, long long M, int Hd, const GateBwdArgs G)
{
    // tsum [B, Hd]: per-image sums T = sum_rows dh * xhat; d scale = w * T and d w = sum_b (1 + scale[b]) * T are formed
    // by rmsnorm_bwd_finalize_kernel (one atomic target per (image, column) instead of 1024 blocks hammering dw[Hd])
    const int c = 4 * threadIdx.x;
    if (c >= Hd) return;
    const long long r_begin = M * blockIdx.x / gridDim.x, r_end = M * (blockIdx.x + 1) / gridDim.x;
    float wv[4];
    {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c));
        wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
    }
    for (long long seg = r_begin; seg < r_end;) {
        const long long b = seg / L;
        const long long seg_end = min(r_end, (b + 1) * L);
        float sc1[4];
        {
            const uint2 q = *reinterpret_cast<const uint2*>(scale + b * mod_stride + c);
            const float2 q0 = unpack_bf2(q.x), q1 = unpack_bf2(q.y);
            sc1[0] = 1.0f + q0.x; sc1[1] = 1.0f + q0.y; sc1[2] = 1.0f + q1.x; sc1[3] = 1.0f + q1.y;
        }
        float a_sh[4] = {0.f, 0.f, 0.f, 0.f}, a_t[4] = {0.f, 0.f, 0.f, 0.f};
        float g2[4] = {0.f, 0.f, 0.f, 0.f}, accg[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
        if (kGate) {
            const uint2 q = *reinterpret_cast<const uint2*>(G.gate + b * G.gate_stride + c);
            const float2 q0 = unpack_bf2(q.x), q1 = unpack_bf2(q.y);
            g2[0] = q0.x; g2[1] = q0.y; g2[2] = q1.x; g2[3] = q1.y;
        }
#pragma unroll 4
        for (long long r = seg; r < seg_end; ++r) {
            const float2 st = __ldg(stats + r);
            const float4 x4 = *reinterpret_cast<const float4*>(x + r * Hd + c);
            const uint2 d2 = *reinterpret_cast<const uint2*>(dh + r * Hd + c);
            float4* dp = reinterpret_cast<float4*>(ds + r * Hd + c);
            float4 d4 = *dp;
            const float2 d0 = unpack_bf2(d2.x), d1 = unpack_bf2(d2.y);
            const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, dv[4] = {d0.x, d0.y, d1.x, d1.y};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                o[e] = st.x * dv[e] * wv[e] * sc1[e] - xv[e] * st.y;
                a_sh[e] += dv[e];
                a_t[e] = fmaf(dv[e] * xv[e], st.x, a_t[e]);
            }
            d4.x += o[0]; d4.y += o[1]; d4.z += o[2]; d4.w += o[3];
            *dp = d4;
            if (kGate) {
                const uint2 qa = *reinterpret_cast<const uint2*>(G.a + r * Hd + c);
                const float2 a0 = unpack_bf2(qa.x), a1 = unpack_bf2(qa.y);
                *reinterpret_cast<uint2*>(G.da + r * Hd + c) = make_uint2(pack_bf2(g2[0] * d4.x, g2[1] * d4.y), pack_bf2(g2[2] * d4.z, g2[3] * d4.w));
                accg[0] = fmaf(d4.x, a0.x, accg[0]); accg[1] = fmaf(d4.y, a0.y, accg[1]);
                accg[2] = fmaf(d4.z, a1.x, accg[2]); accg[3] = fmaf(d4.w, a1.y, accg[3]);
                accb[0] += d4.x; accb[1] += d4.y; accb[2] += d4.z; accb[3] += d4.w;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            atomicAdd(dshift + b * dmod_stride + c + e, a_sh[e]);
            atomicAdd(tsum + b * Hd + c + e, a_t[e]);
            if (kGate) {
                atomicAdd(G.dgate + b * G.dgate_stride + c + e, accg[e]);
                if (G.usum) atomicAdd(G.usum + b * Hd + c + e, accb[e]);
            }
        }
        seg = seg_end;
    }
}

// dscale[b][c] += w[c] * T[b][c];  dw[c] += sum_b (1 + scale[b][c]) * T[b][c];  block = 32 columns x 32 image lanes
__global__ void __launch_bounds__(1024) rmsnorm_bwd_finalize_kernel(const float* __restrict__ tsum, const float* __restrict__ w,
                                                                    const __nv_bfloat16* __restrict__ scale, long long mod_stride,
                                                                    float* __restrict__ dscale, long long dmod_stride,
                                                                    float* __restrict__ dw, int B, int Hd)
{
    __shared__ float red[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < Hd) {
        const float wc = __ldg(w + c);
        for (int b = threadIdx.y; b < B; b += 32) {
            const float t = tsum[(size_t)b * Hd + c];
            dscale[b * dmod_stride + c] += wc * t;
            acc = fmaf(1.0f + bf2f(scale[b * mod_stride + c]), t, acc);
        }
    }
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < Hd) {
        for (int j = 1; j < 32; ++j) acc += red[j][threadIdx.x];
        dw[c] += acc;
    }
}

// ---------------------------------------------------------------- backward of per-head RMSNorm (+ RoPE), in place on the
// gradient buffer: g [M, g_stride] holds d(out) for the segment on entry and d(raw) on exit; raw [M, raw_stride] is the
// GEMM output the forward normalised.  forward (elementwise.cu qknorm_rope_kernel): n = raw * rs, a = w * n,
// out = rot(a).  G lanes share one (token, head) vector (3 x 24 elements for D = 72, 4 x 16 for D = 64): a lane keeps its
// 16-byte chunks packed and its share of the weight gradient in registers (~80 registers; the first version gave a thread the
// whole vector: 255 registers, 8 warps per SM, 27 us for 57 MB -- ncu: 17 % of DRAM throughput, 68 % of cycles without an
// eligible warp).  The two row scalars are reduced over the G lanes by shuffles; the weight gradient goes per vector group
// into shared memory once per block, is summed per column and leaves with one atomic per column and block.
template <int D> struct HeadGroup { static constexpr int G = (D == 72) ? 3 : 4; };

template <int D>
__global__ void __launch_bounds__(256, 3) headnorm_rope_bwd_kernel(__nv_bfloat16* __restrict__ g, long long g_stride,
                                                                const __nv_bfloat16* __restrict__ raw, long long raw_stride,
                                                                int col, const float* __restrict__ wv,
                                                                const float2* __restrict__ rope, float* __restrict__ dw,
                                                                long long M, int heads, int L, float eps)
{
    constexpr int G = HeadGroup<D>::G;
    constexpr int CPL = D / 8 / G;            // 16-byte chunks per lane
    constexpr int VPW = 32 / G;               // vectors per warp (D = 72: lanes 30, 31 idle)
    constexpr int GPB = 8 * VPW;              // vector groups per block
    static_assert(D % (8 * G) == 0, "head_dim must split into 16-byte chunks per lane");
    __shared__ __align__(16) float sw[D];
    __shared__ float sred[D][GPB + 1];
    for (int i = threadIdx.x; i < D; i += blockDim.x) sw[i] = wv[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, sub = lane % G;
    const bool lane_live = grp < VPW;
    const int base = grp * G;                 // first lane of this vector group
    float dwacc[8 * CPL];
#pragma unroll
    for (int e = 0; e < 8 * CPL; ++e) dwacc[e] = 0.f;
    const long long total = M * heads;
    const long long stride = (long long)gridDim.x * GPB;
    for (long long it0 = (long long)blockIdx.x * GPB + warp * VPW; it0 < total; it0 += stride) {      // warp-uniform trip count
        const long long item = it0 + grp;
        const bool live = lane_live && item < total;
        const long long tok = live ? item / heads : 0;
        const int head = live ? (int)(item % heads) : 0;
        __nv_bfloat16* gp = g + tok * g_stride + col + (long long)head * D;
        const __nv_bfloat16* rp = raw + tok * raw_stride + col + (long long)head * D;
        const float2* cs = rope ? rope + (long long)(tok % L) * (D / 2) : nullptr;
        uint4 rq[CPL], gq[CPL];
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int ch = sub + G * k;
            rq[k] = live ? *reinterpret_cast<const uint4*>(rp + ch * 8) : make_uint4(0u, 0u, 0u, 0u);
            gq[k] = live ? *reinterpret_cast<const uint4*>(gp + ch * 8) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t w4[4] = {rq[k].x, rq[k].y, rq[k].z, rq[k].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 a = unpack_bf2(w4[e]); ss = fmaf(a.x, a.x, fmaf(a.y, a.y, ss)); }
        }
        float tot = 0.f;
#pragma unroll
        for (int j = 0; j < G; ++j) tot += __shfl_sync(0xffffffffu, ss, (base + j) & 31);
        const float rs = rsqrtf(tot / (float)D + eps);
        // d a = rot^T(g); d w += d a * n; d n = d a * w (kept for the output pass); dot = mean(d n * n)
        float dn[8 * CPL];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int ch = sub + G * k;
            const uint32_t r4[4] = {rq[k].x, rq[k].y, rq[k].z, rq[k].w}, g4[4] = {gq[k].x, gq[k].y, gq[k].z, gq[k].w};
            float4 t01 = make_float4(1.f, 0.f, 1.f, 0.f), t23 = t01;
            if (cs && live) { t01 = __ldg(reinterpret_cast<const float4*>(cs + ch * 4)); t23 = __ldg(reinterpret_cast<const float4*>(cs + ch * 4 + 2)); }
            const float4 w03 = *reinterpret_cast<const float4*>(sw + ch * 8), w47 = *reinterpret_cast<const float4*>(sw + ch * 8 + 4);
            const float cc[4] = {t01.x, t01.z, t23.x, t23.z}, sn[4] = {t01.y, t01.w, t23.y, t23.w};
            const float ww[8] = {w03.x, w03.y, w03.z, w03.w, w47.x, w47.y, w47.z, w47.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = unpack_bf2(r4[e]), gg = unpack_bf2(g4[e]);
                const float da0 = gg.x * cc[e] + gg.y * sn[e], da1 = -gg.x * sn[e] + gg.y * cc[e];     // transpose of the rotation
                const float n0 = a.x * rs, n1 = a.y * rs;
                dwacc[8 * k + 2 * e] = fmaf(da0, n0, dwacc[8 * k + 2 * e]);
                dwacc[8 * k + 2 * e + 1] = fmaf(da1, n1, dwacc[8 * k + 2 * e + 1]);
                dn[8 * k + 2 * e] = da0 * ww[2 * e];
                dn[8 * k + 2 * e + 1] = da1 * ww[2 * e + 1];
                dot = fmaf(dn[8 * k + 2 * e], n0, fmaf(dn[8 * k + 2 * e + 1], n1, dot));
            }
        }
        float dtot = 0.f;
#pragma unroll
        for (int j = 0; j < G; ++j) dtot += __shfl_sync(0xffffffffu, dot, (base + j) & 31);
        const float k2 = rs * rs * dtot / (float)D;       // d raw = rs * d n - raw * rs^2 * mean(d n * n)
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int ch = sub + G * k;
            const uint32_t r4[4] = {rq[k].x, rq[k].y, rq[k].z, rq[k].w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = unpack_bf2(r4[e]);
                o[e] = pack_bf2(fmaf(rs, dn[8 * k + 2 * e], -a.x * k2), fmaf(rs, dn[8 * k + 2 * e + 1], -a.y * k2));
            }
            if (live) *reinterpret_cast<uint4*>(gp + ch * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    // weight gradient: this lane's columns, one slot per vector group of the block
    if (lane_live) {
#pragma unroll
        for (int k = 0; k < CPL; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) sred[(sub + G * k) * 8 + e][warp * VPW + grp] = dwacc[8 * k + e];
    }
    __syncthreads();
    if (threadIdx.x < D) {
        float a = 0.f;
        for (int i = 0; i < GPB; ++i) a += sred[threadIdx.x][i];
        atomicAdd(dw + threadIdx.x, a);
    }
}

// ---------------------------------------------------------------- backward of c = silu(temb + table[label])  (:493-494)
//   dpre = dc * silu'(pre);  dtemb += dpre;  dtable[label] += dpre
__global__ void cond_combine_bwd_kernel(const float* __restrict__ dc, const __nv_bfloat16* __restrict__ temb,
                                        const float* __restrict__ table, const long long* __restrict__ labels,
                                        float* __restrict__ dtemb, float* __restrict__ dtable, int B, int Hd, int num_rows)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Hd) return;
    const int b = i / Hd, h = i % Hd;
    long long lab = labels[b];
    if (lab < 0 || lab >= num_rows) lab = num_rows - 1;
    const float pre = bf2f(temb[i]) + __ldg(table + lab * Hd + h);
    const float d = dc[i] * dsilu_f(pre);
    dtemb[i] += d;
    atomicAdd(dtable + lab * Hd + h, d);
}

// ---------------------------------------------------------------- dz = dy * silu'(z)   (t_embedder.mlp[1], :55-57)
__global__ void silu_bwd_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dy,
                                __nv_bfloat16* __restrict__ dz, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dz[i] = f2bf(bf2f(dy[i]) * dsilu_f(bf2f(z[i])));
}

static inline unsigned bgrid(long long work, int threads) {
    long long b = (work + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// rows per block of the per-image reductions: a power of two <= 32 dividing L, halved until the grid has >= 2 CTAs per SM.
// Each block ends with one fp32 atomic per column and sum: fewer rows per block = more warps to hide latency but more
// atomics, and a REDG costs the SM ~1.3 clocks per lane: at 8 rows per block (4 CTAs per SM, the first setting) the norm
// backward issued 4.7 M of them per call, ~20 us of LSU time next to ~26 us of HBM time (train step 34.95 -> 34.73 ms).
static inline int rows_block(int L, long long M, int ctas_per_sm = -1) {
    static int env_ctas = -2;      // DECO_ROWS_BLOCK_CTAS overrides the target CTAs per SM (A/B measurements)
    if (env_ctas == -2) { const char* e = getenv("DECO_ROWS_BLOCK_CTAS"); env_ctas = e ? atoi(e) : -1; }
    if (ctas_per_sm < 0) ctas_per_sm = env_ctas > 0 ? env_ctas : 2;
    int rb = 32;
    while (rb > 1 && (L % rb || M / rb < (long long)ctas_per_sm * kNumSMs)) rb >>= 1;
    return rb;
}

}  // namespace deco

extern "C" int deco_transpose_cast(const void* src, int src_is_f32, long long lds, void* dst_bf16, long long ldd,
                                   int R, int C, int Rp, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(src && dst_bf16 && R > 0 && C > 0 && Rp >= R, "transpose_cast: bad arguments");
    DECO_CHECK_ARG(C % 2 == 0 && lds % 2 == 0 && ldd % 2 == 0 && Rp % 2 == 0 && ldd >= Rp,
                   "transpose_cast: C, Rp and leading dimensions must be even (C=%d Rp=%d lds=%lld ldd=%lld)", C, Rp, lds, ldd);
    dim3 grid((C + 63) / 64, (Rp + 63) / 64);
    DECO_CHECK_ARG(grid.y <= 65535, "transpose_cast: too many rows (%d)", Rp);
    if (src_is_f32)
        transpose_cast_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, lds, (__nv_bfloat16*)dst_bf16, ldd, R, C, Rp);
    else
        transpose_cast_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst_bf16, ldd, R, C, Rp);
    DECO_CHECK_LAUNCH("transpose_cast_kernel");
    return DECO_OK;
}

extern "C" int deco_colsum(const void* x, int x_is_f32, long long ldx, float* out_accum, long long M, int N, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(x && out_accum && M > 0 && N > 0 && N % 2 == 0 && ldx % 2 == 0, "colsum: bad arguments");
    const int rpb = 256;
    dim3 grid((N + 63) / 64, (unsigned)((M + rpb - 1) / rpb));
    DECO_CHECK_ARG(grid.y <= 65535, "colsum: too many rows");
    if (x_is_f32) colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, ldx, out_accum, M, N, rpb);
    else colsum_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, out_accum, M, N, rpb);
    DECO_CHECK_LAUNCH("colsum_kernel");
    return DECO_OK;
}

extern "C" int deco_gate_residual(const float* s, const void* a_bf16, const void* gate_bf16, long long gate_stride,
                                  float* out, int rows_per_image, long long M, int hidden, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(s && out && a_bf16 && gate_bf16 && M > 0 && hidden > 0 && hidden % 2 == 0 && rows_per_image > 0 &&
                   gate_stride % 2 == 0, "gate_residual: bad arguments");
    gate_residual_kernel<<<bgrid(M * (hidden / 2), 256), 256, 0, (cudaStream_t)stream>>>(
        s, (const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)gate_bf16, gate_stride, out, rows_per_image, M, hidden);
    DECO_CHECK_LAUNCH("gate_residual_kernel");
    return DECO_OK;
}

extern "C" int deco_gate_bwd(const float* ds, const void* a_bf16, const void* gate_bf16, long long gate_stride,
                             void* da_bf16, float* dgate_accum, long long dgate_stride, float* dbias_accum, float* img_ws,
                             int rows_per_image, long long M, int hidden, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(ds && a_bf16 && gate_bf16 && da_bf16 && dgate_accum, "gate_bwd: null pointer");
    DECO_CHECK_ARG(!dbias_accum || img_ws, "gate_bwd: the bias gradient needs the zeroed [B, hidden] workspace");
    DECO_CHECK_ARG(M > 0 && hidden > 0 && hidden % 4 == 0 && hidden <= 2048 && rows_per_image > 0 &&
                   M % rows_per_image == 0 && gate_stride % 4 == 0, "gate_bwd: bad shape M=%lld hidden=%d L=%d", M, hidden, rows_per_image);
    const int rb = rows_block(rows_per_image, M);
    gate_bwd_kernel<<<(unsigned)(M / rb), ((hidden / 4) + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(
        ds, (const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)gate_bf16, gate_stride, (__nv_bfloat16*)da_bf16,
        dgate_accum, dgate_stride, dbias_accum ? img_ws : nullptr, rows_per_image, rb, hidden);
    DECO_CHECK_LAUNCH("gate_bwd_kernel");
    if (dbias_accum) {
        gate_bias_finalize_kernel<<<(hidden + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)gate_bf16, gate_stride, img_ws, dbias_accum, (int)(M / rows_per_image), hidden);
        DECO_CHECK_LAUNCH("gate_bias_finalize_kernel");
    }
    return DECO_OK;
}

extern "C" int deco_silu_add_rows_bwd(const void* dout_bf16, const float* x, const void* row_bf16, float* dx,
                                      float* drow_accum, int rows_per_image, long long M, int hidden, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(dout_bf16 && x && row_bf16 && dx && drow_accum, "silu_add_rows_bwd: null pointer");
    DECO_CHECK_ARG(M > 0 && hidden > 0 && hidden % 2 == 0 && hidden <= 512 * kMaxPairIters && rows_per_image > 0 &&
                   M % rows_per_image == 0, "silu_add_rows_bwd: bad shape");
    const int rb = rows_block(rows_per_image, M);
    silu_add_rows_bwd_kernel<<<(unsigned)(M / rb), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dout_bf16, x, (const __nv_bfloat16*)row_bf16, dx, drow_accum, rows_per_image, rb, hidden);
    DECO_CHECK_LAUNCH("silu_add_rows_bwd_kernel");
    return DECO_OK;
}

extern "C" int deco_swiglu_fwd(const void* y13_bf16, void* u_bf16, long long M, int ffn_pad, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(y13_bf16 && u_bf16 && M > 0 && ffn_pad > 0 && ffn_pad % 16 == 0, "swiglu_fwd: bad arguments");
    swiglu_kernel<false><<<bgrid(M * (ffn_pad / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)y13_bf16, nullptr, (__nv_bfloat16*)u_bf16, M, ffn_pad);
    DECO_CHECK_LAUNCH("swiglu_kernel<fwd>");
    return DECO_OK;
}

extern "C" int deco_swiglu_bwd(const void* y13_bf16, const void* du_bf16, void* dy13_bf16, long long M, int ffn_pad,
                               void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(y13_bf16 && du_bf16 && dy13_bf16 && M > 0 && ffn_pad > 0 && ffn_pad % 16 == 0, "swiglu_bwd: bad arguments");
    swiglu_kernel<true><<<bgrid(M * (ffn_pad / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)y13_bf16, (const __nv_bfloat16*)du_bf16, (__nv_bfloat16*)dy13_bf16, M, ffn_pad);
    DECO_CHECK_LAUNCH("swiglu_kernel<bwd>");
    return DECO_OK;
}

static int rmsnorm_modulate_bwd_impl(const void* dh_bf16, const float* x, const float* weight, const void* scale_bf16,
                                     long long mod_row_stride, float* ds_accum, float* dweight_accum,
                                     float* dshift_accum, float* dscale_accum, long long dmod_row_stride,
                                     float* row_ws, float* img_ws, int rows_per_image, long long M, int hidden, float eps,
                                     const deco::GateBwdArgs* gate, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(dh_bf16 && x && weight && scale_bf16 && ds_accum && dweight_accum && dshift_accum && dscale_accum && row_ws &&
                   img_ws, "rmsnorm_modulate_bwd: null pointer");
    DECO_CHECK_ARG(M > 0 && hidden > 0 && hidden % 4 == 0 && hidden <= 2048 && rows_per_image > 0 &&
                   M % rows_per_image == 0 && mod_row_stride % 4 == 0 && ((uintptr_t)row_ws & 7) == 0,
                   "rmsnorm_modulate_bwd: bad shape");
    const __nv_bfloat16* dhp = (const __nv_bfloat16*)dh_bf16;
    const __nv_bfloat16* scp = (const __nv_bfloat16*)scale_bf16;
    float2* stats = reinterpret_cast<float2*>(row_ws);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g1 = (unsigned)((M + 7) / 8);
    if (hidden <= 1024)
        rmsnorm_bwd_rowstats_kernel<8><<<g1, 256, 0, st>>>(dhp, x, weight, scp, mod_row_stride, stats, rows_per_image, M, hidden, eps);
    else
        rmsnorm_bwd_rowstats_kernel<16><<<g1, 256, 0, st>>>(dhp, x, weight, scp, mod_row_stride, stats, rows_per_image, M, hidden, eps);
    DECO_CHECK_LAUNCH("rmsnorm_bwd_rowstats_kernel");
    // one resident wave of blocks, each with an equal share of contiguous rows
    const int threads = ((hidden / 4) + 31) / 32 * 32;
    static int occ_cache[2][17] = {};           // resident blocks per SM by (gate variant, warps per block); 0 = not asked yet
    int& per_sm = occ_cache[gate ? 1 : 0][threads / 32];
    if (per_sm == 0) {
        int v = 0;
        cudaError_t oe = gate ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, rmsnorm_modulate_bwd_kernel<true>, threads, 0)
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, rmsnorm_modulate_bwd_kernel<false>, threads, 0);
        per_sm = (oe != cudaSuccess || v < 1) ? 1 : v;
    }
    long long nblk = (long long)per_sm * kNumSMs;
    if (nblk > M) nblk = M;
    if (gate)
        rmsnorm_modulate_bwd_kernel<true><<<(unsigned)nblk, threads, 0, st>>>(
            dhp, x, weight, scp, mod_row_stride, stats, ds_accum, img_ws, dshift_accum, dmod_row_stride, rows_per_image, M, hidden, *gate);
    else
        rmsnorm_modulate_bwd_kernel<false><<<(unsigned)nblk, threads, 0, st>>>(
            dhp, x, weight, scp, mod_row_stride, stats, ds_accum, img_ws, dshift_accum, dmod_row_stride, rows_per_image, M, hidden,
            GateBwdArgs{});
    DECO_CHECK_LAUNCH("rmsnorm_modulate_bwd_kernel");
    rmsnorm_bwd_finalize_kernel<<<(hidden + 31) / 32, dim3(32, 32), 0, st>>>(img_ws, weight, scp, mod_row_stride, dscale_accum,
                                                                       dmod_row_stride, dweight_accum,
                                                                       (int)(M / rows_per_image), hidden);
    DECO_CHECK_LAUNCH("rmsnorm_bwd_finalize_kernel");
    return DECO_OK;
}

extern "C" int deco_rmsnorm_modulate_bwd(const void* dh_bf16, const float* x, const float* weight, const void* scale_bf16,
                                         long long mod_row_stride, float* ds_accum, float* dweight_accum,
                                         float* dshift_accum, float* dscale_accum, long long dmod_row_stride,
                                         float* row_ws, float* img_ws, int rows_per_image, long long M, int hidden, float eps,
                                         void* stream)
{
    return rmsnorm_modulate_bwd_impl(dh_bf16, x, weight, scale_bf16, mod_row_stride, ds_accum, dweight_accum, dshift_accum,
                                     dscale_accum, dmod_row_stride, row_ws, img_ws, rows_per_image, M, hidden, eps, nullptr, stream);
}

// deco_rmsnorm_modulate_bwd followed by deco_gate_bwd on the UPDATED stream gradient, in one pass over it: the backward of
// "s_mid = s + gate * a; h = norm(s_mid)" seen from the norm's side (training backward; dit_c2i_DeCo.py:236-244 reversed).
// gate_img_ws [B, hidden] zeroed fp32 is needed only with dbias_accum.
extern "C" int deco_rmsnorm_modulate_bwd_gate(const void* dh_bf16, const float* x, const float* weight, const void* scale_bf16,
                                              long long mod_row_stride, float* ds_accum, float* dweight_accum,
                                              float* dshift_accum, float* dscale_accum, long long dmod_row_stride,
                                              float* row_ws, float* img_ws, int rows_per_image, long long M, int hidden, float eps,
                                              const void* a_bf16, const void* gate_bf16, long long gate_stride, void* da_bf16,
                                              float* dgate_accum, long long dgate_stride, float* dbias_accum, float* gate_img_ws,
                                              void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(a_bf16 && gate_bf16 && da_bf16 && dgate_accum, "rmsnorm_modulate_bwd_gate: null pointer");
    DECO_CHECK_ARG(!dbias_accum || gate_img_ws, "rmsnorm_modulate_bwd_gate: the bias gradient needs the zeroed [B, hidden] workspace");
    DECO_CHECK_ARG(gate_stride % 4 == 0, "rmsnorm_modulate_bwd_gate: bad gate stride");
    GateBwdArgs G;
    G.a = (const __nv_bfloat16*)a_bf16; G.gate = (const __nv_bfloat16*)gate_bf16; G.gate_stride = gate_stride;
    G.da = (__nv_bfloat16*)da_bf16; G.dgate = dgate_accum; G.dgate_stride = dgate_stride; G.usum = dbias_accum ? gate_img_ws : nullptr;
    int rc = rmsnorm_modulate_bwd_impl(dh_bf16, x, weight, scale_bf16, mod_row_stride, ds_accum, dweight_accum, dshift_accum,
                                       dscale_accum, dmod_row_stride, row_ws, img_ws, rows_per_image, M, hidden, eps, &G, stream);
    if (rc) return rc;
    if (dbias_accum) {
        gate_bias_finalize_kernel<<<(hidden + 31) / 32, dim3(32, 32), 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)gate_bf16, gate_stride, gate_img_ws, dbias_accum, (int)(M / rows_per_image), hidden);
        DECO_CHECK_LAUNCH("gate_bias_finalize_kernel");
    }
    return DECO_OK;
}

extern "C" int deco_headnorm_rope_bwd(void* g_bf16, long long g_stride, const void* raw_bf16, long long raw_stride,
                                      int col, const float* weight, const float* rope_cos_sin, float* dweight_accum,
                                      long long M, int heads, int head_dim, int L, float eps, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(g_bf16 && raw_bf16 && weight && dweight_accum && M > 0 && heads > 0 && L > 0, "headnorm_rope_bwd: bad arguments");
    DECO_CHECK_ARG(g_stride % 8 == 0 && raw_stride % 8 == 0 && col % 8 == 0, "headnorm_rope_bwd: strides / col must be multiples of 8");
    const long long items = M * heads;
    const int gpb = 8 * (32 / (head_dim == 72 ? 3 : 4));   // vector groups per 256-thread block
    long long nblk = (items + gpb - 1) / gpb;
    if (nblk > 3LL * kNumSMs) nblk = 3LL * kNumSMs;        // 3 resident blocks per SM; groups loop over their vectors
    const unsigned grid = (unsigned)nblk;
    const float2* rp = (const float2*)rope_cos_sin;
    DECO_CHECK_ARG(!rp || ((uintptr_t)rp & 15) == 0, "headnorm_rope_bwd: rope table must be 16-byte aligned");
    if (head_dim == 72)
        headnorm_rope_bwd_kernel<72><<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)g_bf16, g_stride, (const __nv_bfloat16*)raw_bf16, raw_stride, col, weight, rp, dweight_accum, M, heads, L, eps);
    else if (head_dim == 64)
        headnorm_rope_bwd_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)g_bf16, g_stride, (const __nv_bfloat16*)raw_bf16, raw_stride, col, weight, rp, dweight_accum, M, heads, L, eps);
    else {
        deco_set_error("headnorm_rope_bwd: head_dim %d not built (64, 72)", head_dim);
        return DECO_ERR_UNSUPPORTED;
    }
    DECO_CHECK_LAUNCH("headnorm_rope_bwd_kernel");
    return DECO_OK;
}

extern "C" int deco_cond_combine_bwd(const float* dc, const void* temb_bf16, const float* table, const long long* labels,
                                     float* dtemb_accum, float* dtable_accum, int B, int hidden, int num_rows, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(dc && temb_bf16 && table && labels && dtemb_accum && dtable_accum && B > 0 && hidden > 0 && num_rows > 0,
                   "cond_combine_bwd: bad arguments");
    const int n = B * hidden;
    cond_combine_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dc, (const __nv_bfloat16*)temb_bf16, table, labels,
                                                                               dtemb_accum, dtable_accum, B, hidden, num_rows);
    DECO_CHECK_LAUNCH("cond_combine_bwd_kernel");
    return DECO_OK;
}

extern "C" int deco_silu_bwd(const void* z_bf16, const void* dy_bf16, void* dz_bf16, long long n, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(z_bf16 && dy_bf16 && dz_bf16 && n > 0, "silu_bwd: bad arguments");
    silu_bwd_kernel<<<bgrid(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z_bf16, (const __nv_bfloat16*)dy_bf16,
                                                                     (__nv_bfloat16*)dz_bf16, n);
    DECO_CHECK_LAUNCH("silu_bwd_kernel");
    return DECO_OK;
}
