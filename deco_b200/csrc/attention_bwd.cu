// Backward of non-causal softmax attention (training step; F.scaled_dot_product_attention at
// /root/reference/src/models/transformer/dit_c2i_DeCo.py:181-185 under autograd).
//
// Two kernels, no atomics, deterministic:
//   attn_bwd_dq_kernel : CTA = 64 queries of one (image, head).  Pass 1 over the key blocks rebuilds the softmax statistics
//                        (the forward kernel keeps none): lse2 = log2 sum exp2(scale' s); pass 2 recomputes P, forms
//                        dP = dO V^T, dS = P (dP - delta) scale and accumulates dQ = dS K.  Writes dQ, lse2, delta.
//   attn_bwd_dkv_kernel: CTA = 64 keys of one (image, head), loops over the query blocks with the transposed products:
//                        P^T = exp2(scale' K Q^T - lse2), dV += P^T dO, dK += dS^T Q.
// Both use legacy warp MMAs (mma.sync m16n8k16 bf16, fp32 accumulate) on 64 x DP bf16 tiles staged in shared memory
// (DP = head_dim padded to a multiple of 16, zero-filled); a warp owns 16 rows.  Attention is 4 % of the step's FLOPs and
// its backward is 2.5x the forward, so this kernel is sized for correctness first; the tcgen05 forward stays the fast path.
#include "common.cuh"
#include <math.h>

namespace deco {

template <int D> struct BwdCfg {
    static constexpr int DP = (D + 15) / 16 * 16;
    static constexpr int LD = DP + 8;          // bf16 elements per smem row: 16-byte aligned, odd multiple of 16 B
    static constexpr int KS = DP / 16;         // k-steps over the head dimension
    static constexpr int NT = DP / 8;          // n-tiles over the head dimension
};

struct AttnBwdParams {
    const __nv_bfloat16 *q, *k, *v, *o, *dout;
    __nv_bfloat16 *dq, *dk, *dv;
    float *lse2, *delta;                       // [B * heads, Lq]
    long long q_stride, kv_stride, o_stride, do_stride, dq_stride, dkv_stride;
    int B, heads, Lq, Lk;
    float scale, scale_log2;
    int have_lse;                              // lse2 was written by the forward (deco_attention_fwd_lse): skip pass 1
};

// 64 x DP tile of rows [row0, row0 + 64) of a [rows, stride] matrix (columns [0, D)), zero-filled outside
template <int D>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, long long stride, int row0, int nrows)
{
    using C = BwdCfg<D>;
    constexpr int CH = C::DP / 8;
    for (int i = threadIdx.x; i < 64 * CH; i += blockDim.x) {
        const int r = i / CH, c = i % CH;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row0 + r < nrows && c * 8 < D) v = *reinterpret_cast<const uint4*>(g + (long long)(row0 + r) * stride + c * 8);
        *reinterpret_cast<uint4*>(s + r * C::LD + c * 8) = v;
    }
}

// same, asynchronously (cp.async 16 B with zero-fill); completion via cp_async_wait + __syncthreads
template <int D>
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* s, const __nv_bfloat16* g, long long stride, int row0, int nrows)
{
    using C = BwdCfg<D>;
    constexpr int CH = C::DP / 8;
    for (int i = threadIdx.x; i < 64 * CH; i += blockDim.x) {
        const int r = i / CH, c = i % CH;
        const bool ok = row0 + r < nrows && c * 8 < D;
        cp_async16(s + r * C::LD + c * 8, ok ? g + (long long)(row0 + r) * stride + c * 8 : g, ok);
    }
}

// A fragments (16 rows x DP) of rows [r0, r0 + 16) of a smem tile
template <int D>
__device__ __forceinline__ void load_afrags(uint32_t (&a)[BwdCfg<D>::KS][4], const __nv_bfloat16* s, int r0, int lane)
{
    using C = BwdCfg<D>;
#pragma unroll
    for (int ks = 0; ks < C::KS; ++ks)
        ldmatrix_x4(a[ks], s + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * C::LD + ks * 16 + (lane >> 4) * 8);
}

// acc[8][4] (16 x 64) = A (16 x DP, fragments) . T^T where T is a 64 x DP smem tile ([n][k] storage)
template <int D>
__device__ __forceinline__ void mma_a_tT(float (&acc)[8][4], const uint32_t (&a)[BwdCfg<D>::KS][4], const __nv_bfloat16* t, int lane)
{
    using C = BwdCfg<D>;
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
#pragma unroll
        for (int ks = 0; ks < C::KS; ++ks) {
            uint32_t b[4];
            ldmatrix_x4(b, t + (np * 16 + (lane & 7) + (lane >> 4) * 8) * C::LD + ks * 16 + ((lane >> 3) & 1) * 8);
            const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
            mma_bf16_16816(acc[2 * np], a[ks], b0);
            mma_bf16_16816(acc[2 * np + 1], a[ks], b1);
        }
    }
}

// out[NT][4] (16 x DP) += P (16 x 64, accumulator layout -> A fragments) . T where T is a 64 x DP smem tile ([k][n] storage)
template <int D>
__device__ __forceinline__ void mma_p_t(float (&out)[BwdCfg<D>::NT][4], const float (&p)[8][4], const __nv_bfloat16* t, int lane)
{
    using C = BwdCfg<D>;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
        a[0] = pack_bf2(p[2 * ks][0], p[2 * ks][1]);
        a[1] = pack_bf2(p[2 * ks][2], p[2 * ks][3]);
        a[2] = pack_bf2(p[2 * ks + 1][0], p[2 * ks + 1][1]);
        a[3] = pack_bf2(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
        for (int np = 0; np < C::NT / 2; ++np) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, t + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * C::LD + np * 16 + (lane >> 4) * 8);
            const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
            mma_bf16_16816(out[2 * np], a, b0);
            mma_bf16_16816(out[2 * np + 1], a, b1);
        }
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// store a 16 x DP accumulator tile (rows r0 + g, r0 + g + 8) as bf16, columns < D, rows < nrows
template <int D>
__device__ __forceinline__ void store_acc(const float (&acc)[BwdCfg<D>::NT][4], __nv_bfloat16* g, long long stride, int r0,
                                          int nrows, int lane)
{
    const int gq = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < BwdCfg<D>::NT; ++j) {
        const int col = 8 * j + 2 * t;
        if (col < D) {
            if (r0 + gq < nrows) *reinterpret_cast<uint32_t*>(g + (long long)(r0 + gq) * stride + col) = pack_bf2(acc[j][0], acc[j][1]);
            if (r0 + gq + 8 < nrows) *reinterpret_cast<uint32_t*>(g + (long long)(r0 + gq + 8) * stride + col) = pack_bf2(acc[j][2], acc[j][3]);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnBwdParams P)
{
    using C = BwdCfg<D>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE = 64 * C::LD;
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* sdO = sQ + TILE;
    __nv_bfloat16* sKb = sdO + TILE;            // [2] double-buffered key tiles
    __nv_bfloat16* sVb = sKb + 2 * TILE;        // [2] double-buffered value tiles
    float* sDelta = reinterpret_cast<float*>(sVb + 2 * TILE);

    const int qb = blockIdx.x, bh = blockIdx.y;
    const int b = bh / P.heads, h = bh % P.heads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int q0 = qb * 64;
    const __nv_bfloat16* qg = P.q + (long long)b * P.Lq * P.q_stride + (long long)h * D;
    const __nv_bfloat16* kg = P.k + (long long)b * P.Lk * P.kv_stride + (long long)h * D;
    const __nv_bfloat16* vg = P.v + (long long)b * P.Lk * P.kv_stride + (long long)h * D;
    const __nv_bfloat16* og = P.o + (long long)b * P.Lq * P.o_stride + (long long)h * D;
    const __nv_bfloat16* dog = P.dout + (long long)b * P.Lq * P.do_stride + (long long)h * D;

    load_tile<D>(sQ, qg, P.q_stride, q0, P.Lq);
    load_tile<D>(sdO, dog, P.do_stride, q0, P.Lq);
    // delta[row] = sum_d dO . O : two threads per row
    {
        const int r = threadIdx.x >> 1, half = threadIdx.x & 1;
        float acc = 0.f;
        if (q0 + r < P.Lq) {
            const __nv_bfloat16* op = og + (long long)(q0 + r) * P.o_stride;
            const __nv_bfloat16* dp = dog + (long long)(q0 + r) * P.do_stride;
            for (int c = half * 8; c < D; c += 16) {
                const uint4 a = *reinterpret_cast<const uint4*>(op + c), d = *reinterpret_cast<const uint4*>(dp + c);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 x = unpack_bf2(aw[e]), y = unpack_bf2(dw[e]);
                    acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
                }
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (half == 0) {
            sDelta[r] = acc;
            if (q0 + r < P.Lq) P.delta[(long long)bh * P.Lq + q0 + r] = acc;
        }
    }
    __syncthreads();
    uint32_t qf[C::KS][4], dof[C::KS][4];
    load_afrags<D>(qf, sQ, warp * 16, lane);
    load_afrags<D>(dof, sdO, warp * 16, lane);
    const float del0 = sDelta[warp * 16 + g], del1 = sDelta[warp * 16 + g + 8];
    const int nkb = (P.Lk + 63) / 64;
    // K (and, in pass 2, V) tiles stream through a 2-deep cp.async pipeline: step st = pass * nkb + kb uses buffer st & 1
    const int npass1 = P.have_lse ? 0 : nkb;
    const int total = npass1 + nkb;
    int step = 0;
    auto issue = [&](int st) {
        const int kb = st < npass1 ? st : st - npass1, bsel = st & 1;
        load_tile_async<D>(sKb + bsel * TILE, kg, P.kv_stride, kb * 64, P.Lk);
        if (st >= npass1) load_tile_async<D>(sVb + bsel * TILE, vg, P.kv_stride, kb * 64, P.Lk);
        cp_async_commit();
    };
    auto acquire = [&]() {
        if (step + 1 < total) { issue(step + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        return step & 1;
    };
    auto release = [&]() { __syncthreads(); ++step; };
    issue(0);

    // ---- pass 1: softmax statistics of rows g and g + 8
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int kb = 0; kb < npass1; ++kb) {
        const __nv_bfloat16* sK = sKb + acquire() * TILE;
        float s[8][4];
        mma_a_tT<D>(s, qf, sK, lane);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = kb * 64 + 8 * j + 2 * t + (e & 1);
                s[j][e] = key < P.Lk ? s[j][e] * P.scale_log2 : -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = quad_max(mx0); mx1 = quad_max(mx1);
        const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a0 += exp2f(s[j][0] - n0) + exp2f(s[j][1] - n0);
            a1 += exp2f(s[j][2] - n1) + exp2f(s[j][3] - n1);
        }
        a0 = quad_sum(a0); a1 = quad_sum(a1);
        l0 = l0 * exp2f(m0 - n0) + a0; l1 = l1 * exp2f(m1 - n1) + a1;
        m0 = n0; m1 = n1;
        release();
    }
    float lse0 = m0 + log2f(l0), lse1 = m1 + log2f(l1);
    {
        const int r = q0 + warp * 16 + g;
        if (P.have_lse) {
            lse0 = r < P.Lq ? P.lse2[(long long)bh * P.Lq + r] : 0.f;
            lse1 = r + 8 < P.Lq ? P.lse2[(long long)bh * P.Lq + r + 8] : 0.f;
        } else if (t == 0) {
            if (r < P.Lq) P.lse2[(long long)bh * P.Lq + r] = lse0;
            if (r + 8 < P.Lq) P.lse2[(long long)bh * P.Lq + r + 8] = lse1;
        }
    }

    // ---- pass 2: dQ
    float dq[C::NT][4];
#pragma unroll
    for (int j = 0; j < C::NT; ++j) { dq[j][0] = 0.f; dq[j][1] = 0.f; dq[j][2] = 0.f; dq[j][3] = 0.f; }
    for (int kb = 0; kb < nkb; ++kb) {
        const int bsel = acquire();
        const __nv_bfloat16* sK = sKb + bsel * TILE;
        const __nv_bfloat16* sV = sVb + bsel * TILE;
        float s[8][4], dp[8][4];
        mma_a_tT<D>(s, qf, sK, lane);
        mma_a_tT<D>(dp, dof, sV, lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = kb * 64 + 8 * j + 2 * t + (e & 1);
                const float p = key < P.Lk ? exp2f(s[j][e] * P.scale_log2 - (e < 2 ? lse0 : lse1)) : 0.f;
                s[j][e] = p * (dp[j][e] - (e < 2 ? del0 : del1)) * P.scale;     // dS
            }
        }
        mma_p_t<D>(dq, s, sK, lane);
        release();
    }
    __nv_bfloat16* dqg = P.dq + (long long)b * P.Lq * P.dq_stride + (long long)h * D;
    store_acc<D>(dq, dqg, P.dq_stride, q0 + warp * 16, P.Lq, lane);
}

template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnBwdParams P)
{
    using C = BwdCfg<D>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TILE = 64 * C::LD;
    __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* sV = sK + TILE;
    __nv_bfloat16* sQb = sV + TILE;             // [2] double-buffered query tiles
    __nv_bfloat16* sdOb = sQb + 2 * TILE;       // [2] double-buffered dO tiles
    float* sLseb = reinterpret_cast<float*>(sdOb + 2 * TILE);   // [2][64]
    float* sDeltab = sLseb + 128;                                // [2][64]

    const int kb = blockIdx.x, bh = blockIdx.y;
    const int b = bh / P.heads, h = bh % P.heads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3;
    const int k0 = kb * 64;
    const __nv_bfloat16* qg = P.q + (long long)b * P.Lq * P.q_stride + (long long)h * D;
    const __nv_bfloat16* kg = P.k + (long long)b * P.Lk * P.kv_stride + (long long)h * D;
    const __nv_bfloat16* vg = P.v + (long long)b * P.Lk * P.kv_stride + (long long)h * D;
    const __nv_bfloat16* dog = P.dout + (long long)b * P.Lq * P.do_stride + (long long)h * D;

    load_tile<D>(sK, kg, P.kv_stride, k0, P.Lk);
    load_tile<D>(sV, vg, P.kv_stride, k0, P.Lk);
    __syncthreads();
    uint32_t kf[C::KS][4], vf[C::KS][4];
    load_afrags<D>(kf, sK, warp * 16, lane);
    load_afrags<D>(vf, sV, warp * 16, lane);

    float dk[C::NT][4], dv[C::NT][4];
#pragma unroll
    for (int j = 0; j < C::NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dk[j][e] = 0.f; dv[j][e] = 0.f; }

    const int nqb = (P.Lq + 63) / 64;
    auto issue = [&](int qb) {
        const int bsel = qb & 1;
        load_tile_async<D>(sQb + bsel * TILE, qg, P.q_stride, qb * 64, P.Lq);
        load_tile_async<D>(sdOb + bsel * TILE, dog, P.do_stride, qb * 64, P.Lq);
        cp_async_commit();
        if (threadIdx.x < 64) {
            const int r = qb * 64 + threadIdx.x;
            sLseb[bsel * 64 + threadIdx.x] = r < P.Lq ? P.lse2[(long long)bh * P.Lq + r] : INFINITY;
            sDeltab[bsel * 64 + threadIdx.x] = r < P.Lq ? P.delta[(long long)bh * P.Lq + r] : 0.f;
        }
    };
    issue(0);
    for (int qb = 0; qb < nqb; ++qb) {
        if (qb + 1 < nqb) { issue(qb + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        const __nv_bfloat16* sQ = sQb + (qb & 1) * TILE;
        const __nv_bfloat16* sdO = sdOb + (qb & 1) * TILE;
        const float* sLse = sLseb + (qb & 1) * 64;
        const float* sDelta = sDeltab + (qb & 1) * 64;
        float s[8][4], dp[8][4];
        mma_a_tT<D>(s, kf, sQ, lane);      // S^T  : keys x queries
        mma_a_tT<D>(dp, vf, sdO, lane);    // dP^T : keys x queries
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int qc = 8 * j + 2 * t + (e & 1);
                const float p = exp2f(s[j][e] * P.scale_log2 - sLse[qc]);
                dp[j][e] = p * (dp[j][e] - sDelta[qc]) * P.scale;   // dS^T
                s[j][e] = p;                                         // P^T
            }
        }
        mma_p_t<D>(dv, s, sdO, lane);
        mma_p_t<D>(dk, dp, sQ, lane);
        __syncthreads();     // everyone is done with buffer qb & 1 before the next-but-one tile lands in it
    }
    __nv_bfloat16* dkg = P.dk + (long long)b * P.Lk * P.dkv_stride + (long long)h * D;
    __nv_bfloat16* dvg = P.dv + (long long)b * P.Lk * P.dkv_stride + (long long)h * D;
    store_acc<D>(dk, dkg, P.dkv_stride, k0 + warp * 16, P.Lk, lane);
    store_acc<D>(dv, dvg, P.dkv_stride, k0 + warp * 16, P.Lk, lane);
}

template <int D>
static int launch_attn_bwd(const AttnBwdParams& P, cudaStream_t st)
{
    using C = BwdCfg<D>;
    const int smem = 6 * 64 * C::LD * 2 + 4 * 64 * 4;
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { deco_set_error("attention_bwd attr: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 gq((P.Lq + 63) / 64, P.B * P.heads), gk((P.Lk + 63) / 64, P.B * P.heads);
    attn_bwd_dq_kernel<D><<<gq, 128, smem, st>>>(P);
    DECO_CHECK_LAUNCH("attn_bwd_dq_kernel");
    attn_bwd_dkv_kernel<D><<<gk, 128, smem, st>>>(P);
    DECO_CHECK_LAUNCH("attn_bwd_dkv_kernel");
    return DECO_OK;
}

}  // namespace deco

extern "C" int deco_attention_bwd(const void* q, long long q_stride, const void* k, const void* v, long long kv_stride,
                                  const void* o, long long o_stride, const void* dout, long long do_stride,
                                  void* dq, long long dq_stride, void* dk, void* dv, long long dkv_stride,
                                  float* lse2_ws, float* delta_ws, int have_lse, int B, int heads, int Lq, int Lk, int head_dim,
                                  float scale, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(q && k && v && o && dout && dq && dk && dv && lse2_ws && delta_ws, "attention_bwd: null pointer");
    DECO_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk > 0 && B * heads <= 65535, "attention_bwd: bad shape");
    DECO_CHECK_ARG(q_stride % 8 == 0 && kv_stride % 8 == 0 && o_stride % 8 == 0 && do_stride % 8 == 0 &&
                   dq_stride % 2 == 0 && dkv_stride % 2 == 0, "attention_bwd: strides must be multiples of 8 elements");
    DECO_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o | (uintptr_t)dout) & 15) == 0 &&
                   (((uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 3) == 0, "attention_bwd: misaligned pointer");
    AttnBwdParams P;
    P.q = (const __nv_bfloat16*)q; P.k = (const __nv_bfloat16*)k; P.v = (const __nv_bfloat16*)v;
    P.o = (const __nv_bfloat16*)o; P.dout = (const __nv_bfloat16*)dout;
    P.dq = (__nv_bfloat16*)dq; P.dk = (__nv_bfloat16*)dk; P.dv = (__nv_bfloat16*)dv;
    P.lse2 = lse2_ws; P.delta = delta_ws;
    P.q_stride = q_stride; P.kv_stride = kv_stride; P.o_stride = o_stride; P.do_stride = do_stride;
    P.dq_stride = dq_stride; P.dkv_stride = dkv_stride;
    P.B = B; P.heads = heads; P.Lq = Lq; P.Lk = Lk;
    P.scale = scale; P.scale_log2 = scale * 1.4426950408889634f; P.have_lse = have_lse;
    if (head_dim == 72) return launch_attn_bwd<72>(P, (cudaStream_t)stream);
    if (head_dim == 64) return launch_attn_bwd<64>(P, (cudaStream_t)stream);
    deco_set_error("attention_bwd: head_dim %d not built (64, 72)", head_dim);
    return DECO_ERR_UNSUPPORTED;
}
