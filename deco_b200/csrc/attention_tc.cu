// Non-causal multi-head attention on the 5th-generation tensor cores, with q_norm / k_norm (RMSNorm over the head
// dimension) and the 2-D axial RoPE fused into the operand load.
//
// Replaces (reference, /root/reference): src/models/transformer/dit_c2i_DeCo.py:176-187 (split -> q_norm, k_norm ->
// apply_rotary_emb -> transposes -> scaled_dot_product_attention) and src/models/layers/attention_op.py:4; the second
// key/value segment covers the t2i joint attention (src/models/transformer/dit_t2i_pixnerd.py:46-59: keys =
// [image || text], k_norm on both, RoPE on the image part only).
//
// One CTA = one (batch row, head, 128-query tile); two CTAs are resident per SM (256 TMEM columns and ~101 KB of
// shared memory each), so one CTA's softmax overlaps the other's MMAs and loads.
//   4 warps (thread t <-> query row t <-> TMEM lane t):
//       gmem -> registers -> RMSNorm + RoPE -> bf16 -> shared memory in the 32-byte-swizzled, chunk-major operand
//       layout (sw32_offset) for Q and every 128-key block of K; V is copied raw with cp.async into the same layout;
//       S row from TMEM (tcgen05.ld), exact block max, exp2, row sum, P (bf16) written back over S (tcgen05.st).
//   thread 0, after each arrival round: S = Q K^T (tcgen05.mma SS, K-major x K-major) and O += P V (tcgen05.mma TS:
//       A = P in tensor memory, B = V MN-major), completion via tcgen05.commit -> mbarrier.  (A separate MMA warp would
//       cap the kernel at 168 registers per thread for two CTAs per SM; the S row alone is 128.)
// Online softmax without a per-block rescale of O: the exponent reference m only moves when a block's maximum exceeds
// it by more than 2^8 (then O and the row sum are rescaled through TMEM); otherwise P simply carries values up to 256.
// bf16/fp32 have the exponent range for that and the final 1/l normalisation uses the same reference, so the result is
// the exact softmax.  Operand layouts were validated stand-alone by scripts/umma_probe.cu.
#include "tcgen05.cuh"

namespace deco {

template <int D> struct TcAttnCfg {
    static constexpr int DP = (D + 15) / 16 * 16;   // head dim padded to the MMA K granularity (72 -> 80)
    static constexpr int NCH = DP / 16;             // 32-byte chunks per operand row
    static constexpr int C8 = D / 8;                // 16-byte chunks of real data per row
    static constexpr uint32_t TILE = NCH * 128 * 32;
    static constexpr uint32_t oQ = 0, oK = TILE, oV = 3 * TILE, oW = 5 * TILE;
    static constexpr uint32_t oBar = oW + 2 * D * 4;
    static constexpr uint32_t oSlot = oBar + 32;
    static constexpr uint32_t SMEM = oSlot + 16 + 256;   // + alignment slack
};

constexpr int kTcRows = 128;        // queries per CTA = keys per block = TMEM lanes
constexpr int kTcThreads = 128;     // thread t <-> query row t; thread 0 also issues the MMAs
constexpr uint32_t kColS = 0, kColO = 128, kTcTmemCols = 256;
constexpr float kRescaleThreshold = 8.0f;   // log2 units

struct TcAttnParams {
    const __nv_bfloat16* q; long long q_stride;          // row strides in elements
    const __nv_bfloat16* k[2]; const __nv_bfloat16* v[2]; long long kv_stride[2]; int Lk[2];
    __nv_bfloat16* o; long long o_stride;
    const float* qw; const float* kw;                     // RMSNorm weights [D] (both or neither)
    const float2* rope;                                   // (cos, sin) [Lq, D/2] for q and key segment 0, or null
    int Lq, heads;
    float scale_log2, eps;
};

__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D>
__device__ __forceinline__ void load_row(uint4 (&raw)[D / 8], const __nv_bfloat16* src, bool ok) {
#pragma unroll
    for (int c = 0; c < D / 8; ++c) raw[c] = ok ? ldg_nc16(src + c * 8) : make_uint4(0u, 0u, 0u, 0u);
}

// One operand row: optional RMSNorm (weights in shared memory) and RoPE, bf16, swizzled store.  Numerically identical
// to the stand-alone qknorm_rope_kernel (normalised value rounded to bf16 before the fp32 weight, dit_c2i_DeCo.py:99).
template <int D>
__device__ __forceinline__ void prep_row_store(const uint4 (&raw)[D / 8], uint8_t* tile, int row, const float* w,
                                               const float2* rope_row, float eps)
{
    using C = TcAttnCfg<D>;
    if (w == nullptr) {
#pragma unroll
        for (int c = 0; c < C::C8; ++c) *reinterpret_cast<uint4*>(tile + sw32_offset(row, c * 8, kTcRows)) = raw[c];
    } else {
        float v[D];
#pragma unroll
        for (int c = 0; c < C::C8; ++c) {
            const float2 a = unpack_bf2(raw[c].x), b = unpack_bf2(raw[c].y), cc = unpack_bf2(raw[c].z), d = unpack_bf2(raw[c].w);
            v[c * 8 + 0] = a.x; v[c * 8 + 1] = a.y; v[c * 8 + 2] = b.x; v[c * 8 + 3] = b.y;
            v[c * 8 + 4] = cc.x; v[c * 8 + 5] = cc.y; v[c * 8 + 6] = d.x; v[c * 8 + 7] = d.y;
        }
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < D; ++e) ss = fmaf(v[e], v[e], ss);
        const float rs = rsqrtf(ss / (float)D + eps);
#pragma unroll
        for (int c = 0; c < C::C8; ++c) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = c * 4 + e;   // pair index
                const float a = w[2 * j] * round_bf(v[2 * j] * rs);
                const float b = w[2 * j + 1] * round_bf(v[2 * j + 1] * rs);
                if (rope_row != nullptr) {
                    const float2 cs = __ldg(rope_row + j);
                    o[e] = pack_bf2(a * cs.x - b * cs.y, a * cs.y + b * cs.x);
                } else {
                    o[e] = pack_bf2(a, b);
                }
            }
            *reinterpret_cast<uint4*>(tile + sw32_offset(row, c * 8, kTcRows)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    if constexpr (C::DP > D) {   // contraction padding must be zero in both Q and K
#pragma unroll
        for (int c = C::C8; c < C::DP / 8; ++c)
            *reinterpret_cast<uint4*>(tile + sw32_offset(row, c * 8, kTcRows)) = make_uint4(0u, 0u, 0u, 0u);
    }
}

template <int D>
__global__ void __launch_bounds__(kTcThreads, 2) attention_tc_kernel(TcAttnParams P)
{
    using C = TcAttnCfg<D>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 255u) & ~255u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    float* sw = reinterpret_cast<float*>(gen + C::oW);
    const uint32_t bar_s = base + C::oBar, bar_o = bar_s + 8, bar_p = bar_s + 16, slot = base + C::oSlot;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int q0 = blockIdx.x * kTcRows;
    const int head = blockIdx.y;
    const long long b = blockIdx.z;
    const bool do_norm = P.qw != nullptr;

    const int nt0 = (P.Lk[0] + kTcRows - 1) / kTcRows;
    const int nt1 = (P.Lk[1] + kTcRows - 1) / kTcRows;
    const int nblk = nt0 + nt1;

    if (tid == 0) {
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        mbar_init(bar_p, 128);
        fence_barrier_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(slot, kTcTmemCols);
    if (do_norm)
        for (int i = tid; i < D; i += kTcThreads) { sw[i] = P.qw[i]; sw[D + i] = P.kw[i]; }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + C::oSlot);

    // ---------------------------------------------------------------------- MMA issue (thread 0 only)
    // phase 0: Q, K0, V0 staged -> S_0.  phase j + 1: P_j, K_{j+1}, V_{j+1} ready -> PV_j, then (once PV_j has
    // finished reading P_j, which aliases S) S_{j+1}.
    auto issue_phase = [&](int ph) {
        const uint32_t idesc_s = make_idesc_major(128, 128, 0, 0);
        const uint32_t idesc_o = make_idesc_major(128, C::DP, 0, 1);
        mbar_wait(bar_p, ph & 1);
        tc_fence_after();
        if (ph > 0) {
            const int j = ph - 1;
            const uint32_t sV = base + C::oV + (uint32_t)(j & 1) * C::TILE;
#pragma unroll
            for (int ks = 0; ks < kTcRows / 16; ++ks)
                umma_bf16_ts(tmem + kColO, tmem + kColS + (uint32_t)(ks * 8),
                             make_umma_desc(sV + ks * 512, kTcRows * 32, 256, 6), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
            umma_commit(bar_o);
        }
        if (ph < nblk) {
            if (ph > 0) {
                mbar_wait(bar_o, (ph - 1) & 1);
                tc_fence_after();
            }
            const uint32_t sK = base + C::oK + (uint32_t)(ph & 1) * C::TILE;
#pragma unroll
            for (int kc = 0; kc < C::NCH; ++kc)
                umma_bf16(tmem + kColS, make_umma_desc(base + C::oQ + kc * kTcRows * 32, 16, 256, 6),
                          make_umma_desc(sK + kc * kTcRows * 32, 16, 256, 6), idesc_s, kc ? 1u : 0u);
            umma_commit(bar_s);
        }
    };

    {
        // ------------------------------------------------------------------ load + softmax warps
        const int t = tid;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const float c = P.scale_log2;

        auto k_src = [&](int blk, bool& ok, const float2*& rp) -> const __nv_bfloat16* {
            const int seg = blk < nt0 ? 0 : 1;
            const int kk = (seg ? blk - nt0 : blk) * kTcRows + t;
            ok = kk < P.Lk[seg];
            rp = (seg == 0 && P.rope != nullptr && ok) ? P.rope + (long long)kk * (D / 2) : nullptr;
            return P.k[seg] + (b * P.Lk[seg] + (ok ? kk : 0)) * P.kv_stride[seg] + (long long)head * D;
        };
        auto v_copy = [&](int blk) {
            const int seg = blk < nt0 ? 0 : 1;
            const int kk = (seg ? blk - nt0 : blk) * kTcRows + t;
            const bool ok = kk < P.Lk[seg];
            const __nv_bfloat16* src = P.v[seg] + (b * P.Lk[seg] + (ok ? kk : 0)) * P.kv_stride[seg] + (long long)head * D;
            uint8_t* tile = gen + C::oV + (uint32_t)(blk & 1) * C::TILE;
#pragma unroll
            for (int ch = 0; ch < C::C8; ++ch) cp_async16(tile + sw32_offset(t, ch * 8, kTcRows), src + ch * 8, ok);
            cp_async_commit();
        };

        uint4 raw[C::C8];
        bool k_ok; const float2* k_rp;
        {
            const int qi = q0 + t;
            const bool ok = qi < P.Lq;
            load_row<D>(raw, P.q + (b * P.Lq + (ok ? qi : 0)) * P.q_stride + (long long)head * D, ok);
            uint4 rawk[C::C8];
            const __nv_bfloat16* ks = k_src(0, k_ok, k_rp);
            load_row<D>(rawk, ks, k_ok);
            v_copy(0);
            prep_row_store<D>(raw, gen + C::oQ, t, do_norm ? sw : nullptr,
                              (P.rope != nullptr && ok) ? P.rope + (long long)qi * (D / 2) : nullptr, P.eps);
            prep_row_store<D>(rawk, gen + C::oK, t, do_norm ? sw + D : nullptr, k_rp, P.eps);
        }
        if (nblk > 1) load_row<D>(raw, k_src(1, k_ok, k_rp), k_ok);
        cp_async_wait<0>();
        fence_proxy_async();
        mbar_arrive(bar_p);
        if (tid == 0) issue_phase(0);

        float m_used = 0.f, l = 0.f;
        for (int j = 0; j < nblk; ++j) {
            if (j + 1 < nblk) {       // K_{j+1}: its buffer was last read by S_{j-1}, complete since the last bar_s wait
                prep_row_store<D>(raw, gen + C::oK + (uint32_t)((j + 1) & 1) * C::TILE, t, do_norm ? sw + D : nullptr,
                                  k_rp, P.eps);
                if (j + 2 < nblk) load_row<D>(raw, k_src(j + 2, k_ok, k_rp), k_ok);
            }
            mbar_wait(bar_s, j & 1);      // S_j in TMEM; also implies PV_{j-1} complete (V buffer and O quiescent)
            tc_fence_after();
            if (j + 1 < nblk) v_copy(j + 1);

            uint32_t s[4][32];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) tmem_ld32(tmem + lane_base + kColS + (uint32_t)(qd * 32), s[qd]);
            tmem_ld_wait();

            {   // ragged tail of the segment
                const int seg = j < nt0 ? 0 : 1;
                const int valid = P.Lk[seg] - (seg ? j - nt0 : j) * kTcRows;
                if (valid < kTcRows) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (qd * 32 + i >= valid) s[qd][i] = 0xff800000u;   // -inf
                }
            }
            float mx[4];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                mx[qd] = __uint_as_float(s[qd][0]);
#pragma unroll
                for (int i = 1; i < 32; ++i) mx[qd] = fmaxf(mx[qd], __uint_as_float(s[qd][i]));
            }
            const float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
            if (j == 0) {
                m_used = bm;
            } else {
                const bool need = (bm - m_used) * c > kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    const float m_new = need ? bm : m_used;
                    const float corr = ex2_approx((m_used - m_new) * c);
                    l *= corr;
                    m_used = m_new;
#pragma unroll
                    for (int c0 = 0; c0 < C::DP; c0 += 16) {
                        uint32_t o[16];
                        tmem_ld16(tmem + lane_base + kColO + (uint32_t)c0, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                        tmem_st16(tmem + lane_base + kColO + (uint32_t)c0, o);
                    }
                }
            }
            const float mc = m_used * c;
            float sum[2] = {0.f, 0.f};
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(s[qd][2 * i]), c, -mc));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(s[qd][2 * i + 1]), c, -mc));
                    sum[0] += p0; sum[1] += p1;
                    pk[i] = pack_bf2(p0, p1);
                }
                tmem_st16(tmem + lane_base + kColS + (uint32_t)(qd * 16), pk);   // P over S: column = two keys
            }
            l += sum[0] + sum[1];
            tmem_st_wait();
            cp_async_wait<0>();
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_p);
            if (tid == 0) issue_phase(j + 1);
        }

        // ------------------------------------------------------------------ O / l -> global
        mbar_wait(bar_o, (nblk - 1) & 1);
        tc_fence_after();
        const int qi = q0 + t;
        const float inv = 1.0f / l;
        __nv_bfloat16* dst = P.o + (b * P.Lq + (qi < P.Lq ? qi : 0)) * P.o_stride + (long long)head * D;
        uint32_t o[D];
#pragma unroll
        for (int c0 = 0; c0 + 32 <= D; c0 += 32) tmem_ld32(tmem + lane_base + kColO + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&o[c0]));
        if constexpr (D % 32 == 8) tmem_ld8(tmem + lane_base + kColO + (uint32_t)(D - 8), *reinterpret_cast<uint32_t(*)[8]>(&o[D - 8]));
        static_assert(D % 32 == 0 || D % 32 == 8, "epilogue covers head dims 64 and 72");
        tmem_ld_wait();
        if (qi < P.Lq) {
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 8) {
                uint4 w4;
                w4.x = pack_bf2(__uint_as_float(o[c0 + 0]) * inv, __uint_as_float(o[c0 + 1]) * inv);
                w4.y = pack_bf2(__uint_as_float(o[c0 + 2]) * inv, __uint_as_float(o[c0 + 3]) * inv);
                w4.z = pack_bf2(__uint_as_float(o[c0 + 4]) * inv, __uint_as_float(o[c0 + 5]) * inv);
                w4.w = pack_bf2(__uint_as_float(o[c0 + 6]) * inv, __uint_as_float(o[c0 + 7]) * inv);
                *reinterpret_cast<uint4*>(dst + c0) = w4;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTcTmemCols);
}

template <int D>
static int launch_attention_tc(const TcAttnParams& P, int B, cudaStream_t st) {
    using C = TcAttnCfg<D>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) { deco_set_error("attention attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    dim3 grid((P.Lq + kTcRows - 1) / kTcRows, P.heads, B);
    attention_tc_kernel<D><<<grid, kTcThreads, C::SMEM, st>>>(P);
    DECO_CHECK_LAUNCH("attention_tc_kernel");
    return DECO_OK;
}

}  // namespace deco

extern "C" int deco_attention_fwd(const void* q, long long q_stride,
                                  const void* k0, const void* v0, long long kv0_stride, int Lk0,
                                  const void* k1, const void* v1, long long kv1_stride, int Lk1,
                                  void* out, long long out_stride,
                                  const float* q_norm_w, const float* k_norm_w, const float* rope_cos_sin, float eps,
                                  int B, int heads, int Lq, int head_dim, float scale, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(q && k0 && v0 && out, "attention: null pointer");
    DECO_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk0 > 0 && Lk1 >= 0, "attention: bad shape");
    DECO_CHECK_ARG(Lk1 == 0 || (k1 && v1), "attention: second key/value segment is null");
    DECO_CHECK_ARG(q_stride % 8 == 0 && kv0_stride % 8 == 0 && out_stride % 8 == 0 && (Lk1 == 0 || kv1_stride % 8 == 0),
                   "attention: row strides must keep 16-byte alignment");
    DECO_CHECK_ARG(((uintptr_t)q | (uintptr_t)k0 | (uintptr_t)v0 | (uintptr_t)out | (uintptr_t)k1 | (uintptr_t)v1) % 16 == 0,
                   "attention: pointers must be 16-byte aligned");
    DECO_CHECK_ARG((q_norm_w == nullptr) == (k_norm_w == nullptr), "attention: pass both norm weights or neither");
    DECO_CHECK_ARG(rope_cos_sin == nullptr || Lk0 == Lq, "attention: RoPE needs Lk0 == Lq (self-attention segment)");
    DECO_CHECK_ARG(B <= 65535 && heads <= 65535, "attention: grid too large");
    TcAttnParams P;
    P.q = (const __nv_bfloat16*)q; P.q_stride = q_stride;
    P.k[0] = (const __nv_bfloat16*)k0; P.v[0] = (const __nv_bfloat16*)v0; P.kv_stride[0] = kv0_stride; P.Lk[0] = Lk0;
    P.k[1] = (const __nv_bfloat16*)(Lk1 ? k1 : k0); P.v[1] = (const __nv_bfloat16*)(Lk1 ? v1 : v0);
    P.kv_stride[1] = Lk1 ? kv1_stride : kv0_stride; P.Lk[1] = Lk1;
    P.o = (__nv_bfloat16*)out; P.o_stride = out_stride;
    P.qw = q_norm_w; P.kw = k_norm_w; P.rope = (const float2*)rope_cos_sin;
    P.Lq = Lq; P.heads = heads;
    P.scale_log2 = scale * 1.4426950408889634f; P.eps = eps;
    if (head_dim == 72) return launch_attention_tc<72>(P, B, (cudaStream_t)stream);
    if (head_dim == 64) return launch_attention_tc<64>(P, B, (cudaStream_t)stream);
    deco_set_error("attention: head_dim %d not built (64, 72)", head_dim);
    return DECO_ERR_UNSUPPORTED;
}
