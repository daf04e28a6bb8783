// Non-causal multi-head attention, softmax(Q K^T / sqrt(d)) V, flash-style (online softmax, K/V streamed through
// shared memory with cp.async double buffering).  Q, K, V are read *in place* from the QKV GEMM output
// ([tokens, 3*H] rows, head h at column h*d) so no head transpose is ever materialised, and the output is written
// straight into the [tokens, H] layout the projection GEMM consumes.
//
// Replaces (reference, /root/reference): src/models/transformer/dit_c2i_DeCo.py:181-187 (SDPA + transposes) and
// src/models/layers/attention_op.py:4; two key/value segments cover the t2i joint attention
// (src/models/transformer/dit_t2i_pixnerd.py:52-59: keys = [image || text]).
//
// v1 uses warp-level mma.sync (m16n8k16, bf16 in / fp32 accumulate): 8 warps x 16 query rows per CTA, 64-key tiles.
// Head dims 64 and 72 (72 is zero-padded to 80 along the QK^T contraction in shared memory only).
// A tcgen05/TMEM version is the planned replacement (DESIGN.md "next"); attention is 3.5 % of the step FLOPs.
#include "common.cuh"

namespace deco {

template <int D> struct AttnCfg {
    static constexpr int DP = (D + 15) / 16 * 16;      // contraction length for QK^T
    static constexpr int DS = (D == 64) ? 72 : 88;     // smem row stride (elements): conflict-free ldmatrix
    static constexpr int KSTEPS = DP / 16;
    static constexpr int NT_O = D / 8;                 // output n-tiles
    static constexpr int CHUNKS = D / 8;               // 16-byte chunks per row
};

constexpr int kAttnBM = 128;   // query rows per CTA
constexpr int kAttnBN = 64;    // keys per tile
constexpr int kAttnThreads = 256;

struct AttnParams {
    const __nv_bfloat16* q; long long q_stride;        // row stride in elements
    const __nv_bfloat16* k[2]; const __nv_bfloat16* v[2]; long long kv_stride[2]; int Lk[2];
    __nv_bfloat16* o; long long o_stride;
    int Lq, heads;
    float scale_log2;   // scale * log2(e)
};

template <int D>
__device__ __forceinline__ void attn_load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, long long stride,
                                               int row0, int nrows_valid, int nrows_tile, int tid)
{
    using C = AttnCfg<D>;
    for (int i = tid; i < nrows_tile * C::CHUNKS; i += kAttnThreads) {
        const int r = i / C::CHUNKS, c = i % C::CHUNKS;
        const bool ok = (row0 + r) < nrows_valid;
        const __nv_bfloat16* g = src + (long long)(ok ? row0 + r : 0) * stride + c * 8;
        cp_async16(dst + r * C::DS + c * 8, g, ok);
    }
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 2) attention_fwd_kernel(AttnParams P)
{
    using C = AttnCfg<D>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* sK = sQ + kAttnBM * C::DS;               // [2][BN][DS]
    __nv_bfloat16* sV = sK + 2 * kAttnBN * C::DS;           // [2][BN][DS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * kAttnBM;
    const int head = blockIdx.y;
    const long long b = blockIdx.z;

    // zero the contraction padding (columns D..DP-1) of Q and both K buffers once; cp.async never touches it
    if constexpr (C::DP > D) {
        for (int i = tid; i < (kAttnBM + 2 * kAttnBN) * (C::DP - D); i += kAttnThreads) {
            const int r = i / (C::DP - D), c = i % (C::DP - D);
            __nv_bfloat16* base = (r < kAttnBM) ? (sQ + r * C::DS) : (sK + (r - kAttnBM) * C::DS);
            base[D + c] = f2bf(0.f);
        }
    }

    const __nv_bfloat16* qg = P.q + (b * P.Lq) * P.q_stride + (long long)head * D;
    attn_load_tile<D>(sQ, qg, P.q_stride, q0, P.Lq, kAttnBM, tid);

    const int nt0 = (P.Lk[0] + kAttnBN - 1) / kAttnBN;
    const int nt1 = (P.Lk[1] + kAttnBN - 1) / kAttnBN;
    const int ntiles = nt0 + nt1;

    auto issue_kv = [&](int tile, int buf) {
        const int seg = tile < nt0 ? 0 : 1;
        const int lt = seg ? tile - nt0 : tile;
        const __nv_bfloat16* kg = P.k[seg] + (b * P.Lk[seg]) * P.kv_stride[seg] + (long long)head * D;
        const __nv_bfloat16* vg = P.v[seg] + (b * P.Lk[seg]) * P.kv_stride[seg] + (long long)head * D;
        attn_load_tile<D>(sK + buf * kAttnBN * C::DS, kg, P.kv_stride[seg], lt * kAttnBN, P.Lk[seg], kAttnBN, tid);
        attn_load_tile<D>(sV + buf * kAttnBN * C::DS, vg, P.kv_stride[seg], lt * kAttnBN, P.Lk[seg], kAttnBN, tid);
    };

    issue_kv(0, 0);
    cp_async_commit();

    float o_acc[C::NT_O][4];
#pragma unroll
    for (int j = 0; j < C::NT_O; ++j) { o_acc[j][0] = o_acc[j][1] = o_acc[j][2] = o_acc[j][3] = 0.f; }
    float row_max[2] = {-INFINITY, -INFINITY}, row_sum[2] = {0.f, 0.f};
    uint32_t qf[C::KSTEPS][4];

    for (int tile = 0; tile < ntiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < ntiles) issue_kv(tile + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        if (tile == 0) {
            // Q fragments: rows warp*16 .. +15
#pragma unroll
            for (int s = 0; s < C::KSTEPS; ++s) {
                const __nv_bfloat16* p = sQ + (warp * 16 + (lane & 15)) * C::DS + s * 16 + (lane >> 4) * 8;
                ldmatrix_x4(qf[s], p);
            }
        }
        const __nv_bfloat16* kb = sK + buf * kAttnBN * C::DS;
        const __nv_bfloat16* vb = sV + buf * kAttnBN * C::DS;

        // ---- S = Q K^T  (16 x 64 per warp)
        float s_acc[kAttnBN / 8][4];
#pragma unroll
        for (int j = 0; j < kAttnBN / 8; ++j) {
            s_acc[j][0] = s_acc[j][1] = s_acc[j][2] = s_acc[j][3] = 0.f;
            const __nv_bfloat16* krow = kb + (j * 8 + (lane & 7)) * C::DS + (lane >> 3) * 8;
#pragma unroll
            for (int s2 = 0; s2 + 1 < C::KSTEPS; s2 += 2) {
                uint32_t r[4];
                ldmatrix_x4(r, krow + s2 * 16);
                const uint32_t b0[2] = {r[0], r[1]}, b1[2] = {r[2], r[3]};
                mma_bf16_16816(s_acc[j], qf[s2], b0);
                mma_bf16_16816(s_acc[j], qf[s2 + 1], b1);
            }
            if (C::KSTEPS & 1) {
                uint32_t r[2];
                ldmatrix_x2(r, kb + (j * 8 + (lane & 7)) * C::DS + (C::KSTEPS - 1) * 16 + ((lane >> 3) & 1) * 8);
                mma_bf16_16816(s_acc[j], qf[C::KSTEPS - 1], r);
            }
        }

        // ---- mask the ragged tail of the segment
        {
            const int seg = tile < nt0 ? 0 : 1;
            const int lt = seg ? tile - nt0 : tile;
            const int valid = P.Lk[seg] - lt * kAttnBN;
            if (valid < kAttnBN) {
#pragma unroll
                for (int j = 0; j < kAttnBN / 8; ++j) {
                    const int c0 = j * 8 + 2 * t;
                    if (c0 >= valid) { s_acc[j][0] = -INFINITY; s_acc[j][2] = -INFINITY; }
                    if (c0 + 1 >= valid) { s_acc[j][1] = -INFINITY; s_acc[j][3] = -INFINITY; }
                }
            }
        }

        // ---- online softmax (rows g and g+8)
        float mx[2] = {row_max[0], row_max[1]};
#pragma unroll
        for (int j = 0; j < kAttnBN / 8; ++j) {
            mx[0] = fmaxf(mx[0], fmaxf(s_acc[j][0], s_acc[j][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s_acc[j][2], s_acc[j][3]));
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], 1));
            mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], 2));
        }
        float corr[2], msc[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            corr[e] = exp2f((row_max[e] - mx[e]) * P.scale_log2);   // row_max = -inf on the first tile -> 0
            msc[e] = mx[e] * P.scale_log2;
            row_max[e] = mx[e];
            row_sum[e] *= corr[e];
        }
#pragma unroll
        for (int j = 0; j < C::NT_O; ++j) {
            o_acc[j][0] *= corr[0]; o_acc[j][1] *= corr[0];
            o_acc[j][2] *= corr[1]; o_acc[j][3] *= corr[1];
        }
        uint32_t pf[kAttnBN / 16][4];
#pragma unroll
        for (int j = 0; j < kAttnBN / 8; ++j) {
            const float p0 = exp2f(fmaf(s_acc[j][0], P.scale_log2, -msc[0]));
            const float p1 = exp2f(fmaf(s_acc[j][1], P.scale_log2, -msc[0]));
            const float p2 = exp2f(fmaf(s_acc[j][2], P.scale_log2, -msc[1]));
            const float p3 = exp2f(fmaf(s_acc[j][3], P.scale_log2, -msc[1]));
            row_sum[0] += p0 + p1;
            row_sum[1] += p2 + p3;
            // accumulator layout of S == A-operand layout of P for the PV product
            pf[j >> 1][(j & 1) * 2 + 0] = pack_bf2(p0, p1);
            pf[j >> 1][(j & 1) * 2 + 1] = pack_bf2(p2, p3);
        }

        // ---- O += P V
#pragma unroll
        for (int ks = 0; ks < kAttnBN / 16; ++ks) {
            const __nv_bfloat16* vrow = vb + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * C::DS + (lane >> 4) * 8;
#pragma unroll
            for (int j2 = 0; j2 + 1 < C::NT_O; j2 += 2) {
                uint32_t r[4];
                ldmatrix_x4_trans(r, vrow + j2 * 8);
                const uint32_t b0[2] = {r[0], r[1]}, b1[2] = {r[2], r[3]};
                mma_bf16_16816(o_acc[j2], pf[ks], b0);
                mma_bf16_16816(o_acc[j2 + 1], pf[ks], b1);
            }
            if (C::NT_O & 1) {
                uint32_t r[2];
                ldmatrix_x2_trans(r, vb + (ks * 16 + (lane & 15)) * C::DS + (C::NT_O - 1) * 8);
                mma_bf16_16816(o_acc[C::NT_O - 1], pf[ks], r);
            }
        }
        __syncthreads();   // all warps done with buf before the next prefetch overwrites it
    }

    // ---- normalise and store
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        row_sum[e] += __shfl_xor_sync(0xffffffffu, row_sum[e], 1);
        row_sum[e] += __shfl_xor_sync(0xffffffffu, row_sum[e], 2);
    }
    const float inv0 = 1.0f / row_sum[0], inv1 = 1.0f / row_sum[1];
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __nv_bfloat16* og = P.o + (b * P.Lq) * P.o_stride + (long long)head * D;
#pragma unroll
    for (int j = 0; j < C::NT_O; ++j) {
        const int col = j * 8 + 2 * t;
        if (r0 < P.Lq) *reinterpret_cast<uint32_t*>(og + (long long)r0 * P.o_stride + col) = pack_bf2(o_acc[j][0] * inv0, o_acc[j][1] * inv0);
        if (r1 < P.Lq) *reinterpret_cast<uint32_t*>(og + (long long)r1 * P.o_stride + col) = pack_bf2(o_acc[j][2] * inv1, o_acc[j][3] * inv1);
    }
}

template <int D>
static int launch_attention(const AttnParams& P, int B, cudaStream_t st) {
    using C = AttnCfg<D>;
    const int smem = (kAttnBM + 4 * kAttnBN) * C::DS * (int)sizeof(__nv_bfloat16);
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { deco_set_error("attention attr: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid((P.Lq + kAttnBM - 1) / kAttnBM, P.heads, B);
    attention_fwd_kernel<D><<<grid, kAttnThreads, smem, st>>>(P);
    DECO_CHECK_LAUNCH("attention_fwd_kernel");
    return DECO_OK;
}

}  // namespace deco

extern "C" int deco_attention_fwd(const void* q, long long q_stride,
                                  const void* k0, const void* v0, long long kv0_stride, int Lk0,
                                  const void* k1, const void* v1, long long kv1_stride, int Lk1,
                                  void* out, long long out_stride,
                                  int B, int heads, int Lq, int head_dim, float scale, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(q && k0 && v0 && out, "attention: null pointer");
    DECO_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk0 > 0 && Lk1 >= 0, "attention: bad shape");
    DECO_CHECK_ARG(q_stride % 8 == 0 && kv0_stride % 8 == 0 && out_stride % 2 == 0 && (Lk1 == 0 || kv1_stride % 8 == 0),
                   "attention: row strides must keep 16-byte alignment");
    DECO_CHECK_ARG(B <= 65535 && heads <= 65535, "attention: grid too large");
    AttnParams P;
    P.q = (const __nv_bfloat16*)q; P.q_stride = q_stride;
    P.k[0] = (const __nv_bfloat16*)k0; P.v[0] = (const __nv_bfloat16*)v0; P.kv_stride[0] = kv0_stride; P.Lk[0] = Lk0;
    P.k[1] = (const __nv_bfloat16*)(Lk1 ? k1 : k0); P.v[1] = (const __nv_bfloat16*)(Lk1 ? v1 : v0);
    P.kv_stride[1] = Lk1 ? kv1_stride : kv0_stride; P.Lk[1] = Lk1;
    P.o = (__nv_bfloat16*)out; P.o_stride = out_stride; P.Lq = Lq; P.heads = heads;
    P.scale_log2 = scale * 1.4426950408889634f;
    if (head_dim == 72) return launch_attention<72>(P, B, (cudaStream_t)stream);
    if (head_dim == 64) return launch_attention<64>(P, B, (cudaStream_t)stream);
    deco_set_error("attention: head_dim %d not built (64, 72)", head_dim);
    return DECO_ERR_UNSUPPORTED;
}
