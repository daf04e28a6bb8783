// Hyper-network pixel decoder of the PixNerd baseline (configs_c2i/Baseline_PixNerd.yaml) on tcgen05 / TMEM.
//
// Replaces (reference, /root/reference/src/models/transformer/dit_c2i_pixnerd.py):
//   :212-248  NerfEmbedder (constant positional table, folded on the host into T[pixel] + Wrgb rgb)
//   :250-273  NerfBlock.forward: per PATCH the MLP weights fc1 [64 x 128], fc2 [128 x 64] come out of a Linear on the patch's
//             DiT condition (param_generator1: the big GEMM, run by the caller with the tcgen05 GEMM), are L2-normalised over
//             their input dimension, and x <- x + silu(rmsnorm(x) fc1n) fc2n for the patch's 256 pixels
//   :275-283  NerfFinalLayer (RMSNorm + Linear 64 -> 3);  :376-380 reshape / transpose / F.fold
//
// Geometry as in csrc/decoder_tc.cu: a tile is 128 pixels (half a 16 x 16 patch), pixel <-> TMEM lane <-> epilogue thread;
// the fp32 residual x[64] lives in registers.  A CTA walks over tokens; its two SLOTS take the two halves of the token, so
// both use the same generated weights: the token's fc1 | fc2 (32 KB bf16 per block) are fetched by TMA straight out of the
// generator GEMM's output row into the MN-major operand layout (3-D tensor maps: n, k, token), double buffered.
//   MMA1  D[128 x 128] = A . fc1   (A = bf16(rmsnorm(x) w) in tensor memory, K = 64: 4 tcgen05.mma, B MN-major in smem)
//   E1    t = silu(D / ||fc1[:, n]||) -> bf16 -> A                       (column norms: see below)
//   MMA2  D[128 x 64]  = t . fc2   (K = 128: 8 tcgen05.mma)
//   E2    x += D / ||fc2[:, n]|| ; next block's rmsnorm -> A (or the final norm; MMA_f: N = 16, K = 64 against Wf)
// F.normalize(fc, dim=-2) divides every COLUMN of the generated matrix by its norm, which commutes with the product:
// (h fc / ||fc_n||)_n = (h fc)_n / ||fc_n||, so the raw bf16 weights feed the tensor core and the 128 + 64 reciprocal norms
// (computed from the staged tile by the 256 epilogue threads, one column each, while MMA1 runs) scale the fp32 accumulator.
// Bound: the generator GEMMs (2 x 16.8 MMAC per token) dominate the decoder; this kernel moves 64 KB of generated weights
// per token (HBM) for 2 x 4.2 MFLOP ... x 256 pixels = 8.4 MFLOP per token: memory-bound on the generated weights.
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace deco {
namespace nerf {

constexpr int kHx = 64, kHm = 128;
constexpr int kSlots = 2;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + kSlots + 2) * 32;     // + one issuer per slot, + TMA producer / TMEM allocator, + spare
constexpr int kMaxR = 4;
constexpr uint32_t kColD = 0, kColA = 128, kSlotCols = 192;
constexpr uint32_t kFc1Bytes = kHx * kHm * 2, kFc2Bytes = kHm * kHx * 2, kWBuf = kFc1Bytes + kFc2Bytes;
constexpr int kTabPitch = 68;
// constant blob: Wf tile [16 x 64] bf16 (K-major SW32, 2048 B) | fp32: norm weights [R][64] | final norm [64] | bias_f [4] |
// Wrgb [64][4] | T [256][68]
constexpr uint32_t kWfBytes = 2048;
__host__ __device__ inline uint32_t blob_floats(int R) { return (uint32_t)(R * kHx + kHx + 4 + kHx * 4 + 256 * kTabPitch); }
__host__ __device__ inline uint32_t blob_bytes(int R) { return kWfBytes + blob_floats(R) * 4; }
__host__ __device__ inline uint32_t smem_bytes(int R) {
    return ((blob_bytes(R) + 1023u) & ~1023u) + 2 * kWBuf + 2 * (kHm + kHx) * 4 + 256 /*barriers*/ + 1024 /*align*/;
}

struct Maps { CUtensorMap fc1[kMaxR], fc2[kMaxR]; };

struct Params {
    const float* x;
    const void* blob;
    void* out;
    int out_bf16;
    int R, H, W, Hp, Wp;
    int tokens;
};

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
nerf_decoder_kernel(const __grid_constant__ Maps maps, const Params P)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const int R = P.R;
    const uint32_t sWf = base;
    const float* sNormW = reinterpret_cast<const float*>(gen + kWfBytes);       // [R][64]
    const float* sNormF = sNormW + R * kHx;                                     // [64]
    const float* sBiasF = sNormF + kHx;                                         // [4]
    const float* sWrgb = sBiasF + 4;                                            // [64][4]
    const float* sTab = sWrgb + kHx * 4;                                        // [256][68]
    const uint32_t sWB = base + ((blob_bytes(R) + 1023u) & ~1023u);             // [2][fc1 | fc2]
    const uint32_t oInv = (sWB - base) + 2 * kWBuf;
    float* sInv = reinterpret_cast<float*>(gen + oInv);                         // [2][128 + 64] reciprocal column norms
    const uint32_t bars = base + oInv + 2 * (kHm + kHx) * 4;
    auto d_bar = [&](int k) { return bars + 8u * k; };
    auto a_bar = [&](int k) { return bars + 16u + 8u * k; };
    auto w_full = [&](int b) { return bars + 32u + 8u * b; };
    auto w_empty = [&](int b) { return bars + 48u + 8u * b; };
    const uint32_t tslot = bars + 64u;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int first = blockIdx.x, step = gridDim.x;
    const int nlocal = first < P.tokens ? (P.tokens - first + step - 1) / step : 0;     // tokens of this CTA

    if (tid == 0) {
        for (int k = 0; k < kSlots; ++k) { mbar_init(d_bar(k), 1); mbar_init(a_bar(k), 4); }
        for (int b = 0; b < 2; ++b) { mbar_init(w_full(b), 1); mbar_init(w_empty(b), kSlots); }
        fence_barrier_init();
        for (int j = 0; j < R; ++j) { tma_prefetch_desc(&maps.fc1[j]); tma_prefetch_desc(&maps.fc2[j]); }
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.blob);
        uint4* dst = reinterpret_cast<uint4*>(gen);
        const int n16 = (int)(blob_bytes(R) / 16);
        for (int i = tid; i < n16; i += kThreads) dst[i] = __ldg(src + i);
    }
    if (warp == kEpiWarps + kSlots) tmem_alloc(tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    pdl_launch_dependents();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tslot - base));

    if (warp == kEpiWarps + kSlots) {
        // ================================================================== TMA producer: generated weights of (token, block)
        int q = 0;
        for (int n = 0; n < nlocal; ++n) {
            const int m = first + n * step;
            for (int j = 0; j < R; ++j, ++q) {
                const int b = q & 1;
                if (q >= 2) mbar_wait(w_empty(b), (uint32_t)(((q >> 1) - 1) & 1));
                if (elect_one()) {
                    const uint32_t dst = sWB + (uint32_t)b * kWBuf;
                    mbar_expect_tx(w_full(b), kWBuf);
#pragma unroll
                    for (int c = 0; c < kHm / 16; ++c) tma_load_3d(dst + c * (kHx * 32), &maps.fc1[j], w_full(b), 16 * c, 0, m);
#pragma unroll
                    for (int c = 0; c < kHx / 16; ++c) tma_load_3d(dst + kFc1Bytes + c * (kHm * 32), &maps.fc2[j], w_full(b), 16 * c, 0, m);
                }
                __syncwarp();
            }
        }
    } else if (warp >= kEpiWarps && warp < kEpiWarps + kSlots) {
        // ================================================================== MMA issuer of slot k
        const int k = warp - kEpiWarps;
        constexpr uint32_t id1 = make_idesc_major(128, kHm, 0, 1), id2 = make_idesc_major(128, kHx, 0, 1),
                           idf = make_idesc_major(128, 16, 0, 0);
        const uint32_t tcol = tmem + (uint32_t)k * kSlotCols;
        const uint32_t acol = tcol + kColA;
        const uint64_t dWf = make_umma_desc(sWf, 16, 256, 6);
        uint32_t aphase = 0;
        int q = 0;
        for (int n = 0; n < nlocal; ++n) {
            for (int j = 0; j < R; ++j, ++q) {
                const int b = q & 1;
                const uint64_t d1 = make_umma_desc(sWB + (uint32_t)b * kWBuf, kHx * 32, 256, 6);               // fc1: K = 64 rows per chunk
                const uint64_t d2 = make_umma_desc(sWB + (uint32_t)b * kWBuf + kFc1Bytes, kHm * 32, 256, 6);   // fc2: K = 128 rows per chunk
                mbar_wait(a_bar(k), aphase); aphase ^= 1;                   // A = bf16(rmsnorm(x) w)
                mbar_wait(w_full(b), (uint32_t)((q >> 1) & 1));
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kHx / 16; ++ks)
                        umma_bf16_ts(tcol + kColD, acol + (uint32_t)(ks * 8), d1 + (uint64_t)(ks * (512 >> 4)), id1, ks ? 1u : 0u);
                    umma_commit(d_bar(k));
                }
                __syncwarp();
                mbar_wait(a_bar(k), aphase); aphase ^= 1;                   // A = bf16(silu(D / norm))
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kHm / 16; ++ks)
                        umma_bf16_ts(tcol + kColD, acol + (uint32_t)(ks * 8), d2 + (uint64_t)(ks * (512 >> 4)), id2, ks ? 1u : 0u);
                    umma_commit(d_bar(k));
                    umma_commit(w_empty(b));                                // this slot's last use of the weight buffer
                }
                __syncwarp();
            }
            mbar_wait(a_bar(k), aphase); aphase ^= 1;                       // A = bf16(rmsnorm(x) w_final)
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < kHx / 16; ++ks)
                    umma_bf16_ts(tcol + kColD, acol + (uint32_t)(ks * 8), dWf + (uint64_t)(ks * ((16 * 32) >> 4)), idf, ks ? 1u : 0u);
                umma_commit(d_bar(k));
            }
            __syncwarp();
        }
    } else if (warp < kEpiWarps) {
        // ================================================================== epilogue: thread = pixel = TMEM lane
        const int k = warp >> 2;
        const int l = (warp & 3) * 32 + lane;
        const int et = k * 128 + l;                               // index among the 256 epilogue threads
        const uint32_t tcol = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)k * kSlotCols;
        const size_t plane = (size_t)P.H * P.W;
        const int L = P.Hp * P.Wp;
        const int pix = k * 128 + l;                              // slot k = half k of the patch
        uint32_t dphase = 0;
        auto wait_d = [&]() { mbar_wait(d_bar(k), dphase); dphase ^= 1; tc_fence_after(); };
        auto release = [&]() {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_bar(k));
        };
        float x[kHx];
        // A = bf16(rmsnorm(x) * w): 64 channels -> 32 tensor-memory columns
        auto norm_to_a = [&](const float* w) {
            float sq = 0.f;
#pragma unroll
            for (int c = 0; c < kHx; ++c) sq = fmaf(x[c], x[c], sq);
            const float r = rsqrtf(sq * (1.0f / kHx) + 1e-6f);
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * i);
                pk[2 * i] = pack_bf2(x[4 * i] * r * w4.x, x[4 * i + 1] * r * w4.y);
                pk[2 * i + 1] = pack_bf2(x[4 * i + 2] * r * w4.z, x[4 * i + 3] * r * w4.w);
            }
            tmem_st32(tcol + kColA, pk);
            tmem_st_wait();
        };
        int q = 0;
        for (int n = 0; n < nlocal; ++n) {
            const int m = first + n * step;
            const int img = m / L, tok = m - img * L;
            const int py = tok / P.Wp, px = tok - py * P.Wp;
            const size_t off = (size_t)img * 3 * plane + (size_t)(py * 16 + (pix >> 4)) * P.W + (size_t)px * 16 + (pix & 15);
            {   // x = T[pixel] + Wrgb bf16(rgb)   (NerfEmbedder, fp32)
                const float r0 = round_bf(__ldg(P.x + off)), r1 = round_bf(__ldg(P.x + off + plane)),
                            r2 = round_bf(__ldg(P.x + off + 2 * plane));
                const float4* trow = reinterpret_cast<const float4*>(sTab + pix * kTabPitch);
                const float4* wr = reinterpret_cast<const float4*>(sWrgb);
#pragma unroll
                for (int i = 0; i < kHx / 4; ++i) {
                    const float4 tv = trow[i];
                    const float tt[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float4 w = wr[4 * i + e];
                        x[4 * i + e] = fmaf(w.x, r0, fmaf(w.y, r1, fmaf(w.z, r2, tt[e])));
                    }
                }
            }
            norm_to_a(sNormW);
            release();
            for (int j = 0; j < R; ++j, ++q) {
                const int b = q & 1;
                float* inv = sInv + b * (kHm + kHx);
                // ---- reciprocal column norms of the generated weights (F.normalize(dim=-2), eps 1e-12), one column per thread
                mbar_wait(w_full(b), (uint32_t)((q >> 1) & 1));
                if (et < kHm + kHx) {
                    const bool is1 = et < kHm;
                    const int col = is1 ? et : et - kHm, rows = is1 ? kHx : kHm;
                    const uint8_t* tile = gen + (sWB - base) + (uint32_t)b * kWBuf + (is1 ? 0u : kFc1Bytes);
                    float ss = 0.f;
                    for (int r = 0; r < rows; ++r) {
                        const float v = bf2f(*reinterpret_cast<const __nv_bfloat16*>(tile + sw32_offset(r, col, rows)));
                        ss = fmaf(v, v, ss);
                    }
                    inv[et] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                // ---- E1: t = silu(D / ||fc1_n||) -> A (128 channels -> 64 columns)
                wait_d();
#pragma unroll 1
                for (int ch = 0; ch < kHm / 32; ++ch) {
                    uint32_t d[32];
                    tmem_ld32(tcol + kColD + (uint32_t)(ch * 32), d);
                    tmem_ld_wait();
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 s4 = *reinterpret_cast<const float4*>(inv + ch * 32 + 4 * i);
                        pk[2 * i] = pack_bf2(silu_fast(__uint_as_float(d[4 * i]) * s4.x), silu_fast(__uint_as_float(d[4 * i + 1]) * s4.y));
                        pk[2 * i + 1] = pack_bf2(silu_fast(__uint_as_float(d[4 * i + 2]) * s4.z), silu_fast(__uint_as_float(d[4 * i + 3]) * s4.w));
                    }
                    tmem_st16(tcol + kColA + (uint32_t)(ch * 16), pk);
                }
                tmem_st_wait();
                release();
                // ---- E2: x += D / ||fc2_n||, then the next norm
                wait_d();
#pragma unroll
                for (int ch = 0; ch < kHx / 32; ++ch) {
                    uint32_t d[32];
                    tmem_ld32(tcol + kColD + (uint32_t)(ch * 32), d);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 s4 = *reinterpret_cast<const float4*>(inv + kHm + ch * 32 + 4 * i);
                        x[ch * 32 + 4 * i] = fmaf(__uint_as_float(d[4 * i]), s4.x, x[ch * 32 + 4 * i]);
                        x[ch * 32 + 4 * i + 1] = fmaf(__uint_as_float(d[4 * i + 1]), s4.y, x[ch * 32 + 4 * i + 1]);
                        x[ch * 32 + 4 * i + 2] = fmaf(__uint_as_float(d[4 * i + 2]), s4.z, x[ch * 32 + 4 * i + 2]);
                        x[ch * 32 + 4 * i + 3] = fmaf(__uint_as_float(d[4 * i + 3]), s4.w, x[ch * 32 + 4 * i + 3]);
                    }
                }
                norm_to_a(j + 1 < R ? sNormW + (j + 1) * kHx : sNormF);
                release();
            }
            // ---- final linear: 3 channels
            wait_d();
            uint32_t f[8];
            tmem_ld8(tcol + kColD, f);
            tmem_ld_wait();
            if (P.out_bf16) {
                __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(P.out) + off;
#pragma unroll
                for (int c = 0; c < 3; ++c) op[c * plane] = f2bf(__uint_as_float(f[c]) + sBiasF[c]);
            } else {
                float* op = reinterpret_cast<float*>(P.out) + off;
#pragma unroll
                for (int c = 0; c < 3; ++c) op[c * plane] = __uint_as_float(f[c]) + sBiasF[c];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kEpiWarps + kSlots) tmem_dealloc(tmem, 512);
}

}  // namespace nerf
}  // namespace deco

using namespace deco;

extern "C" int deco_nerf_decoder_blob_bytes(int num_nerf_blocks) { return (int)nerf::blob_bytes(num_nerf_blocks); }

// x fp32 [B,3,H,W]; params[j] = param_generator1 output of NerfBlock j, bf16 [B*L, 2 * 64 * 128] (fc1 [64 x 128] | fc2
// [128 x 64], row-major, un-normalised); blob packed by deco_b200/denoiser_pixnerd.py.  out [B,3,H,W] bf16 / fp32.
extern "C" int deco_nerf_decoder(const float* x, const void* const* params, int num_nerf_blocks, const void* blob,
                                 void* out, int out_is_bf16, int B, int H, int W, int patch, int hidden_x, int mlp_ratio,
                                 void* stream)
{
    using namespace deco::nerf;
    DECO_CHECK_ARG(x && params && blob && out, "nerf_decoder: null pointer");
    if (patch != 16 || hidden_x != kHx || hidden_x * mlp_ratio != kHm) {
        deco_set_error("nerf_decoder: built for patch_size 16, hidden_size_x 64, nerf_mlpratio 2 (got %d, %d, %d)", patch, hidden_x, mlp_ratio);
        return DECO_ERR_UNSUPPORTED;
    }
    DECO_CHECK_ARG(B > 0 && H % 16 == 0 && W % 16 == 0 && num_nerf_blocks >= 1 && num_nerf_blocks <= kMaxR,
                   "nerf_decoder: bad shape B=%d H=%d W=%d R=%d", B, H, W, num_nerf_blocks);
    Params P = {};
    P.x = x; P.blob = blob; P.out = out; P.out_bf16 = out_is_bf16;
    P.R = num_nerf_blocks; P.H = H; P.W = W; P.Hp = H / 16; P.Wp = W / 16;
    const long long tokens = (long long)B * P.Hp * P.Wp;
    DECO_CHECK_ARG(tokens < (1LL << 30), "nerf_decoder: too many tokens");
    P.tokens = (int)tokens;
    PFN_encodeTiled enc = get_tensormap_encoder();
    if (!enc) { deco_set_error("cuTensorMapEncodeTiled entry point not available"); return DECO_ERR_DRIVER; }
    Maps maps;
    const cuuint64_t row_bytes = (cuuint64_t)(2 * kHx * kHm) * 2;
    for (int j = 0; j < kMaxR; ++j) {
        const int jj = j < num_nerf_blocks ? j : 0;
        const char* p = reinterpret_cast<const char*>(params[jj]);
        DECO_CHECK_ARG(p && ((uintptr_t)p & 15) == 0, "nerf_decoder: params[%d] null or misaligned", jj);
        cuuint32_t estr[3] = {1, 1, 1};
        {   // fc1: [k = 64][n = 128] of every token row
            cuuint64_t dims[3] = {(cuuint64_t)kHm, (cuuint64_t)kHx, (cuuint64_t)tokens};
            cuuint64_t strides[2] = {(cuuint64_t)kHm * 2, row_bytes};
            cuuint32_t box[3] = {16, (cuuint32_t)kHx, 1};
            CUresult r = enc(&maps.fc1[j], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<char*>(p), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { deco_set_error("nerf_decoder: tensor map (fc1) failed: %d", (int)r); return DECO_ERR_DRIVER; }
        }
        {   // fc2: [k = 128][n = 64]
            cuuint64_t dims[3] = {(cuuint64_t)kHx, (cuuint64_t)kHm, (cuuint64_t)tokens};
            cuuint64_t strides[2] = {(cuuint64_t)kHx * 2, row_bytes};
            cuuint32_t box[3] = {16, (cuuint32_t)kHm, 1};
            CUresult r = enc(&maps.fc2[j], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<char*>(p + (size_t)kHx * kHm * 2), dims,
                             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { deco_set_error("nerf_decoder: tensor map (fc2) failed: %d", (int)r); return DECO_ERR_DRIVER; }
        }
    }
    const int smem = (int)smem_bytes(num_nerf_blocks);
    static unsigned long long attr_done = 0;
    if (!device_setup_done(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(nerf_decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { deco_set_error("nerf_decoder attr: %s", cudaGetErrorString(e)); return (int)e; }
        mark_device_setup(attr_done);
    }
    long long grid = device_sm_count();
    if (grid > tokens) grid = tokens;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = deco_pdl_enabled() ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, nerf_decoder_kernel, maps, P);
    if (e != cudaSuccess) { deco_set_error("nerf_decoder launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return DECO_OK;
}
