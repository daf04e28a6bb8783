// Non-causal multi-head attention, softmax(Q K^T / sqrt(d)) V, on the 5th-generation tensor cores.
//
// Replaces (reference, /root/reference): src/models/transformer/dit_c2i_DeCo.py:181-187 (transposes +
// scaled_dot_product_attention) and src/models/layers/attention_op.py:4; the second key/value segment covers the t2i
// joint attention (src/models/transformer/dit_t2i_pixnerd.py:52-59: keys = [image || text]).  q_norm / k_norm / RoPE
// (:178-180) are applied before this kernel runs (qknorm_rope_kernel, or the QKV GEMM epilogue).
//
// Persistent kernel, one CTA per SM, 12 warps:
//   warp 0      TMA producer: Q tiles and 128-key K/V blocks straight from the strided [tokens, 3H] QKV matrix through
//               4-D tensor maps (d, token, head, batch) with 16-column boxes and the 32-byte swizzle, i.e. directly in the
//               chunk-major UMMA operand layout (sw32_offset).  Out-of-range coordinates are zero-filled by the TMA unit:
//               that pads head dim 72 -> 80 for the QK^T contraction and zeroes ragged sequence tails.
//   warp 1      MMA issuer (whole warp runs the control flow, one elected lane issues): S = Q K^T (tcgen05.mma SS) and
//               O += P V (tcgen05.mma TS: A = P in tensor memory, B = V MN-major); completion via tcgen05.commit.
//   warps 4-7   softmax warpgroup for query tile 0, warps 8-11 for query tile 1 (thread <-> query row <-> TMEM lane):
//               S row from TMEM, exact maximum, exp2, row sum, P (bf16) back to TMEM, final O / l staged in the item's
//               last (by then dead) K/V stage and written with one TMA store per tile.  Q tiles retire with the item's
//               last Q K^T, so the producer prefetches up to two items ahead.
// A work item is (batch row, head, 256 queries).  Keys are consumed in HALF blocks of 64: per tile the score buffer is
// double buffered in tensor memory (S_a | S_b | P | O = 64 + 64 + 32 + 80 columns), and the moment P_h arrives the
// issuer queues P_h V followed by S_{h+2} = Q K_{h+2}^T, so the next two score blocks are always in flight while the
// warpgroup runs its softmax: the exp2 (MUFU) pipe, not the MMA round trip, sets the pace.
// Online softmax without a per-block rescale of O: the exponent reference only moves when a block's maximum exceeds it
// by more than 2^8 (then O and the row sum are rescaled through TMEM); otherwise P carries values up to 256.  bf16 / fp32
// have the exponent range for that and the final 1/l uses the same reference, so the result is the exact softmax.
// Operand layouts were validated stand-alone by scripts/umma_probe.cu.
#include "tcgen05.cuh"
#include "tma_host.cuh"
#include <cstdlib>

namespace deco {

template <int D> struct TcAttnCfg {
    static constexpr int DP = (D + 15) / 16 * 16;   // head dim padded to the MMA K granularity (72 -> 80)
    static constexpr int NCH = DP / 16;             // 32-byte chunks per operand row = TMA boxes per tile
    static constexpr uint32_t CHUNK = 128 * 32;     // one [128 rows x 16 columns] box
    static constexpr uint32_t TILE = NCH * CHUNK;
    static constexpr int kKvStages = 3;
    static constexpr uint32_t oQ = 0;                         // [2 buffers][2 tiles]
    static constexpr uint32_t oKV = 4 * TILE;                 // [stages][K | V]
    static constexpr uint32_t oBar = oKV + kKvStages * 2 * TILE;
    static constexpr uint32_t SMEM = oBar + 256 + 1024;       // barriers + alignment slack
};

constexpr int kTcRows = 128;        // queries per tile = keys per K/V stage = TMEM lanes
constexpr int kTcHalf = 64;         // keys per softmax step
constexpr int kTcThreads = 384;
// tensor-memory columns of tile t (base t * 256): S buffers, P, O
constexpr uint32_t kColTile = 256, kColS = 0, kColP = 128, kColO = 160, kTcTmemCols = 512;
constexpr float kRescaleThreshold = 8.0f;   // log2 units

struct TcAttnMaps { CUtensorMap q, k0, v0, k1, v1, o; };

struct TcAttnParams {
    int Lq, heads, B;
    int Lk[2];
    float scale_log2;
    float* lse2;             // optional [B * heads, Lq]: log2-sum-exp2 of the scaled scores per query (for the backward)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        :: "l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
template <int kId>
__device__ __forceinline__ void named_bar_sync128() {
    asm volatile("bar.sync %0, 128;" :: "n"(kId) : "memory");
}

#ifdef DECO_ATTN_TRACE
__device__ long long g_attn_trace[3 * 2048];
#define ATTN_TRACE(region, code)                                                                  \
    do {                                                                                          \
        if (blockIdx.x == 0 && trace_n < 1023) {                                                  \
            g_attn_trace[(region) * 2048 + 2 * trace_n] = clock64();                              \
            g_attn_trace[(region) * 2048 + 2 * trace_n + 1] = (code);                             \
            ++trace_n;                                                                            \
        }                                                                                         \
    } while (0)
#else
#define ATTN_TRACE(region, code) do {} while (0)
#endif

template <int D>
__global__ void __launch_bounds__(kTcThreads, 1)
attention_tc_kernel(const __grid_constant__ TcAttnMaps M, const TcAttnParams P)
{
    using C = TcAttnCfg<D>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));

    const uint32_t bars = base + C::oBar;
    auto q_full = [&](int i) { return bars + 8u * i; };                         // 2
    auto q_empty = [&](int i) { return bars + 16u + 8u * i; };                  // 2
    auto kv_full = [&](int i) { return bars + 32u + 8u * i; };                  // kKvStages
    auto kv_empty = [&](int i) { return bars + 64u + 8u * i; };                 // kKvStages
    auto bar_s = [&](int t, int buf) { return bars + 96u + 8u * (2 * t + buf); };   // S_g[t] complete (buffer g & 1)
    auto bar_p = [&](int t) { return bars + 128u + 8u * t; };                   // P_g[t] written (128 arrivals)
    auto bar_o = [&](int t) { return bars + 144u + 8u * t; };                   // P_g V complete
    auto o_empty = [&](int t) { return bars + 160u + 8u * t; };                 // O[t] read out (128 arrivals)
    const uint32_t slot = bars + 176u;
    auto sQ = [&](int buf, int t) { return base + C::oQ + (uint32_t)(buf * 2 + t) * C::TILE; };
    auto sK = [&](int st) { return base + C::oKV + (uint32_t)(2 * st) * C::TILE; };
    auto sV = [&](int st) { return base + C::oKV + (uint32_t)(2 * st + 1) * C::TILE; };

    const int tid = threadIdx.x, warp = tid >> 5;
    // half blocks (64 keys) per segment / item; K/V stages (128 keys) per item
    const int nh0 = (P.Lk[0] + kTcHalf - 1) / kTcHalf, nh1 = (P.Lk[1] + kTcHalf - 1) / kTcHalf;
    const int nb0 = (nh0 + 1) >> 1, nb1 = (nh1 + 1) >> 1;
    const int nh = nh0 + nh1, nblk = nb0 + nb1;
    const int npair = (P.Lq + 2 * kTcRows - 1) / (2 * kTcRows);
    const int nitems = P.B * P.heads * npair;
    const int first = blockIdx.x, step = gridDim.x;
    const int nlocal = first < nitems ? (nitems - first + step - 1) / step : 0;
#ifdef DECO_ATTN_TRACE
    int trace_n = 0;
#endif

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(q_full(i), 1); mbar_init(q_empty(i), 1); }
        for (int i = 0; i < C::kKvStages; ++i) { mbar_init(kv_full(i), 1); mbar_init(kv_empty(i), 2); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_s(t, 0), 1); mbar_init(bar_s(t, 1), 1);
            mbar_init(bar_p(t), 128); mbar_init(bar_o(t), 1); mbar_init(o_empty(t), 128);
        }
        fence_barrier_init();
        tma_prefetch_desc(&M.q); tma_prefetch_desc(&M.k0); tma_prefetch_desc(&M.v0); tma_prefetch_desc(&M.o);
    }
    __syncwarp();
    if (warp == 2) tmem_alloc(slot, kTcTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                   // prologue overlapped the previous kernel's tail; q / k / v are visible from here
    pdl_launch_dependents();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));

    auto decode = [&](int n, int& b, int& h, int& pair) {
        const int it = first + n * step;
        pair = it % npair;
        h = (it / npair) % P.heads;
        b = it / (npair * P.heads);
    };
    // half block hl of an item -> K/V block inside the item and the 64-key half of that block
    auto locate = [&](int hl, int& blk, int& half) {
        const int loc = hl < nh0 ? hl : hl - nh0;
        blk = (hl < nh0 ? 0 : nb0) + (loc >> 1);
        half = loc & 1;
    };

    if (warp == 0) {
        // ================================================================== TMA producer
        int gkv = 0;
        for (int n = 0; n < nlocal; ++n) {
            int b, h, pair;
            decode(n, b, h, pair);
            const int qb = n & 1, fq = n >> 1;
            if (fq > 0) mbar_wait(q_empty(qb), (fq - 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(q_full(qb), 2 * C::TILE);
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int c = 0; c < C::NCH; ++c)
                        tma_load_4d(sQ(qb, t) + c * C::CHUNK, &M.q, q_full(qb), 16 * c, pair * 256 + t * kTcRows, h, b);
            }
            __syncwarp();
            for (int j = 0; j < nblk; ++j, ++gkv) {
                const int st = gkv % C::kKvStages, f = gkv / C::kKvStages;
                if (f > 0) mbar_wait(kv_empty(st), (f - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(kv_full(st), 2 * C::TILE);
                    const bool seg1 = j >= nb0;
                    const CUtensorMap* mk = seg1 ? &M.k1 : &M.k0;
                    const CUtensorMap* mv = seg1 ? &M.v1 : &M.v0;
                    const int row = (seg1 ? j - nb0 : j) * kTcRows;
#pragma unroll
                    for (int c = 0; c < C::NCH; ++c) tma_load_4d(sK(st) + c * C::CHUNK, mk, kv_full(st), 16 * c, row, h, b);
#pragma unroll
                    for (int c = 0; c < C::NCH; ++c) tma_load_4d(sV(st) + c * C::CHUNK, mv, kv_full(st), 16 * c, row, h, b);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        // All index arithmetic is incremental (no divisions) and the shared-memory descriptors are derived from
        // per-buffer bases by adding constants: the issue path must stay far shorter than a 64-key softmax step.
        constexpr uint32_t idesc_s = make_idesc_major(128, kTcHalf, 0, 0);
        constexpr uint32_t idesc_o = make_idesc_major(128, C::DP, 0, 1);
        constexpr uint64_t kChunkStep = C::CHUNK >> 4;          // descriptor address field counts 16-byte units
        const uint64_t dQ0 = make_umma_desc(sQ(0, 0), 16, 256, 6);
        const uint64_t dK0 = make_umma_desc(sK(0), 16, 256, 6);
        const uint64_t dV0 = make_umma_desc(sV(0), kTcRows * 32, 256, 6);
        const int K = nlocal * nh;          // flattened (item, half block) sequence of this CTA
        struct Cursor { int n, hl, gkv_item, st, par; };   // position, CTA-global K/V block of the item's first block
        auto blk_half = [&](int hl, int& blk, int& half) {
            const int loc = hl < nh0 ? hl : hl - nh0;
            blk = (hl < nh0 ? 0 : nb0) + (loc >> 1);
            half = loc & 1;
        };
        int kv_waited = -1, q_waited = -1;
        Cursor cs = {0, 0, 0, 0, 0};        // cursor of the next S to issue
        // S_g[t] (both tiles) for the cursor position; g = CTA-global half-block index
        auto issue_s_wait = [&](int& stage, int& half) {
            int blk;
            blk_half(cs.hl, blk, half);
            const int gkv = cs.gkv_item + blk;
            stage = gkv % C::kKvStages;
            if (gkv > kv_waited) { mbar_wait(kv_full(stage), (gkv / C::kKvStages) & 1); kv_waited = gkv; }
            if (cs.n > q_waited) { mbar_wait(q_full(cs.n & 1), (cs.n >> 1) & 1); q_waited = cs.n; }
        };
        auto issue_s_mma = [&](int t, int g, int stage, int half) {
            const uint64_t da = dQ0 + (uint64_t)((cs.n & 1) * 2 + t) * (C::TILE >> 4);
            const uint64_t db = dK0 + (uint64_t)(2 * stage) * (C::TILE >> 4) + (uint64_t)half * ((kTcHalf * 32) >> 4);
            const uint32_t dcol = tmem + (uint32_t)t * kColTile + kColS + (uint32_t)(g & 1) * kTcHalf;
            if (elect_one()) {
#pragma unroll
                for (int kc = 0; kc < C::NCH; ++kc)
                    umma_bf16(dcol, da + kc * kChunkStep, db + kc * kChunkStep, idesc_s, kc ? 1u : 0u);
                umma_commit(bar_s(t, g & 1));
                if (t == 1 && cs.hl == nh - 1) umma_commit(q_empty(cs.n & 1));   // last S of the item: Q tiles retire with it
            }
            __syncwarp();
        };
        auto advance_s = [&]() {
            if (++cs.hl == nh) { cs.hl = 0; ++cs.n; cs.gkv_item += nblk; }
        };
        for (int g = 0; g < 2 && g < K; ++g) {
            int stage, half;
            issue_s_wait(stage, half);
            tc_fence_after();
            issue_s_mma(0, g, stage, half);
            issue_s_mma(1, g, stage, half);
            advance_s();
        }
        int n = 0, hl = 0, gkv_item = 0;
        for (int k = 0; k < K; ++k) {
            int blk, half;
            blk_half(hl, blk, half);
            const int st = (gkv_item + blk) % C::kKvStages;
            bool last_of_stage = true;
            if (hl + 1 < nh) { int b2, h2; blk_half(hl + 1, b2, h2); last_of_stage = b2 != blk; }
            int s_stage = 0, s_half = 0;
            const bool more = k + 2 < K;
            if (more) issue_s_wait(s_stage, s_half);   // operands of S_{k+2}: waited for ahead of the P wait
            const uint64_t dv = dV0 + (uint64_t)(2 * st) * (C::TILE >> 4) + (uint64_t)(half * 4) * (512 >> 4);
            for (int t = 0; t < 2; ++t) {
                mbar_wait(bar_p(t), k & 1);                                  // P_k[t] in tensor memory (=> S_k[t] was read)
                if (hl == 0 && n > 0) mbar_wait(o_empty(t), (n - 1) & 1);    // previous item's O[t] has been read out
                tc_fence_after();
                ATTN_TRACE(2, 10 + t);
                if (elect_one()) {
                    const uint32_t ocol = tmem + (uint32_t)t * kColTile + kColO;
                    const uint32_t pcol = tmem + (uint32_t)t * kColTile + kColP;
#pragma unroll
                    for (int ks = 0; ks < kTcHalf / 16; ++ks)
                        umma_bf16_ts(ocol, pcol + (uint32_t)(ks * 8), dv + (uint64_t)ks * (512 >> 4), idesc_o, (hl > 0 || ks > 0) ? 1u : 0u);
                    umma_commit(bar_o(t));
                    // every MMA reading this K/V stage is issued: hand it back (two arrivals).  The item's LAST stage is
                    // handed back by the softmax warpgroups instead, which stage their output tiles in it.
                    if (t == 1 && last_of_stage && hl + 1 < nh) { umma_commit(kv_empty(st)); umma_commit(kv_empty(st)); }
                }
                __syncwarp();
                ATTN_TRACE(2, 20 + t);
                if (more) issue_s_mma(t, k + 2, s_stage, s_half);
                ATTN_TRACE(2, 40 + t);
            }
            if (more) advance_s();
            if (++hl == nh) { hl = 0; ++n; gkv_item += nblk; }
        }
    } else if (warp >= 4) {
        // ================================================================== softmax warpgroups
        const int t = (warp - 4) >> 2;                // query tile of this warpgroup
        const int r = tid - 128 - t * 128;            // row inside the tile = TMEM lane
        const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)t * kColTile;
        const uint32_t tP = tbase + kColP, tO = tbase + kColO;
        const float c = P.scale_log2;
        int g = 0;                                    // CTA-global half-block counter of this tile
        int store_stage = -1;
        for (int n = 0; n < nlocal; ++n) {
            float m_used = 0.f, l = 0.f;
            for (int hl = 0; hl < nh; ++hl, ++g) {
                if (r == 0) ATTN_TRACE(t, 1);
                mbar_wait(bar_s(t, g & 1), (g >> 1) & 1);
                tc_fence_after();
                if (r == 0) ATTN_TRACE(t, 2);
                uint32_t s[2][32];
                tmem_ld32(tbase + kColS + (uint32_t)((g & 1) * kTcHalf), s[0]);
                tmem_ld32(tbase + kColS + (uint32_t)((g & 1) * kTcHalf + 32), s[1]);
                tmem_ld_wait();
                if (r == 0) ATTN_TRACE(t, 3);
                {   // ragged tail of the key segment
                    const int seg = hl < nh0 ? 0 : 1;
                    const int valid = P.Lk[seg] - (seg ? hl - nh0 : hl) * kTcHalf;
                    if (valid < kTcHalf) {
#pragma unroll
                        for (int qd = 0; qd < 2; ++qd)
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (qd * 32 + i >= valid) s[qd][i] = 0xff800000u;   // -inf
                    }
                }
                float mx[4];
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    mx[qd] = __uint_as_float(s[qd >> 1][(qd & 1) * 16]);
#pragma unroll
                    for (int i = 1; i < 16; ++i) mx[qd] = fmaxf(mx[qd], __uint_as_float(s[qd >> 1][(qd & 1) * 16 + i]));
                }
                const float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
                if (hl > 0) {        // P V of the previous half block must have drained the P buffer (and O is quiescent)
                    mbar_wait(bar_o(t), (g - 1) & 1);
                    tc_fence_after();
                }
                if (hl == 0) {
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
                            tmem_ld16(tO + (uint32_t)c0, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                            tmem_st16(tO + (uint32_t)c0, o);
                        }
                    }
                }
                const float mc = m_used * c;
                float sum[2] = {0.f, 0.f};
#pragma unroll
                for (int qd = 0; qd < 2; ++qd) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(s[qd][2 * i]), c, -mc));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(s[qd][2 * i + 1]), c, -mc));
                        sum[0] += p0; sum[1] += p1;
                        pk[i] = pack_bf2(p0, p1);
                    }
                    tmem_st16(tP + (uint32_t)(qd * 16), pk);   // one column = two keys
                }
                l += sum[0] + sum[1];
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_p(t));
                if (r == 0) ATTN_TRACE(t, 4);
                if (store_stage >= 0) {             // the previous item's TMA store has long since read its staging tile
                    if (r == 0) { tma_store_wait_read(); mbar_arrive(kv_empty(store_stage)); }
                    store_stage = -1;
                }
            }
            // ---------------------------------------------------------------- O / l -> shared memory -> TMA store
            mbar_wait(bar_o(t), (g - 1) & 1);
            tc_fence_after();
            if (r == 0) ATTN_TRACE(t, 5);
            uint32_t o[D];
#pragma unroll
            for (int c0 = 0; c0 + 32 <= D; c0 += 32) tmem_ld32(tO + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&o[c0]));
            if constexpr (D % 32 == 8) tmem_ld8(tO + (uint32_t)(D - 8), *reinterpret_cast<uint32_t(*)[8]>(&o[D - 8]));
            static_assert(D % 32 == 0 || D % 32 == 8, "epilogue covers head dims 64 and 72");
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(o_empty(t));
            if (r == 0) ATTN_TRACE(t, 6);
            const float inv = 1.0f / l;
            if (P.lse2) {            // softmax statistics for attention_bwd.cu: log2(sum_k exp2(c s_k)) = c m + log2(l)
                int b_, h_, pair_;
                decode(n, b_, h_, pair_);
                const int qrow = pair_ * 256 + t * kTcRows + r;
                if (qrow < P.Lq) P.lse2[((long long)b_ * P.heads + h_) * P.Lq + qrow] = fmaf(m_used, c, log2f(l));
            }
            // staging tile: the item's last K/V stage is dead once its P V has completed (K half for tile 0, V half for
            // tile 1); dense [128][D] bf16 rows, written out by one TMA store
            const int st_last = (n * nblk + nblk - 1) % C::kKvStages;
            const uint32_t stg_addr = t == 0 ? sK(st_last) : sV(st_last);
            uint8_t* stg = gen + (stg_addr - base) + (uint32_t)r * (D * 2);
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 8) {
                uint4 w4;
                w4.x = pack_bf2(__uint_as_float(o[c0 + 0]) * inv, __uint_as_float(o[c0 + 1]) * inv);
                w4.y = pack_bf2(__uint_as_float(o[c0 + 2]) * inv, __uint_as_float(o[c0 + 3]) * inv);
                w4.z = pack_bf2(__uint_as_float(o[c0 + 4]) * inv, __uint_as_float(o[c0 + 5]) * inv);
                w4.w = pack_bf2(__uint_as_float(o[c0 + 6]) * inv, __uint_as_float(o[c0 + 7]) * inv);
                *reinterpret_cast<uint4*>(stg + c0 * 2) = w4;
            }
            fence_proxy_async();
            if (t == 0) named_bar_sync128<1>(); else named_bar_sync128<2>();
            if (r == 0) {
                int b, h, pair;
                decode(n, b, h, pair);
                tma_store_4d(&M.o, stg_addr, 0, pair * 256 + t * kTcRows, h, b);
                tma_store_commit();
            }
            store_stage = st_last;      // handed back to the producer after the next softmax step (or at the end)
            if (r == 0) ATTN_TRACE(t, 7);
        }
        if (r == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, kTcTmemCols);
}

// 4-D view (d, token, head, batch) of a strided [B*L, row_stride] bf16 matrix whose columns are [head][d]
// head_pitch = elements between consecutive heads inside a row (head_dim for a dense [head][d] row; 80 for the padded
// layout the QKV GEMM writes for head_dim 72, which keeps every 16-column TMA box inside one 32-byte sector)
static int make_attn_tmap(CUtensorMap* map, const void* ptr, int D, long long L, int heads, int B, long long row_stride,
                          int box_d, CUtensorMapSwizzle swz, long long head_pitch = 0) {
    if (head_pitch <= 0) head_pitch = D;
    static int promo = -1;
    if (promo < 0) { const char* e = getenv("DECO_ATTN_L2PROMO"); promo = e ? atoi(e) : 128; }
    const CUtensorMapL2promotion l2p = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    PFN_encodeTiled enc = get_tensormap_encoder();
    if (!enc) { deco_set_error("cuTensorMapEncodeTiled entry point not available"); return DECO_ERR_DRIVER; }
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)heads, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)row_stride * 2, (cuuint64_t)head_pitch * 2, (cuuint64_t)L * (cuuint64_t)row_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_d, (cuuint32_t)kTcRows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swz, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { deco_set_error("attention: cuTensorMapEncodeTiled failed: %d", (int)rc); return DECO_ERR_DRIVER; }
    return DECO_OK;
}

template <int D>
static int launch_attention_tc(const TcAttnMaps& M, const TcAttnParams& P, cudaStream_t st) {
    using C = TcAttnCfg<D>;
    static unsigned long long attr_done = 0;
    if (!device_setup_done(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) { deco_set_error("attention attr: %s", cudaGetErrorString(e)); return (int)e; }
        mark_device_setup(attr_done);
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = kNumSMs;
    const long long items = (long long)P.B * P.heads * ((P.Lq + 2 * kTcRows - 1) / (2 * kTcRows));
    const int grid = (int)(items < sms ? items : sms);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = deco_pdl_enabled() ? 1 : 0;
    cudaError_t le = cudaLaunchKernelEx(&cfg, attention_tc_kernel<D>, M, P);
    if (le != cudaSuccess) { deco_set_error("attention launch failed: %s", cudaGetErrorString(le)); return (int)le; }
    DECO_CHECK_LAUNCH("attention_tc_kernel");
    return DECO_OK;
}

}  // namespace deco

static int attention_fwd_impl(const void* q, long long q_stride,
                              const void* k0, const void* v0, long long kv0_stride, int Lk0,
                              const void* k1, const void* v1, long long kv1_stride, int Lk1,
                              void* out, long long out_stride, float* lse2_out,
                              int B, int heads, int Lq, int head_dim, float scale, void* stream,
                              int q_pitch = 0, int kv0_pitch = 0, int kv1_pitch = 0)
{
    using namespace deco;
    DECO_CHECK_ARG(q && k0 && v0 && out, "attention: null pointer");
    DECO_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk0 > 0 && Lk1 >= 0, "attention: bad shape");
    DECO_CHECK_ARG(Lk1 == 0 || (k1 && v1), "attention: second key/value segment is null");
    DECO_CHECK_ARG(head_dim == 64 || head_dim == 72, "attention: head_dim %d not built (64, 72)", head_dim);
    DECO_CHECK_ARG(q_stride % 8 == 0 && kv0_stride % 8 == 0 && out_stride % 8 == 0 && (Lk1 == 0 || kv1_stride % 8 == 0),
                   "attention: row strides must keep 16-byte alignment");
    DECO_CHECK_ARG(((uintptr_t)q | (uintptr_t)k0 | (uintptr_t)v0 | (uintptr_t)out | (uintptr_t)k1 | (uintptr_t)v1) % 16 == 0,
                   "attention: pointers must be 16-byte aligned");
    TcAttnMaps M;
    int rc;
    const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_32B;
    DECO_CHECK_ARG((q_pitch == 0 || (q_pitch >= head_dim && q_pitch % 8 == 0)) && (kv0_pitch == 0 || (kv0_pitch >= head_dim && kv0_pitch % 8 == 0)) &&
                   (kv1_pitch == 0 || (kv1_pitch >= head_dim && kv1_pitch % 8 == 0)), "attention: head pitch must be >= head_dim and a multiple of 8");
    if ((rc = make_attn_tmap(&M.q, q, head_dim, Lq, heads, B, q_stride, 16, sw, q_pitch))) return rc;
    if ((rc = make_attn_tmap(&M.k0, k0, head_dim, Lk0, heads, B, kv0_stride, 16, sw, kv0_pitch))) return rc;
    if ((rc = make_attn_tmap(&M.v0, v0, head_dim, Lk0, heads, B, kv0_stride, 16, sw, kv0_pitch))) return rc;
    if (Lk1) {
        if ((rc = make_attn_tmap(&M.k1, k1, head_dim, Lk1, heads, B, kv1_stride, 16, sw, kv1_pitch))) return rc;
        if ((rc = make_attn_tmap(&M.v1, v1, head_dim, Lk1, heads, B, kv1_stride, 16, sw, kv1_pitch))) return rc;
    } else {
        M.k1 = M.k0; M.v1 = M.v0;
    }
    if ((rc = make_attn_tmap(&M.o, out, head_dim, Lq, heads, B, out_stride, head_dim, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    TcAttnParams P;
    P.Lq = Lq; P.heads = heads; P.B = B; P.Lk[0] = Lk0; P.Lk[1] = Lk1;
    P.scale_log2 = scale * 1.4426950408889634f;
    P.lse2 = lse2_out;
    if (head_dim == 72) return launch_attention_tc<72>(M, P, (cudaStream_t)stream);
    return launch_attention_tc<64>(M, P, (cudaStream_t)stream);
}

extern "C" int deco_attention_fwd(const void* q, long long q_stride,
                                  const void* k0, const void* v0, long long kv0_stride, int Lk0,
                                  const void* k1, const void* v1, long long kv1_stride, int Lk1,
                                  void* out, long long out_stride,
                                  int B, int heads, int Lq, int head_dim, float scale, void* stream)
{
    return attention_fwd_impl(q, q_stride, k0, v0, kv0_stride, Lk0, k1, v1, kv1_stride, Lk1, out, out_stride, nullptr,
                              B, heads, Lq, head_dim, scale, stream);
}

// The same with explicit head pitches (elements between consecutive heads of a q / k / v row; 0 = head_dim): the fused QKV
// GEMM writes head_dim-72 heads at a pitch of 80 so that every 16-column operand box is one aligned 32-byte sector.
extern "C" int deco_attention_fwd_pitched(const void* q, long long q_stride, int q_head_pitch,
                                          const void* k0, const void* v0, long long kv0_stride, int kv0_head_pitch, int Lk0,
                                          const void* k1, const void* v1, long long kv1_stride, int kv1_head_pitch, int Lk1,
                                          void* out, long long out_stride,
                                          int B, int heads, int Lq, int head_dim, float scale, void* stream)
{
    return attention_fwd_impl(q, q_stride, k0, v0, kv0_stride, Lk0, k1, v1, kv1_stride, Lk1, out, out_stride, nullptr,
                              B, heads, Lq, head_dim, scale, stream, q_head_pitch, kv0_head_pitch, kv1_head_pitch);
}

// Training forward: also writes lse2_out [B * heads, Lq] = log2 sum_k exp2(scale log2(e) q.k), which lets the backward
// (deco_attention_bwd with have_lse = 1) skip rebuilding the softmax statistics.
extern "C" int deco_attention_fwd_lse(const void* q, long long q_stride, const void* k, const void* v, long long kv_stride,
                                      int Lk, void* out, long long out_stride, float* lse2_out,
                                      int B, int heads, int Lq, int head_dim, float scale, void* stream)
{
    if (!lse2_out) { deco_set_error("attention_fwd_lse: null lse pointer"); return DECO_ERR_ARG; }
    return attention_fwd_impl(q, q_stride, k, v, kv_stride, Lk, nullptr, nullptr, 0, 0, out, out_stride, lse2_out,
                              B, heads, Lq, head_dim, scale, stream);
}
