// tcgen05 GEMMs whose epilogues absorb the memory-bound neighbours of a DiT block, so that between two GEMMs of
// /root/reference/src/models/transformer/dit_c2i_DeCo.py:206-210 (FlattenDiTBlock.forward) nothing but the attention
// kernel touches HBM:
//
//   FE_STREAM       x <- [x + gate *] (A.W^T + bias)                       the fp32 residual stream update (:208,:209)
//                   + per-row partial sums of squares of the new x          (statistics of the NEXT RMSNorm, :94-99)
//                   + xg = bf16(x * norm_w * (1 + scale_next))              (the next modulate(), :11-12, minus its
//                                                                            row factor rstd and its shift)
//   FE_NORM_QKV     y = rstd[row] * (xg.W^T) + (shift.W^T)[image]           == Linear(modulate(RMSNorm(x)))  (:176)
//                   then per-head q_norm / k_norm + 2-D RoPE (:178-180), bf16 out.  BN = 2 heads.
//   FE_NORM_SWIGLU  same normalisation, then silu(a) * b on the interleaved [16 x w1 | 16 x w3] columns (:113)
//
// RMSNorm commutes with the GEMM because rstd is a per-row scalar:  (rstd * x * g + sh) . W^T = rstd * ((x*g).W^T) +
// sh.W^T ; the second term is one tiny GEMM per block and image batch ([B', H] x [H, N]).
//
// Main loop = the one of gemm_tcgen05.cu (persistent, 2-CTA pairs computing 256 x BN tiles, TMA SW128 ring, one-thread
// tcgen05.mma issue, double-buffered TMEM accumulators).  Epilogue: tcgen05.ld 32x32b gives every thread one ROW, so
// row statistics, per-head norms and RoPE are thread-local; global I/O goes through shared memory and TMA
// (residual tile prefetched with cp.async.bulk.tensor while the MMAs run, results stored with bulk tensor stores), so
// the 128 epilogue threads never issue an uncoalesced global access.
#include "tcgen05.cuh"
#include "tma_host.cuh"
#include <cstdlib>
#include <type_traits>

namespace deco {
namespace fused {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kThreads = 256;
constexpr int kAccStages = 2;
constexpr int kMaxSsqParts = 16;   // column tiles of the producing FE_STREAM GEMM (hidden <= 16 x 128)

// Both FE_STREAM variants stream the fp32 residual through a per-warp RING of 32 x 32 chunk buffers (TMA in, update in
// place, TMA out): the load of chunk g + R - 1 is issued the moment the store of chunk g - 1 has released its buffer, so
// residual traffic never stops at tile boundaries.  FE_STREAM: R = chunks per tile (a tile of look-ahead; the K = hidden
// GEMMs, whose epilogue is HBM-bound).  FE_STREAM_RING: R = 3, 48 KB less staging = two more operand stages and
// 256-wide tiles, for the deep-K (w2) GEMM whose main loop is latency-bound on operand loads.
enum FusedEpi : int { FE_STREAM = 0, FE_NORM_QKV = 1, FE_NORM_SWIGLU = 2, FE_STREAM_RING = 3 };

struct Maps { CUtensorMap a, b, r, o, x; };   // A, W, residual in (fp32), stream out (fp32) / qkv out (bf16), xg out (bf16)

struct Params {
    int M, N, K;
    int L;                          // rows per image: modulation / shift row of `row` is row / L
    // ---- input normalisation (FE_NORM_*)
    const float* ssq_in;            // [ssq_parts][M] partial sums of squares, or null (rstd = 1)
    int ssq_parts;
    float inv_hidden, eps;
    const float* shw;               // fp32 [M / L, shw_stride] or null
    long long shw_stride;
    // ---- FE_STREAM
    const float* bias;              // [N] or null
    const __nv_bfloat16* gate;      // [M / L, gate_stride] or null
    long long gate_stride;
    int has_resid;
    const float* next_w;            // [N] next RMSNorm weight, null = no xg output
    const __nv_bfloat16* next_scale;   // [M / L, next_scale_stride]
    long long next_scale_stride;
    float* ssq_out;                 // [N / BN][M] or null
    __nv_bfloat16* xg;              // [M, ldx] bf16, written when next_w != null
    long long ldx;
    // ---- FE_NORM_QKV
    int seg_cols;                   // heads * head_dim: columns per q / k / v segment
    const float* seg_w[3];          // head-norm weight [d] of the segment, null = pass through
    int seg_rope[3];
    int out_pitch;                  // output columns per head (>= head_dim; 80 for head_dim 72 keeps q / k / v sector-aligned)
    const float2* rope;             // [L][d / 2] (cos, sin)
    int rope_wp;                    // > 0: axial table, tokens per image row (position = (tok / wp, tok % wp)); 0: generic
    float eps_head;
    // ---- FE_NORM_SWIGLU
    __nv_bfloat16* out;
    long long ldo;
};

template <int BN, int EPI> struct Cfg {
    static constexpr int kStageBytes = kBM * kBK * 2 + (BN / 2) * kBK * 2;
    // FE_NORM_QKV tiles hold whole heads: BN = 224 -> 3 heads of 72 (216 useful columns, 8 recomputed by the next tile),
    // BN = 256 -> 4 heads of 64.  Wide tiles matter: the operand fill rate (L2 -> SMEM) caps the MMA rate of narrow ones
    // (measured 850 / 1220 / 1590 TFLOP/s at BN = 128 / 192 / 256, profiles/gemm_bench_r1.txt).
    static constexpr int kHeadDim = EPI == FE_NORM_QKV ? (BN == 224 ? 72 : 64) : 0;
    static constexpr int kHeadsPerTile = EPI == FE_NORM_QKV ? (BN == 224 ? 3 : 4) : 0;
    static constexpr int kTileN = EPI == FE_NORM_QKV ? kHeadDim * kHeadsPerTile : BN;   // column stride between tiles
    // per-warp staging
    //   FE_STREAM      fp32 [BN/32 chunks][32 rows][128 B] (SW128, residual in / stream out) + 3 x [BN] fp32 coefficients
    //   FE_NORM_QKV    2 x bf16 [32 rows][d] (output) + [BN] fp32 shift product
    //   FE_NORM_SWIGLU [BN] fp32 shift product
    static constexpr bool kStream = EPI == FE_STREAM || EPI == FE_STREAM_RING;
    static constexpr int kRing = EPI == FE_STREAM ? BN / 32 : 3;
    static constexpr int kVecBytes = kStream ? 3 * BN * 4 : kTileN * 4;
    // FE_NORM_QKV stages whole heads; a head_dim-72 head may be written at a pitch of 80 columns (zero padded)
    static constexpr int kHeadPitchMax = kHeadDim == 72 ? 80 : kHeadDim;
    static constexpr int kOutStage = kStream ? kRing * 4096
                                   : EPI == FE_NORM_QKV ? 2 * 32 * kHeadPitchMax * 2 : 0;
    // layout: 4 x kOutStage (1024-aligned buffers), then 4 x kVecBytes
    static constexpr int kWarpAll = ((4 * (kOutStage + kVecBytes) + 1023) / 1024) * 1024;
    static_assert(kOutStage % 1024 == 0, "TMA staging buffers must stay 1024-byte aligned");
    // CTA-wide tables of FE_NORM_QKV: axial RoPE (cos, sin) rows, 144-byte pitch, and the 3 head-norm weight vectors
    static constexpr int kRopePitch = 144;
    static constexpr int kRopeMaxPos = 96;
    static constexpr int kTableBytes = EPI == FE_NORM_QKV ? kRopeMaxPos * kRopePitch + 3 * kHeadDim * 4 : 0;
    static constexpr int kEpiBytes = kWarpAll + ((kTableBytes + 1023) / 1024) * 1024;
    static constexpr int kBudget = 227 * 1024 - kEpiBytes - 1024 - 512;
    static constexpr int kStagesRaw = kBudget / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kTmemCols = (kAccStages * BN > 256) ? 512 : 256;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 512 /*barriers*/;
    static_assert(kStages >= 3, "not enough shared memory for the operand ring");
    static_assert(kStageBytes % 1024 == 0, "stage must keep the 1024-byte swizzle alignment");
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_fused_kernel(const __grid_constant__ Maps maps, const Params P)
{
    using C = Cfg<BN, EPI>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = smem_base + C::kStages * C::kStageBytes;
    const uint32_t bar_base = epi_base + C::kEpiBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + kAccStages + s); };
    auto resid_bar = [&](int w) { return bar_base + 8u * (2 * C::kStages + 2 * kAccStages + w); };
    auto ring_bar = [&](int w, int b) { return bar_base + 8u * (2 * C::kStages + 2 * kAccStages + 4 + w * 8 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 2 * kAccStages + 4 + 32);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;
    const int num_m = (P.M + 2 * kBM - 1) / (2 * kBM), num_n = (P.N + C::kTileN - 1) / C::kTileN;
    const int num_tiles = num_m * num_n;
    const int num_k = (P.K + kBK - 1) / kBK;
    const int tile0 = blockIdx.x / 2, tile_stride = gridDim.x / 2;
    // columns of the last column tile (stream GEMMs only; a multiple of 32, else the full-width MMA runs on zero padding)
    const int n_rem = P.N - (num_n - 1) * C::kTileN;
    const int n_last = (C::kStream && n_rem < BN && n_rem % 32 == 0) ? n_rem : BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a);
        tma_prefetch_desc(&maps.b);
        if (EPI != FE_NORM_SWIGLU) tma_prefetch_desc(&maps.o);
        if (C::kStream) tma_prefetch_desc(&maps.r);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); }
        for (int w = 0; w < 4; ++w) { mbar_init(resid_bar(w), 1); for (int b = 0; b < 8; ++b) mbar_init(ring_bar(w, b), 1); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2sm(tmem_slot, C::kTmemCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                   // everything above overlapped the previous kernel's tail; its outputs are visible from here
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
                const int m_blk = tile / num_n, n_blk = tile % num_n;
                const int arow = (m_blk * 2 + (int)cta_rank) * kBM;
                // a ragged last column tile of the stream GEMMs runs a NARROWER MMA (N = what is left) instead of multiplying
                // zero padding: each CTA of the pair then supplies n_last / 2 weight rows, at the head of its B tile
                const int brow = n_blk * C::kTileN + (int)cta_rank * ((C::kStream && n_blk == num_n - 1) ? n_last / 2 : BN / 2);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * C::kStageBytes;
                    const uint32_t sb = sa + kBM * kBK * 2;
                    if (is_leader) mbar_expect_tx(full_bar(stage), 2 * C::kStageBytes);
                    const uint32_t fb = mapa_shared(full_bar(stage), 0);
                    tma_load_2d_2sm(sa, &maps.a, fb, kb * kBK, arow);
                    tma_load_2d_2sm(sb, &maps.b, fb, kb * kBK, brow);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (lane == 0 && is_leader) {
            constexpr uint32_t idesc = make_idesc(2 * kBM, BN);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                const uint32_t idesc_t = (C::kStream && (tile % num_n) == num_n - 1) ? make_idesc(2 * kBM, n_last) : idesc;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * C::kStageBytes;
                    const uint64_t da = make_sw128_desc(sa), db = make_sw128_desc(sa + kBM * kBK * 2);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k)
                        umma_bf16_2sm(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_t, (kb | k) ? 1u : 0u);
                    umma_commit_2sm(empty_bar(stage));
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(tfull_bar(as));
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: thread = one accumulator row =====================
        // Everything a tile needs from global memory besides the accumulator (per-column / per-image vectors, row
        // statistics, the residual tile) is requested one tile AHEAD, so its latency hides behind the previous tile's
        // epilogue and this tile's MMAs; per-image vectors are parked in shared memory and re-read as broadcasts.
        // The `fast` instantiation of each tile body is branch-free; the other one covers ragged shapes (rows of one
        // warp spanning several images, arbitrary RoPE tables) with per-thread global loads.
        const int q = warp & 3;
        int as = 0; uint32_t aphase = 0;
        const uint32_t wstg = epi_base + q * C::kOutStage;
        const uint32_t vec = epi_base + 4 * C::kOutStage + q * C::kVecBytes;
        const bool uni = (P.L % 32) == 0;   // the 32 rows of a warp belong to one image
        auto release_acc = [&]() {
            // TMEM reads are complete (tcgen05.wait::ld) and fenced; the arrive carries no generic-memory data, so it
            // is relaxed: a release would drain this warp's outstanding global stores first (MEMBAR, ~1 us per tile)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(as), 0));
        };
        auto tile_row0 = [&](int tile) { return ((tile / num_n) * 2 + (int)cta_rank) * kBM + q * 32; };

        if constexpr (C::kStream) {
            constexpr int R = C::kRing;           // residual chunk buffers per warp
            constexpr int PEND = R >= 4 ? 2 : 1;  // stores allowed in flight before a buffer is recycled
            static_assert((BN / 32) % 2 == 0, "chunks are processed in pairs");
            constexpr int NC = BN / 32;
            constexpr int NV = BN / 64;        // column pairs per lane
            const bool fast = uni && P.has_resid;
            // chunk sequence number g lives in buffer g % R; the load of chunk g + R - 1 is issued when the store of
            // chunk g - 1 has released that buffer
            int ltile = tile0, lc = 0, lseq = 0, gseq = 0;
            auto issue_next_chunk = [&]() {
                if (ltile >= num_tiles) return;
                const int b = lseq % R;
                mbar_expect_tx(ring_bar(q, b), 4096);
                tma_load_2d(wstg + b * 4096, &maps.r, ring_bar(q, b), (ltile % num_n) * BN + lc * 32, tile_row0(ltile));
                ++lseq;
                if (++lc == NC) { lc = 0; ltile += tile_stride; }
            };
            // raw per-column operands of the coefficient vectors, fetched one tile ahead
            uint32_t rg[NV], rs_[NV];
            float2 rb[NV], rw[NV];
            auto fetch = [&](int tile) {
                const long long row = tile_row0(tile) + lane;
                const long long img = (row < P.M ? row : (long long)P.M - 1) / P.L;
                const int nb = (tile % num_n) * BN;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const int col = nb + 2 * (lane + 32 * k);
                    const bool ok = col < P.N;
                    rg[k] = (ok && P.gate) ? __ldg(reinterpret_cast<const uint32_t*>(P.gate + img * P.gate_stride + col)) : 0x3f803f80u;
                    rb[k] = (ok && P.bias) ? __ldg(reinterpret_cast<const float2*>(P.bias + col)) : make_float2(0.f, 0.f);
                    rw[k] = (ok && P.next_w) ? __ldg(reinterpret_cast<const float2*>(P.next_w + col)) : make_float2(0.f, 0.f);
                    rs_[k] = (ok && P.next_w) ? __ldg(reinterpret_cast<const uint32_t*>(P.next_scale + img * P.next_scale_stride + col)) : 0u;
                }
            };
            if (tile0 < num_tiles) {
                if (P.has_resid && lane == 0)
                    for (int b = 0; b < R; ++b) issue_next_chunk();
                if (uni) fetch(tile0);
            }
            for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
                const int n_blk = tile % num_n;
                const int row0 = tile_row0(tile);
                const long long row = row0 + lane;
                const long long img = (row < P.M ? row : (long long)P.M - 1) / P.L;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
                const int nbase = n_blk * BN;
                const __nv_bfloat16* grow = P.gate ? P.gate + img * P.gate_stride : nullptr;
                const __nv_bfloat16* srow = P.next_w ? P.next_scale + img * P.next_scale_stride : nullptr;
                if (uni) {
                    // coefficient vectors of this tile: G = gate, BG = bias * gate, NC = next_w * (1 + next_scale)
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        const uint32_t o = (uint32_t)(2 * (lane + 32 * k)) * 4;
                        const float2 g = unpack_bf2(rg[k]), sc = unpack_bf2(rs_[k]);
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(vec + o), "f"(g.x), "f"(g.y) : "memory");
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(vec + BN * 4 + o), "f"(rb[k].x * g.x), "f"(rb[k].y * g.y) : "memory");
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(vec + 2 * BN * 4 + o),
                                     "f"(rw[k].x * (1.0f + sc.x)), "f"(rw[k].y * (1.0f + sc.y)) : "memory");
                    }
                    __syncwarp();
                    if (tile + tile_stride < num_tiles) fetch(tile + tile_stride);
                }
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
                float ssq = 0.f;
                auto body = [&](auto fast_tag) {
                    constexpr bool F = decltype(fast_tag)::value;
                    // one chunk: acc = 32 accumulator columns of this thread's row
                    auto chunk = [&](const uint32_t (&acc)[32], int c) {
                        const int n0 = nbase + c * 32;           // may lie beyond N in a ragged last tile: zero operands, clipped stores
                        const int buf = gseq % R;
                        if (P.has_resid) mbar_wait(ring_bar(q, buf), (uint32_t)((gseq / R) & 1));
                        const uint32_t crow = wstg + buf * 4096 + lane * 128;
                        const bool live = n0 < P.N;              // warp-uniform (N % 32 == 0)
                        if (live) {
                            // shared-memory loads are issued in batches ahead of their uses: the warp issues in order,
                            // so a load consumed right away costs a full LDS latency per 4 columns
                            float4 rr[8];
#pragma unroll
                            for (int jq = 0; jq < 8; ++jq)
                                rr[jq] = (F || P.has_resid) ? lds128(crow + ((uint32_t)(jq ^ (lane & 7)) << 4))
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                float4 g4[4], bg4[4], nc4[4];
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const int ci = c * 32 + 4 * (4 * h + jj);    // column inside the tile
                                    const int col = nbase + ci;
                                    if (F || uni) {
                                        g4[jj] = lds128(vec + ci * 4);
                                        bg4[jj] = lds128(vec + BN * 4 + ci * 4);
                                        nc4[jj] = lds128(vec + 2 * BN * 4 + ci * 4);
                                    } else {
                                        g4[jj] = make_float4(1.f, 1.f, 1.f, 1.f);
                                        bg4[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
                                        nc4[jj] = bg4[jj];
                                        if (grow) {
                                            const uint2 gv = __ldg(reinterpret_cast<const uint2*>(grow + col));
                                            const float2 g0 = unpack_bf2(gv.x), g1 = unpack_bf2(gv.y);
                                            g4[jj] = make_float4(g0.x, g0.y, g1.x, g1.y);
                                        }
                                        if (P.bias) {
                                            const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + col));
                                            bg4[jj] = make_float4(b.x * g4[jj].x, b.y * g4[jj].y, b.z * g4[jj].z, b.w * g4[jj].w);
                                        }
                                        if (srow) {
                                            const float4 w = __ldg(reinterpret_cast<const float4*>(P.next_w + col));
                                            const uint2 sv = __ldg(reinterpret_cast<const uint2*>(srow + col));
                                            const float2 s0 = unpack_bf2(sv.x), s1 = unpack_bf2(sv.y);
                                            nc4[jj] = make_float4(w.x * (1.0f + s0.x), w.y * (1.0f + s0.y), w.z * (1.0f + s1.x), w.w * (1.0f + s1.y));
                                        }
                                    }
                                }
                                uint32_t xb[8];
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    const int jq = 4 * h + jj;
                                    float4 v;
                                    v.x = fmaf(__uint_as_float(acc[4 * jq]), g4[jj].x, bg4[jj].x) + rr[jq].x;
                                    v.y = fmaf(__uint_as_float(acc[4 * jq + 1]), g4[jj].y, bg4[jj].y) + rr[jq].y;
                                    v.z = fmaf(__uint_as_float(acc[4 * jq + 2]), g4[jj].z, bg4[jj].z) + rr[jq].z;
                                    v.w = fmaf(__uint_as_float(acc[4 * jq + 3]), g4[jj].w, bg4[jj].w) + rr[jq].w;
                                    ssq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ssq))));
                                    sts128(crow + ((uint32_t)(jq ^ (lane & 7)) << 4), v);
                                    xb[2 * jj] = pack_bf2(v.x * nc4[jj].x, v.y * nc4[jj].y);
                                    xb[2 * jj + 1] = pack_bf2(v.z * nc4[jj].z, v.w * nc4[jj].w);
                                }
                                if (srow && row < P.M) {
                                    uint4* xp = reinterpret_cast<uint4*>(P.xg + row * P.ldx + n0 + 16 * h);
                                    xp[0] = make_uint4(xb[0], xb[1], xb[2], xb[3]);
                                    xp[1] = make_uint4(xb[4], xb[5], xb[6], xb[7]);
                                }
                            }
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            if (live) tma_store_2d(&maps.o, wstg + buf * 4096, n0, row0);
                            bulk_commit();                       // (an empty group for a dead chunk keeps the ring's group count)
                            bulk_wait_read<PEND>();              // the store of chunk g - PEND has released its buffer
                            if (P.has_resid && gseq >= PEND) issue_next_chunk();
                        }
                        ++gseq;
                        __syncwarp();
                    };
                    // TMEM reads run one chunk ahead of the arithmetic (two register sets)
                    uint32_t accA[32], accB[32];
                    tmem_ld32(taddr, accA);
#pragma unroll 1
                    for (int c = 0; c < NC; c += 2) {
                        tmem_ld_wait();
                        tmem_ld32(taddr + (uint32_t)((c + 1) * 32), accB);
                        chunk(accA, c);
                        tmem_ld_wait();
                        if (c + 2 < NC) tmem_ld32(taddr + (uint32_t)((c + 2) * 32), accA);
                        else release_acc();                      // accumulator stage is free for the next-but-one tile
                        chunk(accB, c + 1);
                    }
                };
                if (fast) body(std::true_type{}); else body(std::false_type{});
                if (P.ssq_out && row < P.M) P.ssq_out[(long long)n_blk * P.M + row] = ssq;
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
            if (lane == 0) bulk_wait_all();
        } else if constexpr (EPI == FE_NORM_QKV) {
            constexpr int D = C::kHeadDim;
            constexpr int HPT = C::kHeadsPerTile;
            constexpr int TN = C::kTileN;
            const int kRowBytes = P.out_pitch * 2;    // staged row = one head at the output pitch (D or, for D = 72, 80 columns)
            constexpr int NV4 = (TN / 4 + 31) / 32;   // float4 of the shift product per lane
            // ---- CTA-wide tables: axial RoPE rows (x positions, y positions, one identity row), head-norm weights
            const uint32_t tbl = epi_base + C::kWarpAll;
            const uint32_t tblw = tbl + C::kRopeMaxPos * C::kRopePitch;
            const int Wp = P.rope_wp;
            const int Hp = Wp > 0 ? P.L / Wp : 0;
            const bool any_rope = (P.seg_rope[0] | P.seg_rope[1] | P.seg_rope[2]) != 0;
            const bool axial = any_rope && Wp > 0 && (P.L % Wp) == 0 && Wp + Hp < C::kRopeMaxPos;
            const int npos = axial ? Wp + Hp : 0;      // identity row
            const bool fast = (uni || !P.shw) && (axial || !any_rope);
            {
                const int et = threadIdx.x - 128;
                for (int idx = et; idx < (npos + 1) * (D / 4); idx += 128) {
                    const int pos = idx / (D / 4), k = idx % (D / 4);
                    float2 cs = make_float2(1.f, 0.f);
                    if (pos < npos)
                        cs = pos < Wp ? __ldg(P.rope + (long long)pos * (D / 2) + 2 * k)
                                      : __ldg(P.rope + (long long)(pos - Wp) * Wp * (D / 2) + 2 * k + 1);
                    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(tbl + pos * C::kRopePitch + k * 8), "f"(cs.x), "f"(cs.y) : "memory");
                }
                for (int idx = et; idx < 3 * D; idx += 128) {
                    const float* wp = P.seg_w[idx / D];
                    const float wv = wp ? __ldg(wp + idx % D) : 1.0f;
                    asm volatile("st.shared.f32 [%0], %1;" :: "r"(tblw + idx * 4), "f"(wv) : "memory");
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            int hcount = 0;
            float sq[kMaxSsqParts];
            float4 pre[NV4];
            auto fetch = [&](int tile) {
                const long long row = tile_row0(tile) + lane;
                const long long rowc = row < P.M ? row : (long long)P.M - 1;
                const int nb = (tile % num_n) * TN;
#pragma unroll
                for (int p = 0; p < kMaxSsqParts; ++p)
                    sq[p] = (P.ssq_in && p < P.ssq_parts) ? __ldg(P.ssq_in + (long long)p * P.M + rowc) : 0.f;
#pragma unroll
                for (int k = 0; k < NV4; ++k) {
                    const int ci = 4 * (lane + 32 * k);
                    pre[k] = (P.shw && uni && ci < TN && nb + ci < P.N)
                                 ? __ldg(reinterpret_cast<const float4*>(P.shw + (rowc / P.L) * P.shw_stride + nb + ci))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            if (tile0 < num_tiles) fetch(tile0);
            for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
                const int n_blk = tile % num_n;
                const int row0 = tile_row0(tile);
                const long long row = row0 + lane;
                const long long rowc = row < P.M ? row : (long long)P.M - 1;
                const long long img = rowc / P.L;
                const int tok = (int)(rowc % P.L);
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
                const int nbase = n_blk * TN;
                const float* shrow = P.shw ? P.shw + img * P.shw_stride : nullptr;
                float ssum = 0.f;
#pragma unroll
                for (int p = 0; p < kMaxSsqParts; ++p) ssum += sq[p];
                const float rstd = P.ssq_in ? rsqrtf(ssum * P.inv_hidden + P.eps) : 1.0f;
#pragma unroll
                for (int k = 0; k < NV4; ++k) {
                    const int ci = 4 * (lane + 32 * k);
                    if (ci < TN) sts128(vec + ci * 4, pre[k]);
                }
                __syncwarp();
                if (tile + tile_stride < num_tiles) fetch(tile + tile_stride);
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
                auto body = [&](auto fast_tag) {
                    constexpr bool F = decltype(fast_tag)::value;
#pragma unroll 1
                    for (int hh = 0; hh < HPT; ++hh) {
                        const int col0 = nbase + hh * D;
                        // a tile may straddle the q | k | v boundary: the segment is a property of the head
                        const int seg = min(col0 / P.seg_cols, 2);
                        const bool do_norm = P.seg_w[seg] != nullptr;
                        const bool do_rope = P.seg_rope[seg] != 0;
                        const float2* rp = do_rope ? P.rope + (long long)tok * (D / 2) : nullptr;
                        const uint32_t txrow = tbl + ((axial && do_rope) ? tok % Wp : npos) * C::kRopePitch;
                        const uint32_t tyrow = tbl + ((axial && do_rope) ? Wp + tok / Wp : npos) * C::kRopePitch;
                        const uint32_t wrow = tblw + seg * D * 4;
                        const int hb = hcount & 1;               // output staging buffer (double buffered across heads)
                        ++hcount;
                        uint32_t acc[D];
                        {
                            uint32_t t32[32];
                            tmem_ld32(taddr + (uint32_t)(hh * D), t32);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) acc[i] = t32[i];
                            tmem_ld32(taddr + (uint32_t)(hh * D + 32), t32);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) acc[32 + i] = t32[i];
                            if (D == 72) {
                                uint32_t t8[8];
                                tmem_ld8(taddr + (uint32_t)(hh * D + 64), t8);
                                tmem_ld_wait();
#pragma unroll
                                for (int i = 0; i < 8; ++i) acc[(D == 72 ? 64 : 0) + i] = t8[i];
                            }
                        }
                        if (hh == HPT - 1) release_acc();
                        if (col0 >= P.N) continue;
                        float v[D];
                        float ss = 0.f;
                        // shared-memory loads in batches ahead of their uses (in-order issue: see FE_STREAM)
#pragma unroll
                        for (int i0 = 0; i0 < D; i0 += 24) {
                            constexpr int kB = 6;
                            float4 sv[kB];
#pragma unroll
                            for (int b = 0; b < kB; ++b) {
                                const int i = i0 + 4 * b;
                                if (i < D) {
                                    if (F) sv[b] = lds128(vec + (hh * D + i) * 4);
                                    else sv[b] = shrow ? __ldg(reinterpret_cast<const float4*>(shrow + col0 + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                                }
                            }
#pragma unroll
                            for (int b = 0; b < kB; ++b) {
                                const int i = i0 + 4 * b;
                                if (i < D) {
                                    v[i] = fmaf(rstd, __uint_as_float(acc[i]), sv[b].x);
                                    v[i + 1] = fmaf(rstd, __uint_as_float(acc[i + 1]), sv[b].y);
                                    v[i + 2] = fmaf(rstd, __uint_as_float(acc[i + 2]), sv[b].z);
                                    v[i + 3] = fmaf(rstd, __uint_as_float(acc[i + 3]), sv[b].w);
                                    ss = fmaf(v[i], v[i], fmaf(v[i + 1], v[i + 1], fmaf(v[i + 2], v[i + 2], fmaf(v[i + 3], v[i + 3], ss))));
                                }
                            }
                        }
                        // staging buffer hb was last used two heads ago: that store must have finished reading
                        if (lane == 0) bulk_wait_read<1>();
                        __syncwarp();
                        const uint32_t srow = wstg + hb * (32 * kRowBytes) + lane * kRowBytes;
                        const float rs = do_norm ? rsqrtf(ss / (float)D + P.eps_head) : 1.0f;
#pragma unroll
                        for (int c8 = 0; c8 < D / 8; ++c8) {
                            // pairs 4*c8 .. 4*c8+3: even pairs rotate by the x angle k = 2*c8 (+1), odd pairs by the y angle
                            const float4 w0 = lds128(wrow + c8 * 32), w1 = lds128(wrow + c8 * 32 + 16);
                            float4 cx, cy;
                            if (F) {
                                cx = lds128(txrow + c8 * 16);
                                cy = lds128(tyrow + c8 * 16);
                            } else if (do_rope) {
                                const float2 p0 = __ldg(rp + 4 * c8), p1 = __ldg(rp + 4 * c8 + 1);
                                const float2 p2 = __ldg(rp + 4 * c8 + 2), p3 = __ldg(rp + 4 * c8 + 3);
                                cx = make_float4(p0.x, p0.y, p2.x, p2.y);
                                cy = make_float4(p1.x, p1.y, p3.x, p3.y);
                            } else {
                                cx = make_float4(1.f, 0.f, 1.f, 0.f);
                                cy = cx;
                            }
                            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                            const float cc[4] = {cx.x, cy.x, cx.z, cy.z}, sn[4] = {cx.y, cy.y, cx.w, cy.w};
                            uint32_t o[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int j = c8 * 4 + e;      // pair index
                                const float a = v[2 * j] * rs * wv[2 * e], b = v[2 * j + 1] * rs * wv[2 * e + 1];
                                o[e] = pack_bf2(a * cc[e] - b * sn[e], a * sn[e] + b * cc[e]);
                            }
                            // D = 64: 128-byte rows would put a quarter-warp on one bank group; rotate the 16-byte chunks
                            // by the row (the output map uses SWIZZLE_128B for D = 64, none for D = 72)
                            const int cpos = (D == 64) ? (c8 ^ (lane & 7)) : c8;
                            sts128u(srow + cpos * 16, make_uint4(o[0], o[1], o[2], o[3]));
                        }
                        if (D == 72 && P.out_pitch == 80) sts128u(srow + 9 * 16, make_uint4(0u, 0u, 0u, 0u));   // pad columns 72..79
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&maps.o, wstg + hb * (32 * kRowBytes), (col0 / D) * P.out_pitch, row0);
                            bulk_commit();
                        }
                    }
                };
                if (fast) body(std::true_type{}); else body(std::false_type{});
                __syncwarp();
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
            if (lane == 0) bulk_wait_all();
        } else {
            // FE_NORM_SWIGLU: chunk of 32 accumulator columns = [16 x w1 | 16 x w3] -> 16 outputs = 32 bytes per row
            constexpr int NV4 = BN / 128;
            const bool fast = uni || !P.shw;
            float sq[kMaxSsqParts];
            float4 pre[NV4];
            auto fetch = [&](int tile) {
                const long long row = tile_row0(tile) + lane;
                const long long rowc = row < P.M ? row : (long long)P.M - 1;
                const int nb = (tile % num_n) * BN;
#pragma unroll
                for (int p = 0; p < kMaxSsqParts; ++p)
                    sq[p] = (P.ssq_in && p < P.ssq_parts) ? __ldg(P.ssq_in + (long long)p * P.M + rowc) : 0.f;
#pragma unroll
                for (int k = 0; k < NV4; ++k) {
                    const int ci = 4 * (lane + 32 * k);
                    pre[k] = (P.shw && uni && nb + ci < P.N)
                                 ? __ldg(reinterpret_cast<const float4*>(P.shw + (rowc / P.L) * P.shw_stride + nb + ci))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            if (tile0 < num_tiles) fetch(tile0);
            for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
                const int n_blk = tile % num_n;
                const int row0 = tile_row0(tile);
                const long long row = row0 + lane;
                const long long rowc = row < P.M ? row : (long long)P.M - 1;
                const long long img = rowc / P.L;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
                const int nbase = n_blk * BN;
                const float* shrow = P.shw ? P.shw + img * P.shw_stride : nullptr;
                float ssum = 0.f;
#pragma unroll
                for (int p = 0; p < kMaxSsqParts; ++p) ssum += sq[p];
                const float rstd = P.ssq_in ? rsqrtf(ssum * P.inv_hidden + P.eps) : 1.0f;
#pragma unroll
                for (int k = 0; k < NV4; ++k) sts128(vec + 4 * (lane + 32 * k) * 4, pre[k]);
                __syncwarp();
                if (tile + tile_stride < num_tiles) fetch(tile + tile_stride);
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
                auto body = [&](auto fast_tag) {
                    constexpr bool F = decltype(fast_tag)::value;
#pragma unroll 1
                    for (int c = 0; c < BN / 32; ++c) {
                        uint32_t acc[32];
                        tmem_ld32(taddr + (uint32_t)(c * 32), acc);
                        tmem_ld_wait();
                        if (c == BN / 32 - 1) release_acc();
                        const int n0 = nbase + c * 32;
                        if (n0 >= P.N || row >= P.M) continue;
                        float4 sv[8];               // all shift loads first, then the arithmetic (in-order issue)
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (F) sv[i] = lds128(vec + (c * 32 + 4 * i) * 4);
                            else sv[i] = shrow ? __ldg(reinterpret_cast<const float4*>(shrow + n0 + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 s4 = sv[i >> 2];
                            v[i] = fmaf(rstd, __uint_as_float(acc[i]), s4.x);
                            v[i + 1] = fmaf(rstd, __uint_as_float(acc[i + 1]), s4.y);
                            v[i + 2] = fmaf(rstd, __uint_as_float(acc[i + 2]), s4.z);
                            v[i + 3] = fmaf(rstd, __uint_as_float(acc[i + 3]), s4.w);
                        }
                        uint32_t w[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            w[i] = pack_bf2(silu_f(v[2 * i]) * v[16 + 2 * i], silu_f(v[2 * i + 1]) * v[16 + 2 * i + 1]);
                        __nv_bfloat16* o = P.out + row * P.ldo + (n0 >> 1);
                        *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4*>(o + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                };
                if (fast) body(std::true_type{}); else body(std::false_type{});
                __syncwarp();
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc_2sm(tmem_base, C::kTmemCols);
}

// ------------------------------------------------------------------ host side
static int make_map(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* ptr, long long rows, long long cols,
                    long long ld, int box_cols, int box_rows, CUtensorMapSwizzle sw) {
    PFN_encodeTiled enc = get_tensormap_encoder();
    if (!enc) { deco_set_error("cuTensorMapEncodeTiled entry point not available"); return DECO_ERR_DRIVER; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { deco_set_error("fused gemm: cuTensorMapEncodeTiled failed: %d", (int)r); return DECO_ERR_DRIVER; }
    return DECO_OK;
}

static int sm_count() { return device_sm_count() - deco_reserved_sms(); }

template <int BN, int EPI>
static int launch(const Maps& maps, const Params& P, cudaStream_t st) {
    using C = Cfg<BN, EPI>;
    auto kern = gemm_fused_kernel<BN, EPI>;
    static unsigned long long attr_done = 0;
    if (!device_setup_done(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
        if (e != cudaSuccess) { deco_set_error("fused gemm smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        mark_device_setup(attr_done);
    }
    const int tiles = ((P.M + 2 * kBM - 1) / (2 * kBM)) * ((P.N + BN - 1) / BN);
    int groups = sm_count() / 2;
    if (tiles < groups) groups = tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups * 2);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = deco_pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps, P);
    if (e != cudaSuccess) { deco_set_error("fused gemm launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return DECO_OK;
}

// K above which the ring-staged FE_STREAM variant is used (DECO_STREAM_RING_K overrides, 0 = never)
static bool stream_uses_ring(int K) {
    static int ring_k = -1;
    if (ring_k < 0) { const char* e = getenv("DECO_STREAM_RING_K"); ring_k = e ? atoi(e) : 2048; }
    return ring_k > 0 && K > ring_k;
}
// Column tile: the ring variant affords 256-wide tiles (a ragged last tile is cheaper than narrow MMAs); whole-tile
// staging fits 192.
static int stream_tile_n(int N, int K) {
    static int ring_bn = -1;      // DECO_STREAM_RING_BN = 128 / 192 / 256 overrides the ring variant's tile width (A/B measurements)
    if (ring_bn < 0) { const char* e = getenv("DECO_STREAM_RING_BN"); ring_bn = e ? atoi(e) : 0; }
    if (stream_uses_ring(K) && ring_bn > 0 && N % ring_bn == 0) return ring_bn;
    if (stream_uses_ring(K) && N >= 256) return 256;
    static int plain_bn = -1;     // DECO_STREAM_BN = 128 / 192 overrides the whole-tile-staging variant's tile width (A/B measurements)
    if (plain_bn < 0) { const char* e = getenv("DECO_STREAM_BN"); plain_bn = e ? atoi(e) : 0; }
    if (!stream_uses_ring(K) && (plain_bn == 128 || (plain_bn == 192 && N % 192 == 0))) return plain_bn;
    return (N % 192 == 0) ? 192 : 128;
}

static int check_ab(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, int n_mult = 32) {
    DECO_CHECK_ARG(A && W, "fused gemm: null operand");
    DECO_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && N % n_mult == 0,
                   "fused gemm: bad shape M=%d N=%d K=%d (K, lda, ldw %% 8 == 0; N %% %d == 0)", M, N, K, n_mult);
    DECO_CHECK_ARG((((uintptr_t)A | (uintptr_t)W) & 15) == 0, "fused gemm: operands must be 16-byte aligned");
    return DECO_OK;
}

}  // namespace fused
}  // namespace deco

using namespace deco;
using namespace deco::fused;

extern "C" int deco_gemm_stream_parts(int N, int K) { return N > 0 ? (N + stream_tile_n(N, K) - 1) / stream_tile_n(N, K) : 0; }

extern "C" int deco_gemm_stream(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
                                const float* bias, const float* resid, long long ldr, float* out, long long ldo,
                                const void* gate, long long gate_stride, int rows_per_image,
                                const float* next_norm_w, const void* next_scale, long long next_scale_stride,
                                void* xg_out, long long ldx, float* ssq_out, void* stream)
{
    int rc = check_ab(A, lda, W, ldw, M, N, K);
    if (rc) return rc;
    DECO_CHECK_ARG(out && ldo % 4 == 0 && ((uintptr_t)out & 15) == 0, "gemm_stream: bad output");
    DECO_CHECK_ARG(!resid || (ldr % 4 == 0 && ((uintptr_t)resid & 15) == 0), "gemm_stream: bad residual");
    DECO_CHECK_ARG(rows_per_image > 0, "gemm_stream: rows_per_image must be positive");
    DECO_CHECK_ARG(!gate || (gate_stride % 4 == 0 && ((uintptr_t)gate & 7) == 0), "gemm_stream: bad gate");
    DECO_CHECK_ARG(!next_norm_w || (next_scale && xg_out && next_scale_stride % 4 == 0 && ldx % 8 == 0 &&
                                    ((uintptr_t)xg_out & 15) == 0 && ((uintptr_t)next_scale & 7) == 0),
                   "gemm_stream: next-norm arguments invalid");
    const int bn = stream_tile_n(N, K);
    Maps maps;
    if ((rc = make_map(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda, kBK, kBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map(&maps.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, N, K, ldw, kBK, bn / 2, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map(&maps.o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, M, N, ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if (resid) {
        if ((rc = make_map(&maps.r, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, resid, M, N, ldr, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    } else maps.r = maps.o;
    maps.x = maps.o;
    Params P = {};
    P.M = M; P.N = N; P.K = K; P.L = rows_per_image;
    P.bias = bias; P.gate = (const __nv_bfloat16*)gate; P.gate_stride = gate_stride; P.has_resid = resid ? 1 : 0;
    P.next_w = next_norm_w; P.next_scale = (const __nv_bfloat16*)next_scale; P.next_scale_stride = next_scale_stride;
    P.ssq_out = ssq_out; P.xg = (__nv_bfloat16*)xg_out; P.ldx = ldx;
    if (stream_uses_ring(K)) {
        if (bn == 256) return launch<256, FE_STREAM_RING>(maps, P, (cudaStream_t)stream);
        if (bn == 192) return launch<192, FE_STREAM_RING>(maps, P, (cudaStream_t)stream);
        return launch<128, FE_STREAM_RING>(maps, P, (cudaStream_t)stream);
    }
    if (bn == 192) return launch<192, FE_STREAM>(maps, P, (cudaStream_t)stream);
    return launch<128, FE_STREAM>(maps, P, (cudaStream_t)stream);
}

extern "C" int deco_gemm_norm_qkv(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                                  int M, int N, int K, int rows_per_image,
                                  const float* ssq_in, int ssq_parts, int norm_hidden, float norm_eps,
                                  const float* shw, long long shw_stride,
                                  int heads, int head_dim, const float* w_seg0, const float* w_seg1, const float* w_seg2,
                                  int rope_mask, const float* rope_cos_sin, int rope_tokens_per_row, float head_eps,
                                  int out_head_pitch, void* stream)
{
    int rc = check_ab(A, lda, W, ldw, M, N, K, 8);
    if (out_head_pitch <= 0) out_head_pitch = head_dim;
    DECO_CHECK_ARG(out_head_pitch == head_dim || (head_dim == 72 && out_head_pitch == 80),
                   "gemm_norm_qkv: output head pitch %d not built (head_dim, or 80 for head_dim 72)", out_head_pitch);
    if (rc) return rc;
    DECO_CHECK_ARG(out && ldo % 8 == 0 && ((uintptr_t)out & 15) == 0, "gemm_norm_qkv: bad output");
    DECO_CHECK_ARG(head_dim == 64 || head_dim == 72, "gemm_norm_qkv: head_dim %d not built (64, 72)", head_dim);
    const int seg = heads * head_dim;
    DECO_CHECK_ARG(heads > 0 && N % seg == 0 && N / seg >= 1 && N / seg <= 3,
                   "gemm_norm_qkv: N must be 1..3 segments of heads*head_dim");
    DECO_CHECK_ARG(rows_per_image > 0 && (!ssq_in || (ssq_parts > 0 && ssq_parts <= kMaxSsqParts && norm_hidden > 0)), "gemm_norm_qkv: bad norm arguments");
    DECO_CHECK_ARG(!rope_mask || rope_cos_sin, "gemm_norm_qkv: rope table missing");
    DECO_CHECK_ARG(!shw || (shw_stride % 4 == 0 && ((uintptr_t)shw & 15) == 0), "gemm_norm_qkv: bad shift-product matrix");
    const int bn = head_dim == 72 ? 224 : 256;
    Maps maps;
    if ((rc = make_map(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda, kBK, kBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map(&maps.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, N, K, ldw, kBK, bn / 2, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map(&maps.o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, M, (long long)(N / head_dim) * out_head_pitch, ldo,
                       out_head_pitch, 32, head_dim == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
    maps.r = maps.o; maps.x = maps.o;
    Params P = {};
    P.M = M; P.N = N; P.K = K; P.L = rows_per_image;
    P.ssq_in = ssq_in; P.ssq_parts = ssq_parts; P.inv_hidden = norm_hidden > 0 ? 1.0f / (float)norm_hidden : 0.f; P.eps = norm_eps;
    P.shw = shw; P.shw_stride = shw_stride;
    P.seg_cols = seg;
    P.out_pitch = out_head_pitch;
    P.seg_w[0] = w_seg0; P.seg_w[1] = w_seg1; P.seg_w[2] = w_seg2;
    for (int i = 0; i < 3; ++i) P.seg_rope[i] = (rope_mask >> i) & 1;
    P.rope = (const float2*)rope_cos_sin; P.rope_wp = rope_tokens_per_row; P.eps_head = head_eps;
    if (head_dim == 72) return launch<224, FE_NORM_QKV>(maps, P, (cudaStream_t)stream);
    return launch<256, FE_NORM_QKV>(maps, P, (cudaStream_t)stream);
}

extern "C" int deco_gemm_norm_swiglu(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                                     int M, int N, int K, int rows_per_image,
                                     const float* ssq_in, int ssq_parts, int norm_hidden, float norm_eps,
                                     const float* shw, long long shw_stride, void* stream)
{
    int rc = check_ab(A, lda, W, ldw, M, N, K);
    if (rc) return rc;
    DECO_CHECK_ARG(out && ldo % 8 == 0 && ((uintptr_t)out & 15) == 0, "gemm_norm_swiglu: bad output");
    DECO_CHECK_ARG(rows_per_image > 0 && (!ssq_in || (ssq_parts > 0 && ssq_parts <= kMaxSsqParts && norm_hidden > 0)), "gemm_norm_swiglu: bad norm arguments");
    DECO_CHECK_ARG(!shw || (shw_stride % 4 == 0 && ((uintptr_t)shw & 15) == 0), "gemm_norm_swiglu: bad shift-product matrix");
    Maps maps;
    if ((rc = make_map(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda, kBK, kBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_map(&maps.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, N, K, ldw, kBK, 128, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    maps.r = maps.a; maps.o = maps.a; maps.x = maps.a;
    Params P = {};
    P.M = M; P.N = N; P.K = K; P.L = rows_per_image;
    P.ssq_in = ssq_in; P.ssq_parts = ssq_parts; P.inv_hidden = norm_hidden > 0 ? 1.0f / (float)norm_hidden : 0.f; P.eps = norm_eps;
    P.shw = shw; P.shw_stride = shw_stride;
    P.out = (__nv_bfloat16*)out; P.ldo = ldo;
    return launch<256, FE_NORM_SWIGLU>(maps, P, (cudaStream_t)stream);
}
