// Backward of non-causal softmax attention on the 5th-generation tensor cores (training step; autograd through
// F.scaled_dot_product_attention at /root/reference/src/models/transformer/dit_c2i_DeCo.py:181-185).
//
// Two deterministic passes as in csrc/attention_bwd.cu (no atomics; S and dP are recomputed in each), every product a
// tcgen05.mma with M = 128 and the accumulators in tensor memory; two threads share a score row (warps w and w + 4
// address the same 32 TMEM lanes and take 32 of a block's 64 columns each):
//   pass 0  item = (image, head, 128 queries), one step per 64-key block:  S = Q K^T, dP = dO V^T  (SS, N = 64)
//           dS / scale = exp2(c S - lse2) (dP - delta) -> bf16 -> tensor memory;  dQ += dS K  (TS, B = K block MN-major)
//   pass 1  item = (image, head, 128 keys), one step per 64-query block, on the TRANSPOSED scores (rows = keys):
//           S^T = K Q^T, dP^T = V dO^T;  P^T -> bf16 -> tensor memory, dV += P^T dO (TS);
//           dS^T / scale -> bf16 -> shared memory (K-major SW32 tile), dK += dS^T Q (SS)
//   the softmax scale is applied once per output element when dQ / dK leave tensor memory.
// delta = rowsum(dO . O) comes from attn_bwd_delta_kernel; lse2 from the forward (deco_attention_fwd_lse).
// Q, K, V, dO tiles come straight from the strided [tokens, 3H] matrices through 4-D TMA maps (16-column x 64-row boxes,
// 32-byte swizzle = the operand layout; head dim 72 zero-padded to 80, ragged sequence tails zero-filled) -- one tile in
// shared memory serves as K-major operand of the score products and as MN-major operand of the gradient products.
// How the steps are pipelined is described above attn_bwd_pipe_kernel.  History of this file (profiles/attn_bwd_bench_r2.txt):
// mma.sync pair 219 us -> unpipelined tcgen05 pair (load -> MMA -> row math -> MMA in sequence) 143 us -> this kernel 95 us
// per XL/16 layer of 32 images.  A cp.async gather in place of the TMA boxes measured slower (109 us): once the loads have
// four issuing warps the row-math warps are the critical path (scripts/abt_trace.py), not the TMA unit.
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace deco {
namespace abt {

constexpr int kRows = 128;                      // rows of an item: one TMEM lane each

struct Params {
    const __nv_bfloat16 *o, *dout;
    __nv_bfloat16 *dq, *dk, *dv;
    const float* lse2;
    float* delta;
    long long o_stride, do_stride, dq_stride, dkv_stride;
    int B, heads, Lq, Lk;
    float scale, scale_log2;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------------------------------------ pipelined kernels
// The row math of step s overlaps the gradient products of step s - 1 and the score products of step s + 1 / s + 2:
//   * 64-column score blocks, two S / dP accumulator pairs in tensor memory (2 x 128 columns);
//   * both passes keep the long dimension of the gradient products on the TMEM lanes: the dK / dV pass works on the
//     TRANSPOSED block S^T = K Q^T (rows = keys), so P^T and dS^T are [keys x queries]; its per-query constants
//     (lse2, delta) vary along the columns and are staged in shared memory next to the column block;
//   * the bf16 operands of the gradient products are NOT written over the scores: pass 0 puts dS in its own TMEM
//     columns (TS form), pass 1 puts P^T in TMEM (TS form, dV) and dS^T in shared memory (K-major SW32 tile, SS form,
//     dK) -- 400 / 480 of the 512 columns.  So the scores of step s + 2 only wait for the row math of step s to have
//     READ buffer s % 2, not for its gradient products;
//   * one thread issues a tcgen05.mma every ~60-90 clocks whatever its shape (profiles/tmem_bench_r2.txt), and a step is
//     14 (pass 0) / 18 (pass 1) of them against ~800 clocks of row math: three issuing warps share them
//     (S = R1.C1^T | dP = R2.C2^T | gradient products), each with its own commits;
//   * warps 0-7 = row math (two threads per row, 32 columns each), warp 8 = TMA producer (row tiles double-buffered per
//     item, column blocks in a ring), warps 9-11 = the issuers, warps 12-14 = three more producers (one per operand
//     stream: a warp gets a TMA load out only every ~110 clocks).  The read-out of an item's accumulators is deferred
//     until the row math of the next item's first step is done, so the issuers never wait for it.
//   pass 0 (dQ):     rows = 128 queries, R1 = Q, R2 = dO; column blocks C1 = K, C2 = V;  dQ += dS . K
//   pass 1 (dK, dV): rows = 128 keys,    R1 = K, R2 = V;  column blocks C1 = Q, C2 = dO; dV += P^T . dO, dK += dS^T . Q
// delta = rowsum(dO . O) comes from attn_bwd_delta_kernel (one thread per (token, head)).
namespace pipe {

#ifdef ABT_TRACE
__device__ unsigned long long g_trace[2][16];
__device__ long long g_tl[2][6][64][4];      // [pass][role][step][event]: CTA 0's clock at the pipeline events
#define TL(role, step, ev) do { if (blockIdx.x == 0 && (step) < 64) g_tl[PASS][role][step][ev] = clock64(); } while (0)
#define TR_T(var) const long long var = clock64()
#define TR_ADD(slot, t0) atomicAdd(&g_trace[PASS][slot], (unsigned long long)(clock64() - (t0)))
#else
#define TR_T(var)
#define TR_ADD(slot, t0)
#define TL(role, step, ev)
#endif

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

constexpr int kCols = 64;                       // score-block width
constexpr int kMathWarps = 8;
constexpr int kThreadsP = (kMathWarps + 7) * 32;      // + 4 producers + 3 issuers (16 warps' worth of registers either way)
constexpr uint32_t kBufCols = 128;              // S (64) | dP (64) of one buffer
constexpr uint32_t kColA = 256;                 // + 32 b: bf16 TS operand of buffer b (dS in pass 0, P^T in pass 1)
constexpr uint32_t kColAcc0 = 320, kColAcc1 = 400;

template <int D, int PASS> struct PCfg {
    static constexpr int DP = (D + 15) / 16 * 16;
    static constexpr int NCH = DP / 16;
    static constexpr uint32_t RCHUNK = 128 * 32, RTILE = NCH * RCHUNK;      // row tile [128 x DP], chunk-major SW32
    static constexpr uint32_t CCHUNK = kCols * 32, CTILE = NCH * CCHUNK;    // column block [64 x DP]
    static constexpr uint32_t VEC = kCols * 8;                              // (lse2, delta) of the block's 64 queries
    static constexpr uint32_t ATILE = PASS == 1 ? 128 * kCols * 2 : 0;      // dS^T [128 keys x 64 queries] bf16 (pass 1)
    static constexpr int kStages = PASS == 0 ? 6 : (D > 64 ? 4 : 6);        // column-block ring
    static constexpr uint32_t OUT = 128 * D * 2;                            // [128 rows][D] bf16: accumulator rows on their way out
    static constexpr uint32_t OFF_A = 4 * RTILE, OFF_C = OFF_A + 2 * ATILE, OFF_VEC = OFF_C + kStages * 2 * CTILE,
                              OFF_OUT = OFF_VEC + kStages * VEC, OFF_BAR = OFF_OUT + OUT;
    static constexpr uint32_t SMEM = OFF_BAR + 256 + 1024;
    static_assert(SMEM <= 227 * 1024, "attention backward: shared memory budget");
};

struct PMaps { CUtensorMap r1, r2, c1, c2; };   // all with 64-row boxes

template <int NCH>
__device__ __forceinline__ void load_rows64(uint32_t dst, uint32_t chunk_bytes, const CUtensorMap* map, uint32_t bar, int row, int h, int b) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) tma_load_4d(dst + c * chunk_bytes, map, bar, 16 * c, row, h, b);
}

__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta,
                                      long long o_stride, long long do_stride, int B, int heads, int Lq, int D)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (token, head), head fastest
    if (idx >= (long long)B * Lq * heads) return;
    const int h = (int)(idx % heads);
    const long long tok = idx / heads;
    const uint4* orow = reinterpret_cast<const uint4*>(o + tok * o_stride + (long long)h * D);
    const uint4* drow = reinterpret_cast<const uint4*>(dout + tok * do_stride + (long long)h * D);
    float acc = 0.f;
    for (int c = 0; c < D / 8; ++c) {
        const uint4 a = __ldg(orow + c), g = __ldg(drow + c);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf2(aw[e]), y = unpack_bf2(gw[e]);
            acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
        }
    }
    const long long b = tok / Lq, row = tok - b * Lq;
    delta[(b * heads + h) * Lq + row] = acc;
}

template <int D, int PASS>
__global__ void __launch_bounds__(kThreadsP, 1) attn_bwd_pipe_kernel(const __grid_constant__ PMaps M, const Params P)
{
    using C = PCfg<D, PASS>;
    constexpr int NS = C::kStages;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    auto sR1 = [&](int n) { return base + (uint32_t)(n & 1) * 2 * C::RTILE; };
    auto sR2 = [&](int n) { return sR1(n) + C::RTILE; };
    auto sA = [&](int b) { return base + C::OFF_A + (uint32_t)b * C::ATILE; };
    auto sC1 = [&](int s) { return base + C::OFF_C + (uint32_t)s * 2 * C::CTILE; };
    auto sC2 = [&](int s) { return sC1(s) + C::CTILE; };
    auto sVec = [&](int s) { return base + C::OFF_VEC + (uint32_t)s * C::VEC; };
    const uint32_t bars = base + C::OFF_BAR;
    auto row_full = [&](int i) { return bars + 8u * i; };            // [2]
    auto row_empty = [&](int i) { return bars + 16 + 8u * i; };      // [2]  both score issuers
    auto s_full = [&](int b) { return bars + 32 + 8u * b; };         // [2]  both score issuers
    auto math_done = [&](int b) { return bars + 48 + 8u * b; };      // [2]  the 8 row-math warps: S / dP read, operands written
    auto g_done = [&](int b) { return bars + 64 + 8u * b; };         // [2]  gradient products of the buffer done: operands free
    const uint32_t acc_full = bars + 80, acc_empty = bars + 88, slot = bars + 96;
    auto ld_full = [&](int s) { return bars + 104 + 8u * s; };       // [NS]
    auto ld_empty = [&](int s) { return bars + 104 + 8u * NS + 8u * s; };   // [NS] all three issuers
    static_assert(104 + 16 * NS <= 256, "barrier area");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(row_full(i), 2); mbar_init(row_empty(i), 2); mbar_init(s_full(i), 2);
            mbar_init(math_done(i), kMathWarps); mbar_init(g_done(i), 1);
        }
        for (int s = 0; s < NS; ++s) { mbar_init(ld_full(s), 2); mbar_init(ld_empty(s), 3); }
        mbar_init(acc_full, 1); mbar_init(acc_empty, kMathWarps);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    // rows = queries (pass 0) or keys (pass 1); columns = the other one
    const int Lr = PASS == 0 ? P.Lq : P.Lk, Lc = PASS == 0 ? P.Lk : P.Lq;
    const int nr = (Lr + kRows - 1) / kRows, spi = (Lc + kCols - 1) / kCols;      // row tiles per (image, head), steps per item
    const int nitems = P.B * P.heads * nr;
    const int nlocal = (int)blockIdx.x < nitems ? (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nsteps = nlocal * spi;
    auto decode = [&](int n, int& ri, int& h, int& b) {
        const int item = blockIdx.x + n * gridDim.x;
        ri = item % nr; h = (item / nr) % P.heads; b = item / (nr * P.heads);
    };

    if (warp == kMathWarps || warp == kMathWarps + 4) {
        // ===================== column-block producers: pw = 0 loads C1 (+ the per-query constants of pass 1), pw = 1 loads C2
        // One thread gets a TMA load out only every ~250 clocks and one warp every ~110 (scripts/tma_bench.cu), so the
        // 16-column boxes of a tile are issued by different lanes (lane = box) once lane 0 has posted the byte count, and
        // the four operand streams (C1, C2, R1, R2) have a warp each.
        const int pw = warp == kMathWarps ? 0 : 1;
        float vl[2] = {0.f, 0.f}, vd[2] = {0.f, 0.f};
        auto fetch_vec = [&](int h, int b, int j) {     // pass 1: (lse2, delta) of the 64 queries of a step, two per lane
            const long long sb = ((long long)b * P.heads + h) * P.Lq;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int qrow = j * kCols + lane + 32 * e;
                vl[e] = __int_as_float(0x7f800000); vd[e] = 0.f;       // a query past the end: P = exp2(-inf) = 0
                if (qrow < P.Lq) { vl[e] = __ldg(P.lse2 + sb + qrow); vd[e] = P.delta[sb + qrow]; }
            }
        };
        TR_T(pall);
        int s = 0;
        uint32_t eph = 1;                         // parity to wait for on ld_empty (first round: free)
        if (nlocal > 0 && PASS == 1 && pw == 0) { int ri, h, b; decode(0, ri, h, b); fetch_vec(h, b, 0); }
        for (int n = 0; n < nlocal; ++n) {
            int ri, h, b, ri1 = 0, h1 = 0, b1 = 0;
            decode(n, ri, h, b);
            if (n + 1 < nlocal) decode(n + 1, ri1, h1, b1);
            for (int j = 0; j < spi; ++j) {
                TR_T(p1);
                if (lane == 0 && pw == 0) TL(4, n * spi + j, 0);
                mbar_wait(ld_empty(s), eph);
                if (lane == 0 && pw == 0) TL(4, n * spi + j, 1);
                if (lane == 0 && pw == 0) { TR_ADD(12, p1); }
                if (lane == 0) {
                    if (PASS == 1 && pw == 0) mbar_expect_tx_only(ld_full(s), C::CTILE);     // the arrival follows the vector stores
                    else mbar_expect_tx(ld_full(s), C::CTILE);
                }
                __syncwarp();
                TR_T(p2);
                if (lane < C::NCH)                // lane = chunk
                    tma_load_4d((pw ? sC2(s) : sC1(s)) + lane * C::CCHUNK, pw ? &M.c2 : &M.c1, ld_full(s), 16 * lane, j * kCols, h, b);
                __syncwarp();
                if (lane == 0 && pw == 0) TL(4, n * spi + j, 2);
                if (lane == 0 && pw == 0) { TR_ADD(13, p2); }
                if (PASS == 1 && pw == 0) {
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(sVec(s) + (uint32_t)(lane + 32 * e) * 8), "f"(vl[e]), "f"(vd[e]) : "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ld_full(s));     // (release: orders the warp's vector stores before it)
                    // the next step's constants: in flight while this warp waits for the next free stage
                    if (j + 1 < spi) fetch_vec(h, b, j + 1);
                    else if (n + 1 < nlocal) fetch_vec(h1, b1, 0);
                }
                if (++s == NS) { s = 0; eph ^= 1; }
            }
        }
        if (lane == 0 && pw == 0) { TR_ADD(14, pall); }
    } else if (warp == kMathWarps + 5 || warp == kMathWarps + 6) {
        // ===================== row-tile producers: pw = 0 loads R1, pw = 1 loads R2 (each item's pair as soon as its buffer is free)
        const int pw = warp - (kMathWarps + 5);
        for (int n = 0; n < nlocal; ++n) {
            int ri, h, b;
            decode(n, ri, h, b);
            mbar_wait(row_empty(n & 1), (uint32_t)(((n >> 1) & 1) ^ 1));
            if (lane == 0) mbar_expect_tx(row_full(n & 1), C::RTILE);
            __syncwarp();
            if (lane < 2 * C::NCH) {              // lane = (64-row half, chunk)
                const int hf = lane / C::NCH, c = lane % C::NCH;
                tma_load_4d((pw ? sR2(n) : sR1(n)) + hf * C::CCHUNK + c * C::RCHUNK, pw ? &M.r2 : &M.r1, row_full(n & 1), 16 * c,
                            ri * kRows + hf * kCols, h, b);
            }
        }
    } else if (warp == kMathWarps + 1 || warp == kMathWarps + 2) {
        // ===================== score issuers: which = 0: S = R1 . C1^T, which = 1: dP = R2 . C2^T =====================
        if (lane == 0) {
            const int which = warp - (kMathWarps + 1);
            constexpr uint32_t idesc_s = make_idesc_major(128, kCols, 0, 0);
            for (int st = 0; st < nsteps; ++st) {
                const int n = st / spi, j = st - n * spi, s = st % NS, b = st & 1;
                TR_T(t0);
                if (j == 0) mbar_wait(row_full(n & 1), (uint32_t)((n >> 1) & 1));
                if (which == 0) { TR_ADD(10, t0); }
                TL(which, st, 0);
                mbar_wait(ld_full(s), (uint32_t)((st / NS) & 1));
                TL(which, st, 1);
                if (which == 0) { TR_ADD(3, t0); }
                TR_T(t1);
                mbar_wait(math_done(b), (uint32_t)(((st >> 1) & 1) ^ 1));      // step st - 2 has read this buffer (free at first)
                if (which == 0) { TR_ADD(0, t1); }
                TL(which, st, 2);
                TR_T(t2);
                tc_fence_after();
                const uint64_t da = make_umma_desc(which ? sR2(n) : sR1(n), 16, 256, 6);
                const uint64_t db = make_umma_desc(which ? sC2(s) : sC1(s), 16, 256, 6);
                const uint32_t d0 = tmem + (uint32_t)b * kBufCols + (uint32_t)which * kCols;
#pragma unroll
                for (int kc = 0; kc < C::NCH; ++kc)
                    umma_bf16(d0, da + (uint64_t)kc * (C::RCHUNK >> 4), db + (uint64_t)kc * (C::CCHUNK >> 4), idesc_s, kc ? 1u : 0u);
                umma_commit(s_full(b));
                umma_commit(ld_empty(s));
                if (j == spi - 1) umma_commit(row_empty(n & 1));
                TL(which, st, 3);
                if (which == 0) { TR_ADD(4, t2); }
            }
        }
    } else if (warp == kMathWarps + 3) {
        // ===================== gradient-product issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_g = make_idesc_major(128, C::DP, 0, 1);      // A K-major (TMEM or shared), B MN-major
            for (int st = 0; st < nsteps; ++st) {
                const int n = st / spi, j = st - n * spi, s = st % NS, b = st & 1;
                TR_T(t0);
                TL(2, st, 0);
                mbar_wait(math_done(b), (uint32_t)((st >> 1) & 1));
                TL(2, st, 1);
                TR_ADD(1, t0);
                if (j == 0 && n > 0) mbar_wait(acc_empty, (uint32_t)((n - 1) & 1));   // the previous item's accumulators were read out
                TR_T(t2);
                tc_fence_after();
                const uint32_t ta = tmem + kColA + (uint32_t)b * 32;
                // the column block read MN-major: [64 (K) x DP (N)], 16 K-rows = 512 bytes
                const uint64_t m1 = make_umma_desc(sC1(s), C::CCHUNK, 256, 6), m2 = make_umma_desc(sC2(s), C::CCHUNK, 256, 6);
                const uint64_t ads = make_umma_desc(sA(b), 16, 256, 6);           // pass 1: dS^T tile, K-major, one 16-query chunk per K-step
#pragma unroll
                for (int ks = 0; ks < kCols / 16; ++ks) {
                    const uint32_t accf = (j > 0 || ks > 0) ? 1u : 0u;
                    if (PASS == 0) {
                        umma_bf16_ts(tmem + kColAcc0, ta + (uint32_t)(ks * 8), m1 + (uint64_t)ks * (512 >> 4), idesc_g, accf);      // dQ += dS . K
                    } else {
                        umma_bf16_ts(tmem + kColAcc0, ta + (uint32_t)(ks * 8), m2 + (uint64_t)ks * (512 >> 4), idesc_g, accf);      // dV += P^T . dO
                        umma_bf16(tmem + kColAcc1, ads + (uint64_t)ks * (4096 >> 4), m1 + (uint64_t)ks * (512 >> 4), idesc_g, accf); // dK += dS^T . Q
                    }
                }
                umma_commit(g_done(b));
                umma_commit(ld_empty(s));
                if (j == spi - 1) umma_commit(acc_full);
                TL(2, st, 3);
                TR_ADD(2, t2);
            }
        }
    } else {
        // ===================== row math: thread = (row, 32 of the block's 64 columns) =====================
        const int r = tid & 127, half = tid >> 7;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        float lse_n = 0.f, delta_n = 0.f;
        auto fetch_row_consts = [&](int n) {      // pass 0: the query's lse2 and delta
            if (PASS != 0 || n >= nlocal) return;
            int ri, h, b;
            decode(n, ri, h, b);
            const int row = ri * kRows + r;
            if (row < P.Lq) {
                const long long si = ((long long)b * P.heads + h) * P.Lq + row;
                lse_n = __ldg(P.lse2 + si);
                delta_n = P.delta[si];
            }
        };
        // accumulators of item n -> global.  A thread holds 32 / 40 columns of its row; stored straight from there a warp
        // store touches 32 rows (32 L1 transactions, ~2000 clocks per accumulator), so the two warps of a 32-row group
        // transpose through a [32 rows][D] scratch and write 16-byte pieces in row-major order instead
        auto read_out = [&](int n) {
            int ri, h, b;
            decode(n, ri, h, b);
            TR_T(e0);
            mbar_wait(acc_full, (uint32_t)(n & 1));
            if (tid == 0) { TR_ADD(7, e0); }
            TR_T(e1);
            if (tid == 0) TL(5, n, 0);
            tc_fence_after();
            constexpr int kSplit = (D / 8 + 1) / 2 * 8;               // 40 | 32 columns of 72 for the halves, 32 | 32 of 64
            constexpr int kPieces = 32 * (D / 8);                     // 16-byte pieces of the group's 32 rows
            const int grp = warp & 3, t64 = half * 32 + lane;         // the group's 64 threads
            const uint32_t scr = base + C::OFF_OUT + (uint32_t)grp * 32 * D * 2;
            const int row0 = ri * kRows + grp * 32;                   // first row of the group
#pragma unroll
            for (int acc = 0; acc < (PASS == 0 ? 1 : 2); ++acc) {
                const uint32_t tacc = trow + (acc == 0 ? kColAcc0 : kColAcc1) + (uint32_t)(half * kSplit);
                uint32_t v[32], w[8];
                tmem_ld32(tacc, v);
                if (kSplit > 32) tmem_ld8(tacc + 32, w);              // (half 1 reads the zero padding of columns 72..79 and drops it)
                tmem_ld_wait();
                if (tid == 0 && acc == 0) TL(5, n, 1);
                const float sc = (PASS == 0 || acc == 1) ? P.scale : 1.0f;      // dQ and dK carry the softmax scale (dV does not)
                const uint32_t srow = scr + (uint32_t)lane * D * 2 + (uint32_t)(half * kSplit) * 2;
#pragma unroll
                for (int c = 0; c < 32; c += 8)
                    sts128u(srow + c * 2, pack_bf2(sc * __uint_as_float(v[c]), sc * __uint_as_float(v[c + 1])),
                            pack_bf2(sc * __uint_as_float(v[c + 2]), sc * __uint_as_float(v[c + 3])),
                            pack_bf2(sc * __uint_as_float(v[c + 4]), sc * __uint_as_float(v[c + 5])),
                            pack_bf2(sc * __uint_as_float(v[c + 6]), sc * __uint_as_float(v[c + 7])));
                if (kSplit > 32 && half == 0)
                    sts128u(srow + 64, pack_bf2(sc * __uint_as_float(w[0]), sc * __uint_as_float(w[1])),
                            pack_bf2(sc * __uint_as_float(w[2]), sc * __uint_as_float(w[3])),
                            pack_bf2(sc * __uint_as_float(w[4]), sc * __uint_as_float(w[5])),
                            pack_bf2(sc * __uint_as_float(w[6]), sc * __uint_as_float(w[7])));
                asm volatile("bar.sync %0, 64;" :: "r"(grp + 1) : "memory");
                if (tid == 0 && acc == 0) TL(5, n, 2);
                __nv_bfloat16* ob = (PASS == 0 ? P.dq : (acc == 0 ? P.dv : P.dk));
                const long long ostr = PASS == 0 ? P.dq_stride : P.dkv_stride;
#pragma unroll
                for (int q = t64; q < kPieces; q += 64) {
                    const int rr = q / (D / 8), cc = q - rr * (D / 8);
                    if (row0 + rr < Lr) {
                        const uint4 val = *reinterpret_cast<const uint4*>(gen + (scr - base) + (uint32_t)q * 16);
                        *reinterpret_cast<uint4*>(ob + ((long long)b * Lr + row0 + rr) * ostr + (long long)h * D + cc * 8) = val;
                    }
                }
                if (PASS == 1 && acc == 0) asm volatile("bar.sync %0, 64;" :: "r"(grp + 1) : "memory");   // scratch is reused for dK
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            if (tid == 0) TL(5, n, 3);
            if (tid == 0) { TR_ADD(9, e1); }
        };
        fetch_row_consts(0);
        TR_T(tall);
        for (int n = 0; n < nlocal; ++n) {
            int ri, h, b;
            decode(n, ri, h, b);
            const int row = ri * kRows + r;
            const bool live = row < Lr;
            const bool tile_full = (ri + 1) * kRows <= Lr;
            const float lse_off = live ? 0.f : -__int_as_float(0x7f800000);
            const float lse = lse_n, delta = delta_n;
            fetch_row_consts(n + 1);               // one item ahead: the latency hides behind this item's steps
            for (int j = 0; j < spi; ++j) {
                const int st = n * spi + j, bb = st & 1;
                const uint32_t tS = trow + (uint32_t)bb * kBufCols + (uint32_t)(32 * half);
                const uint32_t tA = trow + kColA + (uint32_t)bb * 32 + (uint32_t)(16 * half);
                TR_T(m0);
                if (tid == 0) TL(3, st, 0);
                mbar_wait(s_full(bb), (uint32_t)((st >> 1) & 1));
                if (tid == 0) TL(3, st, 1);
                if (tid == 0) { TR_ADD(5, m0); }
                TR_T(m1);
                tc_fence_after();
                uint32_t s[32], dp[32];
                tmem_ld32(tS, s);
                tmem_ld32(tS + kCols, dp);
                tmem_ld_wait();
                if (PASS == 0) {
                    // dS / scale = P (dP - delta): the softmax scale is applied once per output element at the read-out
                    const int cvalid = Lc - j * kCols - 32 * half;        // columns of this thread that exist
                    uint32_t pk[16];
                    if (tile_full && cvalid >= 32) {                       // (uniform: no per-element masks)
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float p0 = ex2f(fmaf(__uint_as_float(s[2 * i]), P.scale_log2, -lse));
                            const float p1 = ex2f(fmaf(__uint_as_float(s[2 * i + 1]), P.scale_log2, -lse));
                            pk[i] = pack_bf2(p0 * (__uint_as_float(dp[2 * i]) - delta), p1 * (__uint_as_float(dp[2 * i + 1]) - delta));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float v[2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const float p = (live && 2 * i + e < cvalid) ? ex2f(fmaf(__uint_as_float(s[2 * i + e]), P.scale_log2, -lse)) : 0.f;
                                v[e] = p * (__uint_as_float(dp[2 * i + e]) - delta);
                            }
                            pk[i] = pack_bf2(v[0], v[1]);
                        }
                    }
                    mbar_wait(g_done(bb), (uint32_t)(((st >> 1) & 1) ^ 1));   // the products of step st - 2 have read these columns
                    tc_fence_after();
                    tmem_st16(tA, pk);
                } else {
                    const uint32_t vp = sVec(st % NS) + (uint32_t)(32 * half) * 8;
                    uint32_t pp[16], pd[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float4 c2 = lds128(vp + (uint32_t)i * 16);          // (lse2, delta) of columns 2i and 2i + 1
                        // (a key row past the end: lse_off = -inf sends its P to 0; dS^T / scale here, scale at the read-out)
                        const float p0 = ex2f(fmaf(__uint_as_float(s[2 * i]), P.scale_log2, lse_off - c2.x));
                        const float p1 = ex2f(fmaf(__uint_as_float(s[2 * i + 1]), P.scale_log2, lse_off - c2.z));
                        pp[i] = pack_bf2(p0, p1);
                        pd[i] = pack_bf2(p0 * (__uint_as_float(dp[2 * i]) - c2.y), p1 * (__uint_as_float(dp[2 * i + 1]) - c2.w));
                    }
                    mbar_wait(g_done(bb), (uint32_t)(((st >> 1) & 1) ^ 1));
                    tc_fence_after();
                    tmem_st16(tA, pp);
                    // dS^T: element (row = key r, col = query) of the [128 x 64] K-major tile, 16-query chunks of 32 bytes per row
                    const uint32_t ab = sA(bb);
#pragma unroll
                    for (int c16 = 0; c16 < 2; ++c16) {
                        const int col = 32 * half + 16 * c16;
                        sts128u(ab + sw32_offset(r, col, kRows), pd[8 * c16], pd[8 * c16 + 1], pd[8 * c16 + 2], pd[8 * c16 + 3]);
                        sts128u(ab + sw32_offset(r, col + 8, kRows), pd[8 * c16 + 4], pd[8 * c16 + 5], pd[8 * c16 + 6], pd[8 * c16 + 7]);
                    }
                    fence_proxy_async();           // generic-proxy stores -> tensor-core (async proxy) reads
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(math_done(bb));
                if (tid == 0) TL(3, st, 2);
                if (tid == 0) { TR_ADD(6, m1); }
                if (j == 0 && n > 0) read_out(n - 1);      // deferred: the issuers already have this item's first steps
            }
        }
        if (nlocal > 0) read_out(nlocal - 1);
        if (tid == 0) { TR_ADD(8, tall); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace pipe

// 4-D view (d, token, head, batch) of a strided [B*L, row_stride] bf16 matrix whose columns are [head][d]
static int make_tmap(CUtensorMap* map, const void* ptr, int D, long long L, int heads, int B, long long row_stride, int box_rows) {
    PFN_encodeTiled enc = get_tensormap_encoder();
    if (!enc) { deco_set_error("cuTensorMapEncodeTiled entry point not available"); return DECO_ERR_DRIVER; }
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)heads, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)row_stride * 2, (cuuint64_t)D * 2, (cuuint64_t)L * (cuuint64_t)row_stride * 2};
    cuuint32_t box[4] = {16, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { deco_set_error("attention_bwd_tc: cuTensorMapEncodeTiled failed: %d", (int)rc); return DECO_ERR_DRIVER; }
    return DECO_OK;
}

template <int D>
static int launch_pipe(const void* q, const void* k, const void* v, const void* dout, long long q_stride, long long kv_stride,
                       long long do_stride, const Params& P, cudaStream_t st) {
    using namespace pipe;
    static unsigned long long attr_done = 0;
    if (!device_setup_done(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_pipe_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCfg<D, 0>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_pipe_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PCfg<D, 1>::SMEM);
        if (e != cudaSuccess) { deco_set_error("attention_bwd_tc attr: %s", cudaGetErrorString(e)); return (int)e; }
        mark_device_setup(attr_done);
    }
    CUtensorMap mq, mk, mv, md;
    int rc;
    if ((rc = make_tmap(&mq, q, D, P.Lq, P.heads, P.B, q_stride, kCols))) return rc;
    if ((rc = make_tmap(&mk, k, D, P.Lk, P.heads, P.B, kv_stride, kCols))) return rc;
    if ((rc = make_tmap(&mv, v, D, P.Lk, P.heads, P.B, kv_stride, kCols))) return rc;
    if ((rc = make_tmap(&md, dout, D, P.Lq, P.heads, P.B, do_stride, kCols))) return rc;
    const long long nd = (long long)P.B * P.Lq * P.heads;
    attn_bwd_delta_kernel<<<(unsigned)((nd + 255) / 256), 256, 0, st>>>(P.o, P.dout, P.delta, P.o_stride, P.do_stride, P.B, P.heads, P.Lq, D);
    DECO_CHECK_LAUNCH("attn_bwd_delta_kernel");
    const int sms = device_sm_count();
    const int items_a = P.B * P.heads * ((P.Lq + kRows - 1) / kRows), items_b = P.B * P.heads * ((P.Lk + kRows - 1) / kRows);
    PMaps A = {mq, md, mk, mv}, Bm = {mk, mv, mq, md};
    attn_bwd_pipe_kernel<D, 0><<<items_a < sms ? items_a : sms, kThreadsP, PCfg<D, 0>::SMEM, st>>>(A, P);
    DECO_CHECK_LAUNCH("attn_bwd_pipe_kernel<dQ>");
    attn_bwd_pipe_kernel<D, 1><<<items_b < sms ? items_b : sms, kThreadsP, PCfg<D, 1>::SMEM, st>>>(Bm, P);
    DECO_CHECK_LAUNCH("attn_bwd_pipe_kernel<dKdV>");
    return DECO_OK;
}

}  // namespace abt
}  // namespace deco

#ifdef ABT_TRACE
extern "C" int deco_abt_timeline_read(long long* host) {
    cudaMemcpyFromSymbol(host, deco::abt::pipe::g_tl, sizeof(long long) * 2 * 6 * 64 * 4);
    return 0;
}
extern "C" int deco_abt_trace_read(unsigned long long* host32, int reset) {
    cudaMemcpyFromSymbol(host32, deco::abt::pipe::g_trace, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(deco::abt::pipe::g_trace, z, sizeof(z)); }
    return 0;
}
#endif

// tcgen05 form of deco_attention_bwd; needs the forward's softmax statistics (lse2 from deco_attention_fwd_lse).
// delta_ws [B * heads * Lq] fp32 is filled by the first pass and read by the second.
extern "C" int deco_attention_bwd_tc(const void* q, long long q_stride, const void* k, const void* v, long long kv_stride,
                                     const void* o, long long o_stride, const void* dout, long long do_stride,
                                     void* dq, long long dq_stride, void* dk, void* dv, long long dkv_stride,
                                     const float* lse2, float* delta_ws, int B, int heads, int Lq, int Lk, int head_dim,
                                     float scale, void* stream)
{
    using namespace deco;
    using namespace deco::abt;
    DECO_CHECK_ARG(q && k && v && o && dout && dq && dk && dv && lse2 && delta_ws, "attention_bwd_tc: null pointer");
    DECO_CHECK_ARG(B > 0 && heads > 0 && Lq > 0 && Lk > 0, "attention_bwd_tc: bad shape");
    DECO_CHECK_ARG(head_dim == 64 || head_dim == 72, "attention_bwd_tc: head_dim %d not built (64, 72)", head_dim);
    DECO_CHECK_ARG(q_stride % 8 == 0 && kv_stride % 8 == 0 && o_stride % 8 == 0 && do_stride % 8 == 0 && dq_stride % 8 == 0 &&
                   dkv_stride % 8 == 0, "attention_bwd_tc: strides must be multiples of 8 elements");
    DECO_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk |
                     (uintptr_t)dv) & 15) == 0, "attention_bwd_tc: pointers must be 16-byte aligned");
    Params P;
    P.o = (const __nv_bfloat16*)o; P.dout = (const __nv_bfloat16*)dout;
    P.dq = (__nv_bfloat16*)dq; P.dk = (__nv_bfloat16*)dk; P.dv = (__nv_bfloat16*)dv;
    P.lse2 = lse2; P.delta = delta_ws;
    P.o_stride = o_stride; P.do_stride = do_stride; P.dq_stride = dq_stride; P.dkv_stride = dkv_stride;
    P.B = B; P.heads = heads; P.Lq = Lq; P.Lk = Lk;
    P.scale = scale; P.scale_log2 = scale * 1.4426950408889634f;
    if (head_dim == 72) return launch_pipe<72>(q, k, v, dout, q_stride, kv_stride, do_stride, P, (cudaStream_t)stream);
    return launch_pipe<64>(q, k, v, dout, q_stride, kv_stride, do_stride, P, (cudaStream_t)stream);
}
