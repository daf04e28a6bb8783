// Backward of non-causal softmax attention on the 5th-generation tensor cores (training step; autograd through
// F.scaled_dot_product_attention at /root/reference/src/models/transformer/dit_c2i_DeCo.py:181-185).
//
// Same two deterministic passes as csrc/attention_bwd.cu (no atomics; S and dP are recomputed in each), every product a
// tcgen05.mma with M = 128 and the accumulators in tensor memory; two threads share a row of the 128 x 128 score block
// (warps w and w + 4 address the same 32 TMEM lanes; each takes 64 of the block's 128 keys):
//   pass A  item = (image, head, 128 queries):  for every 128-key block   S = Q K^T, dP = dO V^T          (SS, N = 128)
//           P = exp2(c S - lse2), dS = P (dP - delta) scale -> bf16 -> tensor memory;  dQ += dS K  (TS, B = K MN-major)
//           delta = rowsum(dO . O) is formed here (thread-local) and left in the workspace for pass B
//   pass B  item = (image, head, 128 keys):     for every 128-query tile  S, dP as above (rows = queries)
//           P and dS -> bf16 -> SHARED memory as [query][key] tiles = the MN-major A operands of
//           dV += P^T dO and dK += dS^T Q   (M = keys, K = queries; B = the dO / Q tiles read MN-major)
// Q, K, V, dO tiles come straight from the strided [tokens, 3H] matrices through 4-D TMA maps (16-column boxes, 32-byte
// swizzle = the operand layout; head dim 72 zero-padded to 80, ragged sequence tails zero-filled) -- one tile in shared
// memory serves as K-major operand of the score products and as MN-major operand of the gradient products.
// lse2 comes from the forward (deco_attention_fwd_lse).  The kernels are synchronous inside a CTA (load -> MMA -> row
// math -> MMA); parallelism comes from one persistent CTA per SM over ~7 items each.  Operand forms not exercised by the
// forward kernel (A MN-major from shared memory) are checked by tests/test_gpu_backward.py against autograd.
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace deco {
namespace abt {

template <int D> struct Cfg {
    static constexpr int DP = (D + 15) / 16 * 16;
    static constexpr int NCH = DP / 16;
    static constexpr uint32_t CHUNK = 128 * 32;
    static constexpr uint32_t TILE = NCH * CHUNK;             // [128 rows x DP] bf16, chunk-major SW32
    static constexpr uint32_t PT = 8 * CHUNK;                 // [128 x 128] bf16 (P or dS as an MN-major A operand)
    static constexpr uint32_t SMEM_A = 8 * TILE + 256 + 1024;             // 2 x (Q, dO), 2 x (K, V): loads run one step ahead
    static constexpr uint32_t SMEM_B = 6 * TILE + 2 * PT + 256 + 1024;    // K, V, 2 x (Q, dO), P, dS
};
constexpr int kRows = 128, kThreads = 256;      // two threads per score row: each takes 64 of the 128 keys of a block
constexpr uint32_t kColS = 0, kColDP = 128, kColA = 256, kColDQ = 320;      // pass A
constexpr uint32_t kColDV = 256, kColDK = 336;                              // pass B

struct Maps { CUtensorMap q, k, v, dout; };

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

template <int D>
__device__ __forceinline__ void load_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar, int row, int h, int b) {
#pragma unroll
    for (int c = 0; c < Cfg<D>::NCH; ++c) tma_load_4d(dst + c * Cfg<D>::CHUNK, map, bar, 16 * c, row, h, b);
}

// D[128 x N] = A[128 x DP] . B[N x DP]^T, both K-major SW32 tiles (score products)
template <int D>
__device__ __forceinline__ void mma_scores(uint32_t dcol, uint32_t sa, uint32_t sb) {
    constexpr uint32_t idesc = make_idesc_major(128, 128, 0, 0);
    const uint64_t da = make_umma_desc(sa, 16, 256, 6), db = make_umma_desc(sb, 16, 256, 6);
#pragma unroll
    for (int kc = 0; kc < Cfg<D>::NCH; ++kc)
        umma_bf16(dcol, da + (uint64_t)kc * (Cfg<D>::CHUNK >> 4), db + (uint64_t)kc * (Cfg<D>::CHUNK >> 4), idesc, kc ? 1u : 0u);
}

// ------------------------------------------------------------------------------------------------ pass A: dQ (+ delta)
template <int D>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_dq_tc_kernel(const __grid_constant__ Maps M, const Params P)
{
    using C = Cfg<D>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    // [2 x (Q, dO)] per item parity, [2 x (K, V)] per step parity: the loads of step s + 1 are issued before step s computes
    auto sQ = [&](int n) { return base + (uint32_t)(n & 1) * 2 * C::TILE; };
    auto sdO = [&](int n) { return sQ(n) + C::TILE; };
    auto sK = [&](int st) { return base + 4 * C::TILE + (uint32_t)(st & 1) * 2 * C::TILE; };
    auto sV = [&](int st) { return sK(st) + C::TILE; };
    const uint32_t bars = base + 8 * C::TILE;
    auto bar_ld = [&](int st) { return bars + 8u * (st & 1); };
    const uint32_t bar_mma = bars + 16, slot = bars + 24;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, half = tid >> 7;                      // score row of this thread, and which 64 keys of a block it takes
    if (tid == 0) { mbar_init(bar_ld(0), 1); mbar_init(bar_ld(1), 1); mbar_init(bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int nq = (P.Lq + kRows - 1) / kRows, nk = (P.Lk + kRows - 1) / kRows;
    const int nitems = P.B * P.heads * nq;
    const int nlocal = (int)blockIdx.x < nitems ? (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nsteps = nlocal * nk;                                // flat (item, key block) sequence of this CTA
    uint32_t ph_mma = 0;
    constexpr uint32_t idesc_dq = make_idesc_major(128, C::DP, 0, 1);
    auto decode = [&](int n, int& qi, int& h, int& b) {
        const int item = blockIdx.x + n * gridDim.x;
        qi = item % nq; h = (item / nq) % P.heads; b = item / (nq * P.heads);
    };
    auto issue_loads = [&](int st) {                               // thread 0 only
        if (st >= nsteps) return;
        const int n = st / nk, j = st - n * nk;
        int qi, h, b;
        decode(n, qi, h, b);
        mbar_expect_tx(bar_ld(st), (j == 0 ? 4 : 2) * C::TILE);
        if (j == 0) {
            load_tile<D>(sQ(n), &M.q, bar_ld(st), qi * kRows, h, b);
            load_tile<D>(sdO(n), &M.dout, bar_ld(st), qi * kRows, h, b);
        }
        load_tile<D>(sK(st), &M.k, bar_ld(st), j * kRows, h, b);
        load_tile<D>(sV(st), &M.v, bar_ld(st), j * kRows, h, b);
    };
    if (tid == 0) issue_loads(0);

    for (int n = 0; n < nlocal; ++n) {
        int qi, h, b;
        decode(n, qi, h, b);
        const int row = qi * kRows + r;                            // query of this thread
        const bool live = row < P.Lq;
        // delta = rowsum(dO . O), thread-local from the two global rows (the tiles in shared memory are swizzled operands)
        float delta = 0.f, lse = 0.f;
        if (live) {
            const uint4* orow = reinterpret_cast<const uint4*>(P.o + ((long long)b * P.Lq + row) * P.o_stride + (long long)h * D);
            const uint4* drow = reinterpret_cast<const uint4*>(P.dout + ((long long)b * P.Lq + row) * P.do_stride + (long long)h * D);
#pragma unroll
            for (int c = 0; c < D / 8; ++c) {
                const uint4 a = __ldg(orow + c), g = __ldg(drow + c);
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 x = unpack_bf2(aw[e]), y = unpack_bf2(gw[e]);
                    delta = fmaf(x.x, y.x, fmaf(x.y, y.y, delta));
                }
            }
            const long long si = ((long long)b * P.heads + h) * P.Lq + row;
            lse = __ldg(P.lse2 + si);
            if (half == 0) P.delta[si] = delta;
        }
        for (int j = 0; j < nk; ++j) {
            const int st = n * nk + j;
            if (tid == 0) issue_loads(st + 1);                     // its buffers were released by step st - 1's last MMA wait
            mbar_wait(bar_ld(st), (uint32_t)((st >> 1) & 1));
            if (tid == 0) {
                tc_fence_after();
                mma_scores<D>(tmem + kColS, sQ(n), sK(st));
                mma_scores<D>(tmem + kColDP, sdO(n), sV(st));
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
            tc_fence_after();
            const int kvalid = P.Lk - j * kRows;                   // keys of this block that exist
#pragma unroll 1
            for (int ch = 2 * half; ch < 2 * half + 2; ++ch) {
                uint32_t s[32], dp[32];
                tmem_ld32(trow + kColS + (uint32_t)(ch * 32), s);
                tmem_ld32(trow + kColDP + (uint32_t)(ch * 32), dp);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float v[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int key = ch * 32 + 2 * i + e;
                        const float p = (live && key < kvalid) ? ex2f(fmaf(__uint_as_float(s[2 * i + e]), P.scale_log2, -lse)) : 0.f;
                        v[e] = p * (__uint_as_float(dp[2 * i + e]) - delta) * P.scale;
                    }
                    pk[i] = pack_bf2(v[0], v[1]);
                }
                tmem_st16(trow + kColA + (uint32_t)(ch * 16), pk);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint64_t dk = make_umma_desc(sK(st), kRows * 32, 256, 6);   // K block read MN-major: [keys (K) x d (N)]
#pragma unroll
                for (int ks = 0; ks < kRows / 16; ++ks)
                    umma_bf16_ts(tmem + kColDQ, tmem + kColA + (uint32_t)(ks * 8), dk + (uint64_t)ks * (512 >> 4), idesc_dq,
                                 (j > 0 || ks > 0) ? 1u : 0u);
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;               // this step's K / V tiles and the A columns are free again
        }
        tc_fence_after();
        {
            // (the loads are warp-wide .sync.aligned instructions: every lane executes them, only live rows store)
            __nv_bfloat16* out = P.dq + ((long long)b * P.Lq + (live ? row : 0)) * P.dq_stride + (long long)h * D;
            constexpr int kSplit = (D / 8 + 1) / 2 * 8;           // the two threads of a row split its D columns
#pragma unroll
            for (int cc = 0; cc < kSplit; cc += 8) {
                const int c0 = half * kSplit + cc;
                if (c0 >= D) continue;                             // (warp-uniform: half is per warp)
                uint32_t r8[8];
                tmem_ld8(trow + kColDQ + (uint32_t)c0, r8);
                tmem_ld_wait();
                if (live)
                    *reinterpret_cast<uint4*>(out + c0) = make_uint4(pack_bf2(__uint_as_float(r8[0]), __uint_as_float(r8[1])),
                                                                     pack_bf2(__uint_as_float(r8[2]), __uint_as_float(r8[3])),
                                                                     pack_bf2(__uint_as_float(r8[4]), __uint_as_float(r8[5])),
                                                                     pack_bf2(__uint_as_float(r8[6]), __uint_as_float(r8[7])));
            }
        }
        tc_fence_before();
        __syncthreads();                                           // the accumulators are read out before the next item's MMAs
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ pass B: dK, dV
template <int D>
__global__ void __launch_bounds__(kThreads, 1) attn_bwd_dkv_tc_kernel(const __grid_constant__ Maps M, const Params P)
{
    using C = Cfg<D>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    // K, V of the item (single buffer: reloaded at the item boundary), [2 x (Q, dO)] per step parity (loaded one step ahead)
    const uint32_t sK = base, sV = base + C::TILE;
    auto sQ = [&](int st) { return base + 2 * C::TILE + (uint32_t)(st & 1) * 2 * C::TILE; };
    auto sdO = [&](int st) { return sQ(st) + C::TILE; };
    const uint32_t sP = base + 6 * C::TILE, sdS = sP + C::PT;
    const uint32_t bars = sdS + C::PT;
    auto bar_ld = [&](int st) { return bars + 8u * (st & 1); };
    const uint32_t bar_kv = bars + 16, bar_mma = bars + 24, slot = bars + 32;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, half = tid >> 7;                      // score row of this thread, and which 64 keys of a block it takes
    if (tid == 0) { mbar_init(bar_ld(0), 1); mbar_init(bar_ld(1), 1); mbar_init(bar_kv, 1); mbar_init(bar_mma, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int nq = (P.Lq + kRows - 1) / kRows, nk = (P.Lk + kRows - 1) / kRows;
    const int nitems = P.B * P.heads * nk;
    const int nlocal = (int)blockIdx.x < nitems ? (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nsteps = nlocal * nq;                                // flat (item, query tile) sequence of this CTA
    uint32_t ph_mma = 0, ph_kv = 0;
    constexpr uint32_t idesc_g = make_idesc_major(128, C::DP, 1, 1);    // A = P / dS tile MN-major, B = dO / Q tile MN-major
    auto decode = [&](int n, int& kj, int& h, int& b) {
        const int item = blockIdx.x + n * gridDim.x;
        kj = item % nk; h = (item / nk) % P.heads; b = item / (nk * P.heads);
    };
    auto issue_qo = [&](int st) {                                  // thread 0 only: query tile of step st
        if (st >= nsteps) return;
        const int n = st / nq, i = st - n * nq;
        int kj, h, b;
        decode(n, kj, h, b);
        mbar_expect_tx(bar_ld(st), 2 * C::TILE);
        load_tile<D>(sQ(st), &M.q, bar_ld(st), i * kRows, h, b);
        load_tile<D>(sdO(st), &M.dout, bar_ld(st), i * kRows, h, b);
    };
    if (tid == 0) issue_qo(0);

    for (int n = 0; n < nlocal; ++n) {
        int kj, h, b;
        decode(n, kj, h, b);
        const int kvalid = P.Lk - kj * kRows;
        if (tid == 0) {             // the previous item's last MMAs (the readers of K / V) were waited for
            mbar_expect_tx(bar_kv, 2 * C::TILE);
            load_tile<D>(sK, &M.k, bar_kv, kj * kRows, h, b);
            load_tile<D>(sV, &M.v, bar_kv, kj * kRows, h, b);
        }
        for (int i = 0; i < nq; ++i) {
            const int st = n * nq + i;
            const int row = i * kRows + r;                         // query of this thread in tile i
            const bool live = row < P.Lq;
            float lse = 0.f, delta = 0.f;
            if (live) {
                const long long si = ((long long)b * P.heads + h) * P.Lq + row;
                lse = __ldg(P.lse2 + si);
                delta = P.delta[si];                               // written by pass A (an earlier kernel in the stream)
            }
            if (tid == 0) issue_qo(st + 1);                        // its buffers were released by step st - 1's MMA wait
            if (i == 0) { mbar_wait(bar_kv, ph_kv); ph_kv ^= 1; }
            mbar_wait(bar_ld(st), (uint32_t)((st >> 1) & 1));
            if (tid == 0) {
                tc_fence_after();
                mma_scores<D>(tmem + kColS, sQ(st), sK);           // rows = queries, columns = keys
                mma_scores<D>(tmem + kColDP, sdO(st), sV);
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int ch = 2 * half; ch < 2 * half + 2; ++ch) {
                uint32_t s[32], dp[32];
                tmem_ld32(trow + kColS + (uint32_t)(ch * 32), s);
                tmem_ld32(trow + kColDP + (uint32_t)(ch * 32), dp);
                tmem_ld_wait();
                uint32_t pp[16], pd[16];
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    float p[2], g[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int key = ch * 32 + 2 * k2 + e;
                        p[e] = (live && key < kvalid) ? ex2f(fmaf(__uint_as_float(s[2 * k2 + e]), P.scale_log2, -lse)) : 0.f;
                        g[e] = p[e] * (__uint_as_float(dp[2 * k2 + e]) - delta) * P.scale;
                    }
                    pp[k2] = pack_bf2(p[0], p[1]);
                    pd[k2] = pack_bf2(g[0], g[1]);
                }
                // element (row = query tid, col = key) of the [query][key] tile: 16-key chunks of 32 bytes per row
#pragma unroll
                for (int c16 = 0; c16 < 2; ++c16) {
                    const int col = ch * 32 + c16 * 16;
                    const uint32_t o0 = sw32_offset(r, col, kRows), o1 = sw32_offset(r, col + 8, kRows);
                    *reinterpret_cast<uint4*>(gen + (sP - base) + o0) = make_uint4(pp[c16 * 8], pp[c16 * 8 + 1], pp[c16 * 8 + 2], pp[c16 * 8 + 3]);
                    *reinterpret_cast<uint4*>(gen + (sP - base) + o1) = make_uint4(pp[c16 * 8 + 4], pp[c16 * 8 + 5], pp[c16 * 8 + 6], pp[c16 * 8 + 7]);
                    *reinterpret_cast<uint4*>(gen + (sdS - base) + o0) = make_uint4(pd[c16 * 8], pd[c16 * 8 + 1], pd[c16 * 8 + 2], pd[c16 * 8 + 3]);
                    *reinterpret_cast<uint4*>(gen + (sdS - base) + o1) = make_uint4(pd[c16 * 8 + 4], pd[c16 * 8 + 5], pd[c16 * 8 + 6], pd[c16 * 8 + 7]);
                }
            }
            fence_proxy_async();                                   // generic-proxy stores -> tensor-core (async proxy) reads
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint64_t dP_ = make_umma_desc(sP, kRows * 32, 256, 6), dS_ = make_umma_desc(sdS, kRows * 32, 256, 6);
                const uint64_t ddo = make_umma_desc(sdO(st), kRows * 32, 256, 6), dq_ = make_umma_desc(sQ(st), kRows * 32, 256, 6);
#pragma unroll
                for (int ks = 0; ks < kRows / 16; ++ks) {          // K = 16 queries per instruction
                    const uint64_t stp = (uint64_t)ks * (512 >> 4);
                    umma_bf16(tmem + kColDV, dP_ + stp, ddo + stp, idesc_g, (i > 0 || ks > 0) ? 1u : 0u);
                    umma_bf16(tmem + kColDK, dS_ + stp, dq_ + stp, idesc_g, (i > 0 || ks > 0) ? 1u : 0u);
                }
                umma_commit(bar_mma);
            }
            mbar_wait(bar_mma, ph_mma); ph_mma ^= 1;               // this step's Q / dO tiles and P / dS are free again
        }
        tc_fence_after();
        {
            const int key = kj * kRows + r;                        // key of this thread (accumulator row)
            constexpr int kSplit = (D / 8 + 1) / 2 * 8;           // the two threads of a row split its D columns
            const bool live = key < P.Lk;
            __nv_bfloat16* okp = P.dk + ((long long)b * P.Lk + (live ? key : 0)) * P.dkv_stride + (long long)h * D;
            __nv_bfloat16* ovp = P.dv + ((long long)b * P.Lk + (live ? key : 0)) * P.dkv_stride + (long long)h * D;
#pragma unroll
            for (int cc = 0; cc < kSplit; cc += 8) {
                const int c0 = half * kSplit + cc;
                if (c0 >= D) continue;
                uint32_t a8[8], b8[8];
                tmem_ld8(trow + kColDK + (uint32_t)c0, a8);
                tmem_ld8(trow + kColDV + (uint32_t)c0, b8);
                tmem_ld_wait();
                if (live) {
                    *reinterpret_cast<uint4*>(okp + c0) = make_uint4(pack_bf2(__uint_as_float(a8[0]), __uint_as_float(a8[1])),
                                                                     pack_bf2(__uint_as_float(a8[2]), __uint_as_float(a8[3])),
                                                                     pack_bf2(__uint_as_float(a8[4]), __uint_as_float(a8[5])),
                                                                     pack_bf2(__uint_as_float(a8[6]), __uint_as_float(a8[7])));
                    *reinterpret_cast<uint4*>(ovp + c0) = make_uint4(pack_bf2(__uint_as_float(b8[0]), __uint_as_float(b8[1])),
                                                                     pack_bf2(__uint_as_float(b8[2]), __uint_as_float(b8[3])),
                                                                     pack_bf2(__uint_as_float(b8[4]), __uint_as_float(b8[5])),
                                                                     pack_bf2(__uint_as_float(b8[6]), __uint_as_float(b8[7])));
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// 4-D view (d, token, head, batch) of a strided [B*L, row_stride] bf16 matrix whose columns are [head][d]
static int make_tmap(CUtensorMap* map, const void* ptr, int D, long long L, int heads, int B, long long row_stride) {
    PFN_encodeTiled enc = get_tensormap_encoder();
    if (!enc) { deco_set_error("cuTensorMapEncodeTiled entry point not available"); return DECO_ERR_DRIVER; }
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)heads, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)row_stride * 2, (cuuint64_t)D * 2, (cuuint64_t)L * (cuuint64_t)row_stride * 2};
    cuuint32_t box[4] = {16, (cuuint32_t)kRows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { deco_set_error("attention_bwd_tc: cuTensorMapEncodeTiled failed: %d", (int)rc); return DECO_ERR_DRIVER; }
    return DECO_OK;
}

template <int D>
static int launch(const Maps& M, const Params& P, cudaStream_t st) {
    using C = Cfg<D>;
    static unsigned long long attr_done = 0;
    if (!device_setup_done(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_A);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_B);
        if (e != cudaSuccess) { deco_set_error("attention_bwd_tc attr: %s", cudaGetErrorString(e)); return (int)e; }
        mark_device_setup(attr_done);
    }
    const int sms = device_sm_count();
    const int items_a = P.B * P.heads * ((P.Lq + kRows - 1) / kRows), items_b = P.B * P.heads * ((P.Lk + kRows - 1) / kRows);
    attn_bwd_dq_tc_kernel<D><<<items_a < sms ? items_a : sms, kThreads, C::SMEM_A, st>>>(M, P);
    DECO_CHECK_LAUNCH("attn_bwd_dq_tc_kernel");
    attn_bwd_dkv_tc_kernel<D><<<items_b < sms ? items_b : sms, kThreads, C::SMEM_B, st>>>(M, P);
    DECO_CHECK_LAUNCH("attn_bwd_dkv_tc_kernel");
    return DECO_OK;
}

}  // namespace abt
}  // namespace deco

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
    Maps M;
    int rc;
    if ((rc = make_tmap(&M.q, q, head_dim, Lq, heads, B, q_stride))) return rc;
    if ((rc = make_tmap(&M.k, k, head_dim, Lk, heads, B, kv_stride))) return rc;
    if ((rc = make_tmap(&M.v, v, head_dim, Lk, heads, B, kv_stride))) return rc;
    if ((rc = make_tmap(&M.dout, dout, head_dim, Lq, heads, B, do_stride))) return rc;
    Params P;
    P.o = (const __nv_bfloat16*)o; P.dout = (const __nv_bfloat16*)dout;
    P.dq = (__nv_bfloat16*)dq; P.dk = (__nv_bfloat16*)dk; P.dv = (__nv_bfloat16*)dv;
    P.lse2 = lse2; P.delta = delta_ws;
    P.o_stride = o_stride; P.do_stride = do_stride; P.dq_stride = dq_stride; P.dkv_stride = dkv_stride;
    P.B = B; P.heads = heads; P.Lq = Lq; P.Lk = Lk;
    P.scale = scale; P.scale_log2 = scale * 1.4426950408889634f;
    if (head_dim == 72) return launch<72>(M, P, (cudaStream_t)stream);
    return launch<64>(M, P, (cudaStream_t)stream);
}
