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


// ------------------------------------------------------------------------------------------------ pipelined kernels
// The two kernels above run load -> MMA -> row math -> MMA in sequence inside a CTA.  The pipelined form below overlaps
// the row math of step s with the gradient products of step s - 1 and the score products of step s + 1 / s + 2:
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
#define TR_T(var) const long long var = clock64()
#define TR_ADD(slot, t0) atomicAdd(&g_trace[PASS][slot], (unsigned long long)(clock64() - (t0)))
#else
#define TR_T(var)
#define TR_ADD(slot, t0)
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
                mbar_wait(ld_empty(s), eph);
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
                mbar_wait(ld_full(s), (uint32_t)((st / NS) & 1));
                if (which == 0) { TR_ADD(3, t0); }
                TR_T(t1);
                mbar_wait(math_done(b), (uint32_t)(((st >> 1) & 1) ^ 1));      // step st - 2 has read this buffer (free at first)
                if (which == 0) { TR_ADD(0, t1); }
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
                mbar_wait(math_done(b), (uint32_t)((st >> 1) & 1));
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
                const uint32_t srow = scr + (uint32_t)lane * D * 2 + (uint32_t)(half * kSplit) * 2;
#pragma unroll
                for (int c = 0; c < 32; c += 8)
                    sts128u(srow + c * 2, pack_bf2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])),
                            pack_bf2(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])),
                            pack_bf2(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5])),
                            pack_bf2(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7])));
                if (kSplit > 32 && half == 0)
                    sts128u(srow + 64, pack_bf2(__uint_as_float(w[0]), __uint_as_float(w[1])), pack_bf2(__uint_as_float(w[2]), __uint_as_float(w[3])),
                            pack_bf2(__uint_as_float(w[4]), __uint_as_float(w[5])), pack_bf2(__uint_as_float(w[6]), __uint_as_float(w[7])));
                asm volatile("bar.sync %0, 64;" :: "r"(grp + 1) : "memory");
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
            if (tid == 0) { TR_ADD(9, e1); }
        };
        fetch_row_consts(0);
        TR_T(tall);
        for (int n = 0; n < nlocal; ++n) {
            int ri, h, b;
            decode(n, ri, h, b);
            const int row = ri * kRows + r;
            const bool live = row < Lr;
            const float lse = lse_n, delta = delta_n;
            fetch_row_consts(n + 1);               // one item ahead: the latency hides behind this item's steps
            for (int j = 0; j < spi; ++j) {
                const int st = n * spi + j, bb = st & 1;
                const uint32_t tS = trow + (uint32_t)bb * kBufCols + (uint32_t)(32 * half);
                const uint32_t tA = trow + kColA + (uint32_t)bb * 32 + (uint32_t)(16 * half);
                TR_T(m0);
                mbar_wait(s_full(bb), (uint32_t)((st >> 1) & 1));
                if (tid == 0) { TR_ADD(5, m0); }
                TR_T(m1);
                tc_fence_after();
                uint32_t s[32], dp[32];
                tmem_ld32(tS, s);
                tmem_ld32(tS + kCols, dp);
                tmem_ld_wait();
                if (PASS == 0) {
                    const int cvalid = Lc - j * kCols - 32 * half;        // columns of this thread that exist
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float v[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float p = (live && 2 * i + e < cvalid) ? ex2f(fmaf(__uint_as_float(s[2 * i + e]), P.scale_log2, -lse)) : 0.f;
                            v[e] = p * (__uint_as_float(dp[2 * i + e]) - delta) * P.scale;
                        }
                        pk[i] = pack_bf2(v[0], v[1]);
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
                        const float p0 = live ? ex2f(fmaf(__uint_as_float(s[2 * i]), P.scale_log2, -c2.x)) : 0.f;
                        const float p1 = live ? ex2f(fmaf(__uint_as_float(s[2 * i + 1]), P.scale_log2, -c2.z)) : 0.f;
                        pp[i] = pack_bf2(p0, p1);
                        pd[i] = pack_bf2(p0 * (__uint_as_float(dp[2 * i]) - c2.y) * P.scale, p1 * (__uint_as_float(dp[2 * i + 1]) - c2.w) * P.scale);
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
static int make_tmap(CUtensorMap* map, const void* ptr, int D, long long L, int heads, int B, long long row_stride, int box_rows = kRows) {
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

// DECO_ATTN_BWD_TC = "sync" selects the unpipelined kernel pair (A/B measurements); default = the pipelined kernels
static bool use_pipelined() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DECO_ATTN_BWD_TC"); v = (e && e[0] == 's') ? 0 : 1; }
    return v == 1;
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
    if (use_pipelined()) {
        if (head_dim == 72) return launch_pipe<72>(q, k, v, dout, q_stride, kv_stride, do_stride, P, (cudaStream_t)stream);
        return launch_pipe<64>(q, k, v, dout, q_stride, kv_stride, do_stride, P, (cudaStream_t)stream);
    }
    Maps M;
    int rc;
    if ((rc = make_tmap(&M.q, q, head_dim, Lq, heads, B, q_stride))) return rc;
    if ((rc = make_tmap(&M.k, k, head_dim, Lk, heads, B, kv_stride))) return rc;
    if ((rc = make_tmap(&M.v, v, head_dim, Lk, heads, B, kv_stride))) return rc;
    if ((rc = make_tmap(&M.dout, dout, head_dim, Lq, heads, B, do_stride))) return rc;
    if (head_dim == 72) return launch<72>(M, P, (cudaStream_t)stream);
    return launch<64>(M, P, (cudaStream_t)stream);
}
