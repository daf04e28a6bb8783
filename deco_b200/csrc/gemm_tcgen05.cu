// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[M,N] = epilogue(A[M,K] . W[N,K]^T)
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a multi-stage shared-memory ring,
//   * tcgen05.mma (kind::f16, bf16 in / fp32 accumulate, K=16) issued by one thread:
//       cta_group::2 (default): a CTA PAIR (2-CTA cluster = one TPC) computes a 256 x BN tile; each CTA stages its own
//                      128 rows of A and HALF of the W tile, the leader's MMA reads both CTAs' shared memory.  Versus
//                      cta_group::1 this cuts L2->SMEM traffic per MAC by 1/3 and the per-SM operand read rate from
//                      ~100 to 64 B/clk (the 128 B/clk SMEM port is what throttled the 1-CTA version),
//       cta_group::1: 128 x BN tile per CTA (kept for M <= 128 and as the A/B reference),
//   * accumulators in TMEM, double buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1,
//   * epilogue warps read TMEM with tcgen05.ld and fuse bias / SiLU / AdaLN gate + residual / SwiGLU.
//
// Replaces every nn.Linear on the hot path of /root/reference/src/models/transformer/dit_c2i_DeCo.py:
//   s_embedder.proj (:496), t_embedder.mlp (:55-57), adaLN_modulation (:207, all blocks batched into one GEMM),
//   attn.qkv (:176), attn.proj + gated residual (:188,:208), mlp.w1/w3 + SiLU-gate (:113), mlp.w2 + gated residual
//   (:113,:209), dec_net.cond_embed (:404).
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31 = tile rows).
#include "tcgen05.cuh"
#include "tma_host.cuh"
#include <vector>

namespace deco {

constexpr int kBM = 128;
constexpr int kBK = 64;            // 64 bf16 = 128 bytes = one swizzle row
constexpr int kGemmThreads = 256;
constexpr int kEpiWarp0 = 4;
constexpr int kAccStages = 2;

enum GemmEpilogue : int {
    EPI_BIAS = 0,           // out = acc + bias
    EPI_BIAS_SILU = 1,      // out = silu(acc + bias)
    EPI_GATE_RESIDUAL = 2,  // out = resid + gate[row / rows_per_gate] * (acc + bias); resid / out are FP32
    EPI_SWIGLU = 3,         // columns interleaved in 16s: out[:, n/2 + i] = silu(acc[n + i]) * acc[n + 16 + i]
    EPI_BIAS_F32 = 4,       // out = acc + bias, FP32 output (starts the fp32 residual stream)
    // training path: the SwiGLU passes live in the epilogues of the GEMMs next to them (row-per-thread form only)
    EPI_SWIGLU_DUAL = 5,    // EPI_SWIGLU + the pre-activation acc itself as bf16 [M, N] into `resid` (what the backward reads)
    EPI_SWIGLU_BWD = 6,     // acc = du [M, N]; `resid` = y13 bf16 [M, 2N] interleaved [16 a | 16 b]; out = dy13 bf16 [M, 2N]:
                            //   dy[32 g + i] = du b silu'(a), dy[32 g + 16 + i] = du silu(a)   (N a multiple of 16)
};

struct GemmParams {
    void* out;                       // [M, ldo] bf16 (fp32 for EPI_GATE_RESIDUAL / EPI_BIAS_F32)
    long long ldo;
    const float* bias;               // [N] fp32 or null
    const float* resid;              // [M, ldr] fp32 (EPI_GATE_RESIDUAL); bf16 y13 (EPI_SWIGLU_DUAL: written, EPI_SWIGLU_BWD: read)
    long long ldr;
    const __nv_bfloat16* gate;       // [M / rows_per_gate, gate_stride]
    long long gate_stride;
    int rows_per_gate;
    int M, N, K;
    int split_k;                     // > 1 (EPI_BIAS_F32 only): the K loop is cut into split_k slices, one tile each, and the
                                     // partial products are reduced with fp32 atomics into a pre-zeroed output
    int deint_rows;                  // > 0 (EPI_BIAS_F32, staged): output rows come in interleaved groups [16 x first | 16 x
                                     // second] (the w1 / w3 row order of the SwiGLU weight); row r is stored at
                                     // ((r >> 4) & 1) * deint_rows + (r >> 5) * 16 + (r & 15): two stacked plain matrices
};

// Epilogue warps: 4 (one per TMEM lane quarter), or 8 for the SwiGLU training epilogues -- two warps per quarter, each
// draining half of the tile's columns: their per-element math and the pre-activation traffic (128 B in + 128 B out per
// row and 32-column chunk) need more warps in flight than four to keep up with the main loop.
__host__ __device__ constexpr int epi_warps(int epi) { return epi >= 5 ? 8 : 4; }

template <int BN, int CG, int EW = 4> struct GemmCfg {
    static constexpr int kStageBytesA = kBM * kBK * 2;               // this CTA's 128 rows of A
    static constexpr int kStageBytesB = (BN / CG) * kBK * 2;         // this CTA's share of the W tile
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    static constexpr int kEpiStageBytes = EW * 32 * 36 * 4;          // EW epilogue warps x [32][36] fp32
    static constexpr int kBudget = 227 * 1024 - kEpiStageBytes - 1024 - 256;
    static constexpr int kStagesRaw = kBudget / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kTmemCols = (kAccStages * BN > 256) ? 512 : 256;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ------------------------------------------------------------------ epilogue on one 32-column chunk
// A warp owns 32 tile rows (TMEM lanes); tcgen05.ld 32x32b hands each thread one ROW of the chunk.
//   direct : each thread reads/writes its own row (row-strided 16-byte accesses, 32 cache lines per instruction).
//   staged : raw fp32 accumulators go through a per-warp shared-memory tile [32 rows][36 floats] (conflict-free for
//            128-bit accesses both ways) and the global side runs with lane = (row % 4, 4 consecutive columns): 8 lanes
//            cover 128 contiguous bytes of a row.  Costs SMEM bandwidth, which only the 2-CTA MMA has to spare.
constexpr int kStageLd = 36;
constexpr int kStageFloatsPerWarp = 32 * kStageLd;

// Residual / gate operands of one staged chunk, fetched EARLY (before the accumulator is ready) so that their
// global-memory latency overlaps the MMA main loop instead of sitting in the epilogue's critical path.
struct ResidChunk {
    float4 rv[8];     // residual, rows it*4 + (lane >> 3), columns 4*(lane & 7) .. +3
    uint2 gv[8];      // gate (bf16 x4) per row; gv[0] only when the 32 rows share one modulation row
};

__device__ __forceinline__ void load_resid_chunk(const GemmParams& P, ResidChunk& rc, long long row0, int n0, int lane,
                                                 bool uniform_gate) {
    const int n = n0 + 4 * (lane & 7), rsub = lane >> 3;
    const bool col_ok = n < P.N;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const long long row = row0 + it * 4 + rsub;
        const bool ok = col_ok && row < P.M;
        rc.rv[it] = ok ? *reinterpret_cast<const float4*>(P.resid + row * P.ldr + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (!uniform_gate || it == 0)
            rc.gv[it] = ok ? __ldg(reinterpret_cast<const uint2*>(P.gate + (row / P.rows_per_gate) * P.gate_stride + n))
                           : make_uint2(0u, 0u);
    }
}

template <int EPI>
__device__ __forceinline__ void epilogue_chunk_staged(const GemmParams& P, const uint32_t (&acc)[32], float* stg,
                                                      long long row0, int n0, int lane, const ResidChunk& rc,
                                                      bool uniform_gate) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * kStageLd + 4 * j) =
            make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                        __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
    __syncwarp();
    const int cg = lane & 7, rsub = lane >> 3;
    const int n = n0 + 4 * cg;
    if (EPI == EPI_SWIGLU) {
        // chunk = [16 x w1-columns | 16 x w3-columns]: lanes cg<4 hold a[4cg..], lanes cg>=4 hold b[4(cg-4)..].
        // Partner lanes (cg ^ 4) swap their float4 and each produces two of the four outputs: all lanes busy.
        const int no = (n0 >> 1) + 4 * (cg & 3) + ((cg >> 2) << 1);
        const bool col_ok = n0 + 4 * (cg & 3) < P.N;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            const long long row = row0 + r;
            const float4 mine = *reinterpret_cast<const float4*>(stg + r * kStageLd + 4 * cg);
            float4 other;
            other.x = __shfl_xor_sync(0xffffffffu, mine.x, 4); other.y = __shfl_xor_sync(0xffffffffu, mine.y, 4);
            other.z = __shfl_xor_sync(0xffffffffu, mine.z, 4); other.w = __shfl_xor_sync(0xffffffffu, mine.w, 4);
            // low lane: outputs 0,1 = silu(a.x)*b.x, silu(a.y)*b.y ; high lane: outputs 2,3
            const bool hi = cg >= 4;
            const float a0 = hi ? other.z : mine.x, a1 = hi ? other.w : mine.y;
            const float b0 = hi ? mine.z : other.x, b1 = hi ? mine.w : other.y;
            if (col_ok && row < P.M)
                *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(P.out) + row * P.ldo + no) =
                    pack_bf2(silu_f(a0) * b0, silu_f(a1) * b1);
        }
        __syncwarp();
        return;
    }
    if (n < P.N) {
        float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
        if (P.bias) bz = __ldg(reinterpret_cast<const float4*>(P.bias + n));
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            const long long row = row0 + r;
            float4 v = *reinterpret_cast<const float4*>(stg + r * kStageLd + 4 * cg);
            v.x += bz.x; v.y += bz.y; v.z += bz.z; v.w += bz.w;
            if (row < P.M) {
                if (EPI == EPI_GATE_RESIDUAL) {
                    // fp32 residual stream: x <- x + gate * (acc + bias); may run in place (out == resid)
                    const float4 rv = rc.rv[it];
                    const uint2 gv = uniform_gate ? rc.gv[0] : rc.gv[it];
                    const float2 g0 = unpack_bf2(gv.x), g1 = unpack_bf2(gv.y);
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(P.out) + row * P.ldo + n) =
                        make_float4(fmaf(g0.x, v.x, rv.x), fmaf(g0.y, v.y, rv.y), fmaf(g1.x, v.z, rv.z), fmaf(g1.y, v.w, rv.w));
                } else if (EPI == EPI_BIAS_F32) {
                    const long long orow = P.deint_rows > 0 ? ((row >> 4) & 1) * P.deint_rows + (row >> 5) * 16 + (row & 15) : row;
                    float* o = reinterpret_cast<float*>(P.out) + orow * P.ldo + n;
                    if (P.split_k > 1) {        // partial product of one K slice (output zeroed by the launcher): ONE vector
                        // reduction per 16 bytes -- a REDG costs the SM ~1.3 clocks per lane whatever its width, and
                        // four scalar ones per float4 made the epilogue of a split tile (42 k clocks) longer than its
                        // half-depth main loop
                        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                     :: "l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                    } else {
                        *reinterpret_cast<float4*>(o) = v;
                    }
                } else {
                    if (EPI == EPI_BIAS_SILU) { v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w); }
                    uint2 w;
                    w.x = pack_bf2(v.x, v.y); w.y = pack_bf2(v.z, v.w);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(P.out) + row * P.ldo + n) = w;
                }
            }
        }
    }
    __syncwarp();
}

// bf16 copy of a warp's 32 x 32 accumulator chunk to dst[row0 + r][n0 + c] through the warp's staging tile: a store
// instruction covers 4 rows x 64 contiguous bytes (row-per-thread stores touch 32 lines per instruction).
__device__ __forceinline__ void store_chunk_bf16_staged(const uint32_t (&acc)[32], float* stg, __nv_bfloat16* dst, long long ld,
                                                        long long row0, int n0, int lane, int M, int N) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * kStageLd + 4 * j) =
            make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                        __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
    __syncwarp();
    const int cg = lane & 7, rsub = lane >> 3;
    const int n = n0 + 4 * cg;
    if (n < N) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            const long long row = row0 + r;
            const float4 v = *reinterpret_cast<const float4*>(stg + r * kStageLd + 4 * cg);
            if (row < M) *reinterpret_cast<uint2*>(dst + row * ld + n) = make_uint2(pack_bf2(v.x, v.y), pack_bf2(v.z, v.w));
        }
    }
    __syncwarp();
}

template <int EPI>
__device__ __forceinline__ void epilogue_chunk_direct(const GemmParams& P, const uint32_t (&acc)[32], long long row, int n0) {
    if (row >= P.M) return;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
    if (P.bias) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            if (n0 + i < P.N) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias + n0 + i));
                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
            }
        }
    }
    if (EPI == EPI_SWIGLU || EPI == EPI_SWIGLU_DUAL) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(P.out) + row * P.ldo + (n0 >> 1);
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            w[i] = pack_bf2(silu_f(v[2 * i]) * v[16 + 2 * i], silu_f(v[2 * i + 1]) * v[16 + 2 * i + 1]);
        *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(o + 8) = make_uint4(w[4], w[5], w[6], w[7]);
        return;
    }
    if (EPI == EPI_GATE_RESIDUAL) {
        float* o = reinterpret_cast<float*>(P.out) + row * P.ldo + n0;
        const float* r = P.resid + row * P.ldr + n0;
        const __nv_bfloat16* gt = P.gate + (row / P.rows_per_gate) * P.gate_stride + n0;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            if (n0 + i < P.N) {
                const float4 r0 = *reinterpret_cast<const float4*>(r + i);
                const float4 r1 = *reinterpret_cast<const float4*>(r + i + 4);
                const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gt + i));
                const float2 g0 = unpack_bf2(gv.x), g1 = unpack_bf2(gv.y), g2 = unpack_bf2(gv.z), g3 = unpack_bf2(gv.w);
                *reinterpret_cast<float4*>(o + i) = make_float4(fmaf(g0.x, v[i], r0.x), fmaf(g0.y, v[i + 1], r0.y),
                                                                fmaf(g1.x, v[i + 2], r0.z), fmaf(g1.y, v[i + 3], r0.w));
                *reinterpret_cast<float4*>(o + i + 4) = make_float4(fmaf(g2.x, v[i + 4], r1.x), fmaf(g2.y, v[i + 5], r1.y),
                                                                    fmaf(g3.x, v[i + 6], r1.z), fmaf(g3.y, v[i + 7], r1.w));
            }
        }
        return;
    }
    if (EPI == EPI_BIAS_F32) {
        float* o = reinterpret_cast<float*>(P.out) + row * P.ldo + n0;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
            if (n0 + i < P.N) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        return;
    }
    if (EPI == EPI_BIAS_SILU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = silu_fast(v[i]);
    }
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(P.out) + row * P.ldo + n0;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        if (n0 + i < P.N) {
            *reinterpret_cast<uint4*>(o + i) = make_uint4(pack_bf2(v[i], v[i + 1]), pack_bf2(v[i + 2], v[i + 3]),
                                                          pack_bf2(v[i + 4], v[i + 5]), pack_bf2(v[i + 6], v[i + 7]));
        }
    }
}

// EPI_SWIGLU_BWD.  The 32 du columns [n0, n0 + 32) of a row are two groups of 16; their pre-activations are the 64 bf16
// = 128 contiguous bytes y13[row][2 n0 ...] = [a0 | b0 | a1 | b1] (16 each), and dy13 goes to the same place of the output.
// Global accesses are coalesced: in iteration `it` lane l owns the 16-byte piece (l & 7) of row 4 it + (l >> 3), so one
// instruction covers 4 rows x 128 bytes; du comes from the warp's staging tile, the (a, b) partner piece by shuffle.
struct SwigluPre { uint4 q[8]; };

__device__ __forceinline__ void load_swiglu_pre(const GemmParams& P, SwigluPre& y, long long row0, int n0, int lane) {
    const int pc = lane & 7, rsub = lane >> 3;
    const bool col_ok = n0 + 16 * (pc >> 2) < P.N;
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(P.resid) + 2 * n0 + 8 * pc;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const long long row = row0 + it * 4 + rsub;
        y.q[it] = (col_ok && row < P.M) ? __ldg(reinterpret_cast<const uint4*>(src + row * P.ldr)) : make_uint4(0u, 0u, 0u, 0u);
    }
}

__device__ __forceinline__ void epilogue_chunk_swiglu_bwd(const GemmParams& P, const uint32_t (&acc)[32], float* stg,
                                                          const SwigluPre& y, long long row0, int n0, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * kStageLd + 4 * j) =
            make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                        __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
    __syncwarp();
    const int pc = lane & 7, rsub = lane >> 3;
    const int grp = pc >> 2, is_b = (pc >> 1) & 1, half = pc & 1;     // piece = group, a / b part, elements 8 half .. + 7
    const bool col_ok = n0 + 16 * grp < P.N;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(P.out) + 2 * n0 + 8 * pc;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + rsub;
        const long long row = row0 + r;
        const uint4 mine = y.q[it];
        uint4 other;                        // the partner's piece: b for an a-lane and vice versa (lane ^ 2)
        other.x = __shfl_xor_sync(0xffffffffu, mine.x, 2); other.y = __shfl_xor_sync(0xffffffffu, mine.y, 2);
        other.z = __shfl_xor_sync(0xffffffffu, mine.z, 2); other.w = __shfl_xor_sync(0xffffffffu, mine.w, 2);
        const uint4 qa = is_b ? other : mine, qb = is_b ? mine : other;
        const float* dp = stg + r * kStageLd + 16 * grp + 8 * half;
        const float4 d0 = *reinterpret_cast<const float4*>(dp), d1 = *reinterpret_cast<const float4*>(dp + 4);
        const float du[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const uint32_t wa[4] = {qa.x, qa.y, qa.z, qa.w}, wb[4] = {qb.x, qb.y, qb.z, qb.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = unpack_bf2(wa[e]), b = unpack_bf2(wb[e]);
            // sigmoid(a) = 1/2 + 1/2 tanh(a/2): one MUFU op; silu = a s, silu' = s (1 + a (1 - s))
            float t0, t1;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(0.5f * a.x));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(0.5f * a.y));
            const float s0 = fmaf(0.5f, t0, 0.5f), s1 = fmaf(0.5f, t1, 0.5f);
            const float f0 = is_b ? a.x * s0 : b.x * (s0 * fmaf(a.x, 1.0f - s0, 1.0f));
            const float f1 = is_b ? a.y * s1 : b.y * (s1 * fmaf(a.y, 1.0f - s1, 1.0f));
            o[e] = pack_bf2(du[2 * e] * f0, du[2 * e + 1] * f1);
        }
        if (col_ok && row < P.M) *reinterpret_cast<uint4*>(dst + row * P.ldo) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncwarp();
}

// ------------------------------------------------------------------ tile schedule of the persistent CTA groups
// Static, but balanced: (1) the tiles of a RAGGED last column (a narrower MMA, a fraction of a full tile's time) are
// numbered after all full-width tiles, (2) successive rounds run over the CTA groups in alternating direction ("snake"),
// so the groups that drew the extra tile of the last full round get the cheap tiles or none.  N = 1152 in 256-wide tiles on
// 8192 rows = 128 full + 32 half tiles on 74 CTA pairs: every pair ends with exactly two tile-times of work (round robin in
// row-major order: three), which is what lets the 1152-wide GEMMs of the training step use 256-wide tiles at all -- in
// 192-wide tiles they are 192 tiles = 2.6 -> 3 rounds of a tile shape that feeds the tensor pipe a quarter slower.
struct TileSchedule {
    int num_n, num_mn, num_tiles, group, groups, full_mn;     // full_mn: tiles of the full-width columns (0 = no ragged column)
    __host__ __device__ __forceinline__ int tile_of_round(int round) const {      // -1 = this group is done
        const int t = round * groups + ((round & 1) ? groups - 1 - group : group);
        return t < num_tiles ? t : -1;
    }
    __host__ __device__ __forceinline__ void coords(int tile, int& m_blk, int& n_blk, int& ks) const {
        const int mn = tile % num_mn;
        ks = tile / num_mn;
        if (full_mn > 0) {
            if (mn < full_mn) { m_blk = mn / (num_n - 1); n_blk = mn % (num_n - 1); }
            else { m_blk = mn - full_mn; n_blk = num_n - 1; }
        } else { m_blk = mn / num_n; n_blk = mn % num_n; }
    }
};

// ------------------------------------------------------------------ kernel
// CG = 1: one CTA per 128 x BN tile.   CG = 2: a 2-CTA cluster per 256 x BN tile (rank 0 = leader issues the MMAs).
// TN = true: both operands are given TRANSPOSED in global memory -- At [K, M] and Wt [K, N], row-major -- and are staged
// MN-major: a {64 (MN) x 64 (K)} TMA box lands as 64 K-rows of 128 bytes (SW128 atoms of 8 rows), 64-wide MN blocks
// 8 KB apart (LBO), 8-row K groups 1 KB apart (SBO); the instruction descriptor flags both operands MN-major.  This is
// the wgrad contraction dW = dY^T . X read straight from the row-major activations (no transposed copies).
template <int BN, int EPI, int CG, bool STAGED, bool TN = false>
__global__ void __launch_bounds__(128 + 32 * epi_warps(EPI), 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmParams P)
{
    constexpr int kEW = epi_warps(EPI);
    using Cfg = GemmCfg<BN, CG, kEW>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128-byte swizzle atom (8 rows x 128 B); identical offsets in both CTAs
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_stage_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
    const uint32_t bar_base = epi_stage_base + Cfg::kEpiStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + kAccStages + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 2 * kAccStages);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool is_leader = cta_rank == 0;
    const int num_m = (P.M + kBM * CG - 1) / (kBM * CG), num_n = (P.N + BN - 1) / BN;
    const int num_mn = num_m * num_n;
    const int split_k = (EPI == EPI_BIAS_F32 && P.split_k > 1) ? P.split_k : 1;
    const int num_tiles = num_mn * split_k;
    const int num_k = (P.K + kBK - 1) / kBK;
    const int k_per = (num_k + split_k - 1) / split_k;      // host guarantees every slice is non-empty
    const int tile0 = blockIdx.x / CG, tile_stride = gridDim.x / CG;
    // a ragged last column tile runs a NARROWER MMA (N = the columns that are left, when a multiple of 32) instead of
    // multiplying zero padding: under the board power cap wasted MMA work costs clock, not just tensor-pipe time
    const int n_rem = P.N - (num_n - 1) * BN;
    const int n_last = (!TN && n_rem < BN && n_rem % 32 == 0) ? n_rem : BN;
    TileSchedule sched;
    sched.num_n = num_n; sched.num_mn = num_mn; sched.num_tiles = num_tiles; sched.group = tile0; sched.groups = tile_stride;
    // ragged tiles last only while A stays in L2 (126 MB): they re-read every row block of A long after the full-width tiles
    // of that block ran.  On the 131072-row inference GEMMs the same order cost 1.2 % of the step in re-streamed A (measured
    // in the fused kernels, profiles/tile_schedule_r2.txt); there the row-major order stays.
    sched.full_mn = (n_last < BN && num_n > 1 && (long long)P.M * P.K * 2 <= (64ll << 20)) ? num_m * (num_n - 1) : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEW * CG); }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CG == 2) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
        else tmem_alloc(tmem_slot, Cfg::kTmemCols);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                   // prologue overlapped the previous kernel's tail; its outputs are visible from here
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer (every CTA loads its own A rows and its share of W) =====================
        if (TN) {
            // MN-major staging: 64-wide MN blocks, each a {64 x 64} box at (MN coordinate, K row) -- kBM / 64 boxes of A and
            // BN / CG / 64 of W per K block.  One thread gets a bulk tensor load out only every ~250 clocks whatever the box
            // (profiles/tma_bench_r2.txt): four boxes from lane 0 were 1000 clocks per K block against 512 clocks of MMA, i.e.
            // the weight-gradient GEMMs sat at half the tensor rate.  Each box is issued by its own lane instead.
            constexpr int kBoxesA = kBM / 64, kBoxesB = (BN / CG) / 64;
            int stage = 0; uint32_t phase = 0;
            for (int round = 0, tile; (tile = sched.tile_of_round(round)) >= 0; ++round) {
                int m_blk, n_blk, ks;
                sched.coords(tile, m_blk, n_blk, ks);
                const int arow = (m_blk * CG + (int)cta_rank) * kBM;
                const int brow = n_blk * BN + (int)cta_rank * (BN / CG);
                const int kb0 = ks * k_per, kb1 = min(num_k, kb0 + k_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                    const uint32_t sb = sa + Cfg::kStageBytesA;
                    if (lane == 0 && (CG == 1 || is_leader)) mbar_expect_tx(full_bar(stage), CG * Cfg::kStageBytes);
                    __syncwarp();
                    const uint32_t fb = CG == 2 ? mapa_shared(full_bar(stage), 0) : full_bar(stage);
                    if (lane < kBoxesA) {
                        if (CG == 2) tma_load_2d_2sm(sa + lane * 8192, &tmap_a, fb, arow + 64 * lane, kb * kBK);
                        else tma_load_2d(sa + lane * 8192, &tmap_a, fb, arow + 64 * lane, kb * kBK);
                    } else if (lane < kBoxesA + kBoxesB) {
                        const int j = lane - kBoxesA;
                        if (CG == 2) tma_load_2d_2sm(sb + j * 8192, &tmap_b, fb, brow + 64 * j, kb * kBK);
                        else tma_load_2d(sb + j * 8192, &tmap_b, fb, brow + 64 * j, kb * kBK);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int round = 0, tile; (tile = sched.tile_of_round(round)) >= 0; ++round) {
                int m_blk, n_blk, ks;
                sched.coords(tile, m_blk, n_blk, ks);
                const int arow = (m_blk * CG + (int)cta_rank) * kBM;
                const int brow = n_blk * BN + (int)cta_rank * ((n_blk == num_n - 1 ? n_last : BN) / CG);
                const int kb0 = ks * k_per, kb1 = min(num_k, kb0 + k_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                    const uint32_t sb = sa + Cfg::kStageBytesA;
                    if (CG == 1) {
                        mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
                        tma_load_2d(sa, &tmap_a, full_bar(stage), kb * kBK, arow);
                        tma_load_2d(sb, &tmap_b, full_bar(stage), kb * kBK, brow);
                    } else {
                        // both CTAs' bytes are credited to the LEADER's full barrier, which expects the pair's total
                        if (is_leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
                        const uint32_t fb = mapa_shared(full_bar(stage), 0);
                        tma_load_2d_2sm(sa, &tmap_a, fb, kb * kBK, arow);
                        tma_load_2d_2sm(sb, &tmap_b, fb, kb * kBK, brow);
                    }
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0 && is_leader) {
            constexpr uint32_t idesc = TN ? make_idesc_major(kBM * CG, BN, 1, 1) : make_idesc(kBM * CG, BN);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int round = 0, tile; (tile = sched.tile_of_round(round)) >= 0; ++round) {
                int m_blk, n_blk, ks;
                sched.coords(tile, m_blk, n_blk, ks);
                mbar_wait(tempty_bar(as), aphase ^ 1);   // epilogues (of both CTAs) have drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                const uint32_t idesc_t = (!TN && n_blk == num_n - 1) ? make_idesc(kBM * CG, n_last) : idesc;
                const int kb0 = ks * k_per, kb1 = min(num_k, kb0 + k_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
                    const uint64_t da = TN ? make_umma_desc(sa, 8192, 1024, 2) : make_sw128_desc(sa);
                    const uint64_t db = TN ? make_umma_desc(sa + Cfg::kStageBytesA, 8192, 1024, 2) : make_sw128_desc(sa + Cfg::kStageBytesA);
                    // K-major: advance 16 elements = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field;
                    // MN-major: 16 K rows of 128 bytes = two 1 KB atoms: +128
                    constexpr uint64_t kstep = TN ? 128 : 2;
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        if (CG == 2) umma_bf16_2sm(tmem_d, da + kstep * k, db + kstep * k, idesc_t, ((kb - kb0) | k) ? 1u : 0u);
                        else umma_bf16(tmem_d, da + kstep * k, db + kstep * k, idesc_t, ((kb - kb0) | k) ? 1u : 0u);
                    }
                    // frees the smem stage (in both CTAs) once these MMAs retire
                    if (CG == 2) umma_commit_2sm(empty_bar(stage)); else umma_commit(empty_bar(stage));
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                if (CG == 2) umma_commit_2sm(tfull_bar(as)); else umma_commit(tfull_bar(as));   // accumulator -> epilogue
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================== epilogue (each CTA drains its own 128 accumulator rows) =====================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        // with 8 epilogue warps, warps 4-7 drain the first half of the tile's 32-column chunks and warps 8-11 the second
        constexpr int kChunks = BN / 32 / (kEW / 4);
        const int c_begin = ((warp - kEpiWarp0) >> 2) * kChunks, c_end = c_begin + kChunks;
        int as = 0; uint32_t aphase = 0;
        float* stg = reinterpret_cast<float*>(smem_raw + (epi_stage_base - smem_u32(smem_raw))) +
                     (warp - kEpiWarp0) * kStageFloatsPerWarp;
        for (int round = 0, tile; (tile = sched.tile_of_round(round)) >= 0; ++round) {
            int m_blk, n_blk, ks;
            sched.coords(tile, m_blk, n_blk, ks);
            const long long row0 = (long long)(m_blk * CG + (int)cta_rank) * kBM + q * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
            const int nbase = n_blk * BN;
            if (STAGED && EPI == EPI_GATE_RESIDUAL) {
                // residual chunks are fetched one chunk ahead; chunk 0 before the accumulator barrier, so that
                // HBM latency hides behind the MMA main loop
                const bool ug = (P.rows_per_gate % 32) == 0;
                ResidChunk rcA, rcB;
                load_resid_chunk(P, rcA, row0, nbase, lane, ug);
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < BN / 32; c += 2) {
                    uint32_t acc[32];
                    load_resid_chunk(P, rcB, row0, nbase + (c + 1) * 32, lane, ug);
                    tmem_ld32(taddr + (uint32_t)(c * 32), acc);
                    tmem_ld_wait();
                    if (row0 < P.M && nbase + c * 32 < P.N)
                        epilogue_chunk_staged<EPI>(P, acc, stg, row0, nbase + c * 32, lane, rcA, ug);
                    if (c + 2 < BN / 32) load_resid_chunk(P, rcA, row0, nbase + (c + 2) * 32, lane, ug);
                    tmem_ld32(taddr + (uint32_t)((c + 1) * 32), acc);
                    tmem_ld_wait();
                    if (row0 < P.M && nbase + (c + 1) * 32 < P.N)
                        epilogue_chunk_staged<EPI>(P, acc, stg, row0, nbase + (c + 1) * 32, lane, rcB, ug);
                }
            } else if (EPI == EPI_SWIGLU_BWD) {
                // the pre-activations of a chunk are fetched one chunk ahead (the first before the accumulator barrier)
                SwigluPre yA, yB;
                load_swiglu_pre(P, yA, row0, nbase + c_begin * 32, lane);
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
#pragma unroll 1
                for (int c = c_begin; c < c_end; c += 2) {
                    uint32_t acc[32];
                    if (c + 1 < c_end) load_swiglu_pre(P, yB, row0, nbase + (c + 1) * 32, lane);
                    tmem_ld32(taddr + (uint32_t)(c * 32), acc);
                    tmem_ld_wait();
                    if (row0 < P.M && nbase + c * 32 < P.N) epilogue_chunk_swiglu_bwd(P, acc, stg, yA, row0, nbase + c * 32, lane);
                    if (c + 1 < c_end) {
                        if (c + 2 < c_end) load_swiglu_pre(P, yA, row0, nbase + (c + 2) * 32, lane);
                        tmem_ld32(taddr + (uint32_t)((c + 1) * 32), acc);
                        tmem_ld_wait();
                        if (row0 < P.M && nbase + (c + 1) * 32 < P.N)
                            epilogue_chunk_swiglu_bwd(P, acc, stg, yB, row0, nbase + (c + 1) * 32, lane);
                    }
                }
            } else {
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t acc[32];
                    tmem_ld32(taddr + (uint32_t)(c * 32), acc);
                    tmem_ld_wait();
                    const int n0 = nbase + c * 32;
                    if (row0 < P.M && n0 < P.N) {             // warp-uniform guard
                        if (EPI == EPI_SWIGLU_DUAL)         // the pre-activation as the backward will read it
                            store_chunk_bf16_staged(acc, stg, reinterpret_cast<__nv_bfloat16*>(const_cast<float*>(P.resid)), P.ldr,
                                                    row0, n0, lane, P.M, P.N);
                        if (STAGED) { ResidChunk none; epilogue_chunk_staged<EPI>(P, acc, stg, row0, n0, lane, none, false); }
                        else epilogue_chunk_direct<EPI>(P, acc, row0 + lane, n0);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(as), 0));   // TMEM hand-over only
                else mbar_arrive(tempty_bar(as));
            }
            if (++as == kAccStages) { as = 0; aphase ^= 1; }
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        if (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
        else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ------------------------------------------------------------------ host side
// 2-D bf16 row-major [rows, cols] tensor with a {64 x box_rows} box and 128-byte swizzle
static int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    PFN_encodeTiled enc = get_tensormap_encoder();
    if (!enc) { deco_set_error("cuTensorMapEncodeTiled entry point not available"); return DECO_ERR_DRIVER; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { deco_set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return DECO_ERR_DRIVER; }
    return DECO_OK;
}

template <int BN, int EPI, int CG, bool STAGED, bool TN = false>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& P, int max_ctas, cudaStream_t st) {
    using Cfg = GemmCfg<BN, CG, epi_warps(EPI)>;
    auto kern = gemm_bf16_tcgen05_kernel<BN, EPI, CG, STAGED, TN>;
    static unsigned long long attr_done = 0;   // per instantiation
    if (!device_setup_done(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) { deco_set_error("gemm smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        mark_device_setup(attr_done);
    }
    const int tiles = ((P.M + kBM * CG - 1) / (kBM * CG)) * ((P.N + BN - 1) / BN) * (EPI == EPI_BIAS_F32 && P.split_k > 1 ? P.split_k : 1);
    int groups = max_ctas / CG;
    if (tiles < groups) groups = tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups * CG);
    cfg.blockDim = dim3(128 + 32 * epi_warps(EPI));
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = deco_pdl_enabled() ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, P);
    if (e != cudaSuccess) { deco_set_error("gemm launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return DECO_OK;
}

template <int BN, int CG, bool STAGED>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& P, int max_ctas, cudaStream_t st) {
    switch (epi) {
        case EPI_BIAS: return launch_gemm<BN, EPI_BIAS, CG, STAGED>(ta, tb, P, max_ctas, st);
        case EPI_BIAS_SILU: return launch_gemm<BN, EPI_BIAS_SILU, CG, STAGED>(ta, tb, P, max_ctas, st);
        case EPI_GATE_RESIDUAL: return launch_gemm<BN, EPI_GATE_RESIDUAL, CG, STAGED>(ta, tb, P, max_ctas, st);
        case EPI_SWIGLU: return launch_gemm<BN, EPI_SWIGLU, CG, STAGED>(ta, tb, P, max_ctas, st);
        case EPI_BIAS_F32: return launch_gemm<BN, EPI_BIAS_F32, CG, STAGED>(ta, tb, P, max_ctas, st);
        case EPI_SWIGLU_DUAL: return launch_gemm<BN, EPI_SWIGLU_DUAL, CG, false>(ta, tb, P, max_ctas, st);   // row-per-thread only
        case EPI_SWIGLU_BWD: return launch_gemm<BN, EPI_SWIGLU_BWD, CG, false>(ta, tb, P, max_ctas, st);
    }
    deco_set_error("gemm: unknown epilogue %d", epi);
    return DECO_ERR_ARG;
}

template <int BN>
static int dispatch_variant(int cg, int staged, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& P,
                            int max_ctas, cudaStream_t st) {
    if (cg == 2) return staged ? dispatch_epi<BN, 2, true>(epi, ta, tb, P, max_ctas, st)
                               : dispatch_epi<BN, 2, false>(epi, ta, tb, P, max_ctas, st);
    return staged ? dispatch_epi<BN, 1, true>(epi, ta, tb, P, max_ctas, st)
                  : dispatch_epi<BN, 1, false>(epi, ta, tb, P, max_ctas, st);
}

// tuning knobs (process-wide; -1 = automatic)
static int g_force_cta_group = -1;
static int g_force_staged = -1;

static int num_sms() { return device_sm_count() - deco_reserved_sms(); }

// Tile width of an M x N problem on `groups` CTA groups of cg CTAs: the smallest rounds x width / efficiency (see
// deco_gemm_bf16).  A ragged last column tile counts as its fraction of a tile when it can run a narrower MMA.
static int pick_tile_n(int M, int N, int cg, int groups) {
    const int num_m = (M + kBM * cg - 1) / (kBM * cg);
    int bn = 0;
    double best = 0.0;
    for (int cand : {256, 192, 128}) {
        if (cand == 192 && N % 192 != 0) continue;
        const int nfull = N / cand, rem = N - nfull * cand;
        const double frac = rem == 0 ? 0.0 : (rem % 32 == 0 ? (double)rem / cand : 1.0);
        const double units = (double)num_m * (nfull + frac);
        double rounds = units / groups;
        rounds = (rounds <= 1.0) ? 1.0 : (double)(long long)(rounds + 0.999);
        const double eff = cand == 256 ? 1.0 : (cand == 192 ? 0.78 : 0.55);
        const double cost = rounds * cand / eff;
        if (bn == 0 || cost < best) { bn = cand; best = cost; }
    }
    return bn;
}

}  // namespace deco

extern "C" int deco_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                              int M, int N, int K, int epilogue, const float* bias,
                              const void* resid, long long ldr, const void* gate, long long gate_stride,
                              int rows_per_gate, int tile_n, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(A && W && out, "gemm: null pointer");
    DECO_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    DECO_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0 && N % 8 == 0,
                   "gemm: K, N and leading dimensions must be multiples of 8 (K=%d N=%d lda=%lld ldw=%lld ldo=%lld)",
                   K, N, lda, ldw, ldo);
    DECO_CHECK_ARG((((uintptr_t)A | (uintptr_t)W | (uintptr_t)out) & 15) == 0, "gemm: pointers must be 16-byte aligned");
    if (epilogue == EPI_GATE_RESIDUAL)
        DECO_CHECK_ARG(resid && gate && rows_per_gate > 0 && ldr % 8 == 0 && gate_stride % 8 == 0 &&
                       (((uintptr_t)resid | (uintptr_t)gate) & 15) == 0, "gemm: gate/residual arguments invalid");
    if (epilogue == EPI_SWIGLU || epilogue == EPI_SWIGLU_DUAL) DECO_CHECK_ARG(N % 32 == 0, "gemm: swiglu epilogue needs N %% 32 == 0");
    if (epilogue == EPI_SWIGLU_DUAL || epilogue == EPI_SWIGLU_BWD)
        DECO_CHECK_ARG(resid && ldr % 8 == 0 && ((uintptr_t)resid & 15) == 0 && N % 16 == 0,
                       "gemm: the SwiGLU training epilogues take the bf16 pre-activation matrix in resid / ldr (N %% 16 == 0)");
    // 2-CTA pairs for anything with at least one full 256-row pair tile; staged (coalesced) epilogue only there
    int cg = (g_force_cta_group > 0) ? g_force_cta_group : (M > kBM ? 2 : 1);
    // Tile width.  Wide tiles feed the tensor pipe better (isolated: 850 / 1220 / 1590 TFLOP/s at 128 / 192 / 256 columns,
    // profiles/gemm_bench_r1.txt); a ragged last column tile runs a narrower MMA and costs its fraction of a tile; the
    // balanced schedule (TileSchedule) packs full and ragged tiles into ceil(units / CTA groups) rounds.  Pick the width with
    // the smallest rounds x width / efficiency: N = 1152 on 8192 rows -> 256 (2 rounds) instead of 192 (3 rounds).
    const int bn = tile_n ? tile_n : pick_tile_n(M, N, cg, num_sms() / cg);
    DECO_CHECK_ARG(bn == 128 || bn == 192 || bn == 256, "gemm: tile_n must be 128, 192 or 256");
    // staged (coalesced) epilogue everywhere except SwiGLU, whose output is half as wide as its accumulator tile and
    // is compute-heavy: the row-per-thread form keeps all 128 epilogue threads busy (1554 vs 1069 TFLOP/s at BN = 256)
    int staged = (g_force_staged >= 0) ? g_force_staged : (epilogue == EPI_SWIGLU ? 0 : 1);
    if (epilogue == EPI_SWIGLU_DUAL || epilogue == EPI_SWIGLU_BWD) staged = 0;
    CUtensorMap ta, tb;
    int rc = make_tmap(&ta, A, M, K, lda, kBM);
    if (rc) return rc;
    rc = make_tmap(&tb, W, N, K, ldw, bn / cg);
    if (rc) return rc;
    GemmParams P;
    P.out = out; P.ldo = ldo; P.bias = bias; P.resid = (const float*)resid; P.ldr = ldr;
    P.gate = (const __nv_bfloat16*)gate; P.gate_stride = gate_stride; P.rows_per_gate = rows_per_gate > 0 ? rows_per_gate : 1;
    P.M = M; P.N = N; P.K = K; P.split_k = 1; P.deint_rows = 0;
    const int ctas = num_sms();
    cudaStream_t st = (cudaStream_t)stream;
    if (bn == 256) return dispatch_variant<256>(cg, staged, epilogue, ta, tb, P, ctas, st);
    if (bn == 192) return dispatch_variant<192>(cg, staged, epilogue, ta, tb, P, ctas, st);
    return dispatch_variant<128>(cg, staged, epilogue, ta, tb, P, ctas, st);
}

// Host-side view of what deco_gemm_bf16 would launch for an M x N x K problem on `ctas` CTAs (0 = this device's SM count):
// the tile width it picks and, by walking the device code's own TileSchedule for every CTA group, the heaviest group's work in
// 1/256ths of a full-width tile (a ragged last-column tile that runs a narrower MMA counts n_last / tile_n of one) and the
// number of tiles nobody or more than one group visits (must be 0).  Needs no GPU: tests/test_abi.py checks the schedule.
extern "C" int deco_gemm_tile_plan(int M, int N, int K, int ctas, int* tile_n_out, int* max_load_256ths, int* bad_tiles)
{
    using namespace deco;
    DECO_CHECK_ARG(M > 0 && N > 0 && K > 0 && tile_n_out && max_load_256ths && bad_tiles, "gemm_tile_plan: bad arguments");
    const int cg = M > kBM ? 2 : 1;
    if (ctas <= 0) ctas = num_sms();
    const int groups = ctas / cg;
    DECO_CHECK_ARG(groups > 0, "gemm_tile_plan: no CTA group");
    const int bn = pick_tile_n(M, N, cg, groups);
    const int num_m = (M + kBM * cg - 1) / (kBM * cg), num_n = (N + bn - 1) / bn;
    const int n_rem = N - (num_n - 1) * bn;
    const int n_last = (n_rem < bn && n_rem % 32 == 0) ? n_rem : bn;
    const int used = num_m * num_n < groups ? num_m * num_n : groups;
    std::vector<int> visits((size_t)num_m * num_n, 0);
    long long worst = 0;
    for (int g = 0; g < used; ++g) {
        TileSchedule sched;
        sched.num_n = num_n; sched.num_mn = num_m * num_n; sched.num_tiles = num_m * num_n; sched.group = g; sched.groups = used;
        sched.full_mn = (n_last < bn && num_n > 1 && (long long)M * K * 2 <= (64ll << 20)) ? num_m * (num_n - 1) : 0;
        long long load = 0;
        for (int round = 0, tile; (tile = sched.tile_of_round(round)) >= 0; ++round) {
            int m_blk, n_blk, ks;
            sched.coords(tile, m_blk, n_blk, ks);
            if (m_blk < 0 || m_blk >= num_m || n_blk < 0 || n_blk >= num_n || ks != 0) { *bad_tiles = -1; return DECO_OK; }
            ++visits[(size_t)m_blk * num_n + n_blk];
            load += (n_blk == num_n - 1 ? n_last : bn) * 256 / bn;
        }
        if (load > worst) worst = load;
    }
    int bad = 0;
    for (int v : visits) bad += (v != 1);
    *tile_n_out = bn; *max_load_256ths = (int)worst; *bad_tiles = bad;
    return DECO_OK;
}

extern "C" int deco_gemm_set_tuning(int cta_group, int staged_epilogue) {
    using namespace deco;
    DECO_CHECK_ARG(cta_group == -1 || cta_group == 1 || cta_group == 2, "gemm tuning: cta_group must be -1, 1 or 2");
    DECO_CHECK_ARG(staged_epilogue >= -1 && staged_epilogue <= 1, "gemm tuning: staged_epilogue must be -1, 0 or 1");
    g_force_cta_group = cta_group;
    g_force_staged = staged_epilogue;
    return DECO_OK;
}

namespace deco {
// K slices for an fp32-output GEMM whose M x N tiles alone would leave most CTA groups idle: as many slices as fit in one
// wave, each at least 4 K-blocks (256 elements) deep.  0 / 1 = no split.
static int auto_split_k(int M, int N, int K, int bn, int cg, int ctas) {
    const int tiles = ((M + kBM * cg - 1) / (kBM * cg)) * ((N + bn - 1) / bn);
    const int groups = ctas / cg;
    const int num_k = (K + kBK - 1) / kBK;
    // Only when the tiles leave at least half of the CTA groups idle.  Splitting a GEMM that already fills a round to shave
    // its last round was measured and is slower (dW13 = 120 tiles: 3 slices = 5 rounds of a third, 95.7 -> 100.5 us): every
    // slice pays its own pipeline fill and a reduction epilogue.
    if (tiles * 2 > groups || num_k < 8) return 1;
    int s = groups / tiles;
    if (s > num_k / 4) s = num_k / 4;
    if (s > 32) s = 32;
    while (s > 1 && ((num_k + s - 1) / s) * (s - 1) >= num_k) --s;     // no empty slice
    return s < 1 ? 1 : s;
}

// largest s' <= s whose ceil(num_k / s')-deep slices are all non-empty
static int legal_split_k(int s, int K) {
    const int num_k = (K + kBK - 1) / kBK;
    if (s > num_k) s = num_k;
    while (s > 1 && ((num_k + s - 1) / s) * (s - 1) >= num_k) --s;
    return s < 1 ? 1 : s;
}

static int zero_output(float* out, long long ldo, int M, int N, cudaStream_t st) {
    cudaError_t e = (ldo == N) ? cudaMemsetAsync(out, 0, (size_t)M * N * 4, st)
                               : cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)N * 4, (size_t)M, st);
    if (e != cudaSuccess) { deco_set_error("gemm split-K: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    return DECO_OK;
}
}  // namespace deco

// out[M, N] fp32 = At^T . Wt with At [K, M] and Wt [K, N] row-major bf16 (the wgrad contraction dW = dY^T . X on the
// activations as they lie in memory).  M, N, lda, ldw, ldo multiples of 8.
static int gemm_tn_impl(const void* At, long long lda, const void* Wt, long long ldw, float* out, long long ldo,
                       int M, int N, int K, int tile_n, int split_k, int deinterleave16, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(!deinterleave16 || M % 32 == 0, "gemm_tn: de-interleaved output needs M %% 32 == 0 (M=%d)", M);
    DECO_CHECK_ARG(At && Wt && out, "gemm_tn: null pointer");
    DECO_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_tn: bad shape M=%d N=%d K=%d", M, N, K);
    DECO_CHECK_ARG(M % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 4 == 0,
                   "gemm_tn: M, N and leading dimensions must be multiples of 8 (M=%d N=%d lda=%lld ldw=%lld ldo=%lld)",
                   M, N, lda, ldw, ldo);
    DECO_CHECK_ARG((((uintptr_t)At | (uintptr_t)Wt | (uintptr_t)out) & 15) == 0, "gemm_tn: pointers must be 16-byte aligned");
    int bn = tile_n ? tile_n : (N > 128 ? 256 : 128);
    DECO_CHECK_ARG(bn == 128 || bn == 256, "gemm_tn: tile_n must be 128 or 256");
    const int cg = (g_force_cta_group > 0) ? g_force_cta_group : (M > kBM ? 2 : 1);
    CUtensorMap ta, tb;
    int rc = make_tmap(&ta, At, K, M, lda, kBK);      // {64 (MN) x 64 (K rows)} boxes
    if (rc) return rc;
    rc = make_tmap(&tb, Wt, K, N, ldw, kBK);
    if (rc) return rc;
    GemmParams P;
    P.out = out; P.ldo = ldo; P.bias = nullptr; P.resid = nullptr; P.ldr = 0;
    P.gate = nullptr; P.gate_stride = 0; P.rows_per_gate = 1;
    P.M = M; P.N = N; P.K = K; P.deint_rows = deinterleave16 ? M / 2 : 0;
    const int ctas = num_sms();
    cudaStream_t st = (cudaStream_t)stream;
    P.split_k = legal_split_k(split_k > 0 ? split_k : auto_split_k(M, N, K, bn, cg, ctas), K);
    if (P.split_k > 1) { rc = zero_output(out, ldo, M, N, st); if (rc) return rc; }
    if (bn == 256) return cg == 2 ? launch_gemm<256, EPI_BIAS_F32, 2, true, true>(ta, tb, P, ctas, st)
                                  : launch_gemm<256, EPI_BIAS_F32, 1, true, true>(ta, tb, P, ctas, st);
    return cg == 2 ? launch_gemm<128, EPI_BIAS_F32, 2, true, true>(ta, tb, P, ctas, st)
                   : launch_gemm<128, EPI_BIAS_F32, 1, true, true>(ta, tb, P, ctas, st);
}

extern "C" int deco_gemm_bf16_tn(const void* At, long long lda, const void* Wt, long long ldw, float* out, long long ldo,
                                 int M, int N, int K, int tile_n, int split_k, void* stream)
{
    return gemm_tn_impl(At, lda, Wt, ldw, out, ldo, M, N, K, tile_n, split_k, 0, stream);
}

// deco_gemm_bf16_tn whose M output rows are the [16 x w1 | 16 x w3] interleaved rows of the SwiGLU weight: they are stored
// de-interleaved, out = [2][M / 2][ldo] = (dW1 ; dW3), so that no copy has to pull the two gradients apart afterwards.
extern "C" int deco_gemm_bf16_tn_deint16(const void* At, long long lda, const void* Wt, long long ldw, float* out, long long ldo,
                                         int M, int N, int K, int tile_n, int split_k, void* stream)
{
    return gemm_tn_impl(At, lda, Wt, ldw, out, ldo, M, N, K, tile_n, split_k, 1, stream);
}

// out[M, N] fp32 = A . W^T (K-major operands as deco_gemm_bf16 with DECO_EPI_BIAS_F32 and no bias) with the K loop split
// over split_k tiles (0 = automatic) and reduced with fp32 atomics: for skinny problems whose M x N tiles cannot fill the
// GPU, e.g. d c = d mod . Wada with M = batch, K = 6 x hidden x blocks.
extern "C" int deco_gemm_bf16_f32_splitk(const void* A, long long lda, const void* W, long long ldw, float* out, long long ldo,
                                         int M, int N, int K, int split_k, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(A && W && out, "gemm_splitk: null pointer");
    DECO_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 4 == 0,
                   "gemm_splitk: bad shape / alignment (M=%d N=%d K=%d)", M, N, K);
    DECO_CHECK_ARG((((uintptr_t)A | (uintptr_t)W | (uintptr_t)out) & 15) == 0, "gemm_splitk: pointers must be 16-byte aligned");
    const int bn = (N % 256 == 0) ? 256 : ((N % 192 == 0) ? 192 : 128);
    const int cg = M > kBM ? 2 : 1;
    CUtensorMap ta, tb;
    int rc = make_tmap(&ta, A, M, K, lda, kBM);
    if (rc) return rc;
    rc = make_tmap(&tb, W, N, K, ldw, bn / cg);
    if (rc) return rc;
    GemmParams P;
    P.out = out; P.ldo = ldo; P.bias = nullptr; P.resid = nullptr; P.ldr = 0;
    P.gate = nullptr; P.gate_stride = 0; P.rows_per_gate = 1;
    P.M = M; P.N = N; P.K = K; P.deint_rows = 0;
    const int ctas = num_sms();
    cudaStream_t st = (cudaStream_t)stream;
    P.split_k = legal_split_k(split_k > 0 ? split_k : auto_split_k(M, N, K, bn, cg, ctas), K);
    if (P.split_k > 1) { rc = zero_output(out, ldo, M, N, st); if (rc) return rc; }
    if (bn == 256) return dispatch_variant<256>(cg, 1, EPI_BIAS_F32, ta, tb, P, ctas, st);
    if (bn == 192) return dispatch_variant<192>(cg, 1, EPI_BIAS_F32, ta, tb, P, ctas, st);
    return dispatch_variant<128>(cg, 1, EPI_BIAS_F32, ta, tb, P, ctas, st);
}
