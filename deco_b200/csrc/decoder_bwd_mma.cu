// Backward of the per-pixel AdaLN-MLP pixel decoder on warp-level tensor-core MMAs (training step).
// Same contract as csrc/decoder_bwd.cu (the scalar fp32 version this one is validated against): inputs x, ycond, dout;
// outputs dycond (bf16) and the gradients of every decoder parameter, in the fp32 blob layout of decoder_bwd.cu.
//
// Reference ops (/root/reference/src/models/transformer/dit_c2i_DeCo.py): :212-248 NerfEmbedder, :313-317 ResBlock.forward,
// :329-332 decoder FinalLayer, :395-415 SimpleMLPAdaLN.forward -- differentiated by hand.
//
// Layout (as the forward, csrc/decoder.cu): one warp = the 16 pixels of one patch row; lane (g = lane/4, t = lane%4) owns
// pixel rows g, g+8 and channels {8j+2t, 8j+2t+1}.  A 16 x 32 activation is kept as eight packed 8x8 bf16 tiles
// (lo[j] = pixels 0-7, hi[j] = pixels 8-15 of channel tile j): {lo[2s], hi[2s], lo[2s+1], hi[2s+1]} IS the m16n8k16 A
// fragment of k-step s, and `movmatrix.trans` of a tile gives the fragments of the transposed operand, so
//   dgrad  dX = dY . W      : A = dY tiles,            B = W^T packed on the host in fragment order
//   wgrad  dW = dY^T . X    : A = movmatrix(dY tiles), B = movmatrix(X tiles), K = the warp's 16 pixels; the bias
//                              gradient rides along as one more n-tile whose B fragment is a column of ones
// The forward of a res-block is recomputed from its saved input (kept per warp in shared memory); weight gradients are
// accumulated per CTA in shared memory by OWNER warps (the warps exchange operand tiles through a staging buffer instead
// of atomically adding results, see wgrad_coop) and flushed once; a warp keeps ONE patch row ky for its whole life so the
// positional-table gradient accumulates in 16 registers.
#include "common.cuh"

namespace deco {

constexpr int kFW = 64;                      // uint32 words per (n-tile, k-step) fragment block
constexpr int kF32 = 4 * 2 * kFW;            // 32 -> 32 layer (4 n-tiles x 2 k-steps)
constexpr int kF96 = 12 * 2 * kFW;           // forward 32 -> 96 == backward 96 -> 32 (4 n-tiles x 6 k-steps)
constexpr int kF8 = 1 * 2 * kFW;
constexpr int kFBlock = kF96 + 2 * kF32;
// forward blob (decoder.cu): frag words [Win | R x (Wada, W0, W2) | Wf8], then floats rgb[32][3] bin[32] R x (bada96 lng lnb b0 b2) bf[8]
constexpr int kVRgb = 0, kVBin = 96, kVBlock0 = 128, kVPerBlock = 224;
__host__ __device__ inline int mf_frag_words(int R) { return kF32 + R * kFBlock + kF8; }
__host__ __device__ inline int mf_vec_floats(int R) { return kVBlock0 + R * kVPerBlock + 8; }
// backward blob: frag words [WinT | R x (WadaT (n-permuted), W0T, W2T)], then floats Wf[3][32]
__host__ __device__ inline int mb_frag_words(int R) { return kF32 + R * kFBlock; }
__host__ __device__ inline int mb_words(int R) { return mb_frag_words(R) + 96; }
// shared gradient accumulators (floats; matrices with row stride 36)
constexpr int kS = 36;
constexpr int kGWin = 0, kGbin = 32 * kS, kGBlock0 = kGbin + 32;
constexpr int kHWada = 0, kHbada = 96 * kS, kHlng = kHbada + 96, kHlnb = kHlng + 32, kHW0 = kHlnb + 32, kHb0 = kHW0 + 32 * kS,
              kHW2 = kHb0 + 32, kHb2 = kHW2 + 32 * kS, kGBlock = kHb2 + 32;
__host__ __device__ inline int mg_final(int R) { return kGBlock0 + R * kGBlock; }          // Wf [3][36], bf[4], Wrgb [32][4]
__host__ __device__ inline int mg_floats(int R) { return mg_final(R) + 3 * kS + 4 + 128; }
// global gradient layout = fp32 blob of decoder_bwd.cu
constexpr int kTWrgb_ = 0, kTWin_ = 96, kTbin_ = 1120, kTBlock0_ = 1152, kTBlock_ = 5344;
constexpr int kBWada_ = 0, kBbada_ = 3072, kBlng_ = 3168, kBlnb_ = 3200, kBW0_ = 3232, kBb0_ = 4256, kBW2_ = 4288, kBb2_ = 5312;

struct Tiles { uint32_t lo[4], hi[4]; };     // 16 x 32 bf16 as 8x8 tiles

__device__ __forceinline__ uint32_t movm(uint32_t v) {
    uint32_t r;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ Tiles pack_tiles(const float (&v)[4][4]) {
    Tiles p;
#pragma unroll
    for (int j = 0; j < 4; ++j) { p.lo[j] = pack_bf2(v[j][0], v[j][1]); p.hi[j] = pack_bf2(v[j][2], v[j][3]); }
    return p;
}
__device__ __forceinline__ Tiles transpose_tiles(const Tiles& p) {
    Tiles q;
#pragma unroll
    for (int j = 0; j < 4; ++j) { q.lo[j] = movm(p.lo[j]); q.hi[j] = movm(p.hi[j]); }
    return q;
}
// acc[j] (+)= A (tiles, KS = 2) . B for n-tiles [j0, j0 + NT) of a fragment-packed weight with KS k-steps per n-tile
template <int NT>
__device__ __forceinline__ void mma_t(float (&acc)[NT][4], const Tiles& a, const uint32_t* sW, int j0, int lane) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const uint2 b = *reinterpret_cast<const uint2*>(sW + ((j0 + j) * 2 + s) * kFW + lane * 2);
            const uint32_t af[4] = {a.lo[2 * s], a.hi[2 * s], a.lo[2 * s + 1], a.hi[2 * s + 1]};
            const uint32_t bb[2] = {b.x, b.y};
            mma_bf16_16816(acc[j], af, bb);
        }
    }
}
__device__ __forceinline__ void quad_sum2m(float& a, float& b) {
    a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2);
}
__device__ __forceinline__ void ln_stats_m(const float (&x)[4][4], float& m0, float& r0, float& m1, float& r1) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { s0 += x[j][0] + x[j][1]; s1 += x[j][2] + x[j][3]; }
    quad_sum2m(s0, s1);
    m0 = s0 * (1.0f / 32); m1 = s1 * (1.0f / 32);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float d;
        d = x[j][0] - m0; q0 = fmaf(d, d, q0); d = x[j][1] - m0; q0 = fmaf(d, d, q0);
        d = x[j][2] - m1; q1 = fmaf(d, d, q1); d = x[j][3] - m1; q1 = fmaf(d, d, q1);
    }
    quad_sum2m(q0, q1);
    r0 = rsqrtf(q0 * (1.0f / 32) + 1e-6f);
    r1 = rsqrtf(q1 * (1.0f / 32) + 1e-6f);
}
// dh (+)= LayerNorm backward: r * (dhn - mean(dhn) - hn * mean(dhn * hn)) per pixel row
template <bool SET>
__device__ __forceinline__ void ln_bwd_m(const float (&dhn)[4][4], const float (&hn)[4][4], float r0, float r1, float (&dh)[4][4]) {
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s0 += dhn[j][0] + dhn[j][1]; s1 += dhn[j][2] + dhn[j][3];
        q0 = fmaf(dhn[j][0], hn[j][0], fmaf(dhn[j][1], hn[j][1], q0));
        q1 = fmaf(dhn[j][2], hn[j][2], fmaf(dhn[j][3], hn[j][3], q1));
    }
    quad_sum2m(s0, s1); quad_sum2m(q0, q1);
    s0 *= (1.0f / 32); s1 *= (1.0f / 32); q0 *= (1.0f / 32); q1 *= (1.0f / 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float v0 = r0 * (dhn[j][0] - s0 - hn[j][0] * q0), v1 = r0 * (dhn[j][1] - s0 - hn[j][1] * q0);
        const float v2 = r1 * (dhn[j][2] - s1 - hn[j][2] * q1), v3 = r1 * (dhn[j][3] - s1 - hn[j][3] * q1);
        if (SET) { dh[j][0] = v0; dh[j][1] = v1; dh[j][2] = v2; dh[j][3] = v3; }
        else { dh[j][0] += v0; dh[j][1] += v1; dh[j][2] += v2; dh[j][3] += v3; }
    }
}
// Cooperative wgrad of one 32-row slab: gW[(row0 + i) * 36 + col] += sum over the CTA's 8 x 16 pixels of dY[p][i] X[p][col],
// gb[row0 + i] += sum_p dY[p][i].  Shared-memory fp32 atomicAdd is a compare-and-swap loop on this architecture (80 % of
// the first version's stall samples), so instead of every warp adding its own 16-pixel products into the CTA's
// accumulators, the warps EXCHANGE OPERANDS: each stages its transposed tiles (8 + 8 words per lane), and after one barrier
// warp w -- the exclusive owner of output tile (m-tile w / 4, n-tile w % 4) -- runs the 8 MMAs of all eight warps' pixels
// for that tile and adds the result to the accumulator it owns with plain loads and stores.  The bias gradient is a ninth
// n-tile whose B fragment is a column of ones (owners: the n-tile 0 warps).  Two staging buffers alternate (parity), so one
// barrier per call suffices.  PERM: n-tile q, column c <-> channel 8 (c / 2) + 2 q + (c % 2) (the adaLN input layout).
// All 8 warps of a CTA run the same number of items (they share the token index), so the barriers are uniform.
constexpr int kStageWords = 8 * 16 * 32;     // one staging buffer: [warp][16 words][lane]
template <bool PERM>
__device__ __forceinline__ void wgrad_coop(uint32_t* stage, int& parity, float* gW, float* gb, int row0,
                                           const Tiles& dyT, const Tiles& xT, int warp, int lane) {
    uint32_t* buf = stage + parity * kStageWords;
    parity ^= 1;
    uint32_t* my = buf + warp * 16 * 32 + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        my[j * 32] = dyT.lo[j]; my[(4 + j) * 32] = dyT.hi[j];
        my[(8 + j) * 32] = xT.lo[j]; my[(12 + j) * 32] = xT.hi[j];
    }
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    const int mt = warp >> 2, nt = warp & 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f}, cb[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t one = g == 0 ? 0x3F803F80u : 0u;          // bf16 (1, 1) in column n = 0
    const uint32_t ones[2] = {one, one};
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t* src = buf + w * 16 * 32 + lane;
        const uint32_t af[4] = {src[(2 * mt) * 32], src[(2 * mt + 1) * 32], src[(4 + 2 * mt) * 32], src[(4 + 2 * mt + 1) * 32]};
        const uint32_t bb[2] = {src[(8 + nt) * 32], src[(12 + nt) * 32]};
        mma_bf16_16816(c, af, bb);
        if (nt == 0) mma_bf16_16816(cb, af, ones);
    }
    const int col = PERM ? 8 * t + 2 * nt : 8 * nt + 2 * t;
    float* p0 = gW + (row0 + 16 * mt + g) * kS + col;
    p0[0] += c[0]; p0[1] += c[1]; p0[8 * kS] += c[2]; p0[8 * kS + 1] += c[3];
    if (nt == 0 && t == 0) { gb[row0 + 16 * mt + g] += cb[0]; gb[row0 + 16 * mt + g + 8] += cb[2]; }
}
// Column sums over the CTA's pixels of two 16 x 32 tensors (LayerNorm weight / bias gradients) by the same exchange:
// v1 tiles travel in the A slots (owners: warps 0 and 4), v2 tiles in the B slots (owners: warps 1 and 5).
__device__ __forceinline__ void vecsum_coop(uint32_t* stage, int& parity, float* g1, float* g2,
                                            const Tiles& v1T, const Tiles& v2T, int warp, int lane) {
    uint32_t* buf = stage + parity * kStageWords;
    parity ^= 1;
    uint32_t* my = buf + warp * 16 * 32 + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        my[j * 32] = v1T.lo[j]; my[(4 + j) * 32] = v1T.hi[j];
        my[(8 + j) * 32] = v2T.lo[j]; my[(12 + j) * 32] = v2T.hi[j];
    }
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    const int mt = warp >> 2, which = warp & 3;
    if (which < 2) {
        float cb[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t one = g == 0 ? 0x3F803F80u : 0u;
        const uint32_t ones[2] = {one, one};
        const int o = which * 8;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t* src = buf + w * 16 * 32 + lane;
            const uint32_t af[4] = {src[(o + 2 * mt) * 32], src[(o + 2 * mt + 1) * 32], src[(o + 4 + 2 * mt) * 32], src[(o + 4 + 2 * mt + 1) * 32]};
            mma_bf16_16816(cb, af, ones);
        }
        float* gv = which == 0 ? g1 : g2;
        if (t == 0) { gv[16 * mt + g] += cb[0]; gv[16 * mt + g + 8] += cb[2]; }
    }
}
__device__ __forceinline__ float dsilu_m(float x) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-x));
    return s * fmaf(x, 1.0f - s, 1.0f);
}

struct DecBwdMmaParams {
    const float* x; const __nv_bfloat16* ycond; const float* dout;
    const uint32_t* fblob;       // forward blob (decoder.cu layout)
    const uint32_t* bblob;       // backward blob
    const float* postab;
    __nv_bfloat16* dycond;
    float* gblob;                // fp32 gradient accumulators, layout of decoder_bwd.cu (+ d postab)
    int R, H, W, Hp, Wp;
    long long M;
};

__global__ void __launch_bounds__(256, 1) pixel_decoder_bwd_mma_kernel(DecBwdMmaParams P)
{
    extern __shared__ __align__(16) uint32_t msm[];
    const int R = P.R;
    const int nff = mf_frag_words(R), nfv = mf_vec_floats(R), nbw = mb_words(R), ng = mg_floats(R);
    uint32_t* sF = msm;                                   // forward frags + vec
    uint32_t* sBw = sF + ((nff + nfv + 3) & ~3);          // backward frags + Wf
    float* sG = reinterpret_cast<float*>(sBw + ((nbw + 3) & ~3));
    float* sH = sG + ((ng + 3) & ~3);                     // block inputs: [warp][R][16][32]
    uint32_t* sStage = reinterpret_cast<uint32_t*>(sH + 8 * R * 16 * 32);   // operand exchange, 2 x kStageWords
    int parity = 0;
    for (int i = threadIdx.x; i < nff + nfv; i += blockDim.x) sF[i] = __ldg(P.fblob + i);
    for (int i = threadIdx.x; i < nbw; i += blockDim.x) sBw[i] = __ldg(P.bblob + i);
    for (int i = threadIdx.x; i < ng; i += blockDim.x) sG[i] = 0.f;
    __syncthreads();
    const uint32_t* sW = sF;
    const float* sV = reinterpret_cast<const float*>(sF + nff);
    const uint32_t* sWT = sBw;
    const float* sWf = reinterpret_cast<const float*>(sBw + mb_frag_words(R));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    float* hsave = sH + (size_t)warp * R * 16 * 32;
    const int L = P.Hp * P.Wp;
    const long long gw = (long long)blockIdx.x * 8 + warp;       // global warp id; total warps is a multiple of 16
    const long long nwarps = (long long)gridDim.x * 8;
    const int ky = (int)(gw & 15);
    const size_t plane = (size_t)P.H * P.W;
    float dpos[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { dpos[j][0] = 0.f; dpos[j][1] = 0.f; dpos[j][2] = 0.f; dpos[j][3] = 0.f; }

    for (long long m = gw >> 4; m < P.M; m += nwarps >> 4) {
        const long long b = m / L;
        const int tok = (int)(m % L);
        const int py = tok / P.Wp, px = tok % P.Wp;
        const __nv_bfloat16* yrow = P.ycond + (m * 256 + ky * 16) * 32;
        const uint4 y0 = ld_stream16(yrow + g * 32 + 8 * t);
        const uint4 y1 = ld_stream16(yrow + (g + 8) * 32 + 8 * t);
        const size_t pix0 = (size_t)(py * 16 + ky) * P.W + px * 16;
        const float* xb = P.x + (size_t)b * 3 * plane + pix0;
        const float* db = P.dout + (size_t)b * 3 * plane + pix0;
        float rgb[2][3], dout[2][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            rgb[0][c] = round_bf(__ldg(xb + c * plane + g)); rgb[1][c] = round_bf(__ldg(xb + c * plane + g + 8));
            dout[0][c] = __ldg(db + c * plane + g); dout[1][c] = __ldg(db + c * plane + g + 8);
        }
        // silu(y): A fragments (K-permuted) of the adaLN layers, and its transposed tiles for the adaLN wgrad
        const uint32_t yw0[4] = {y0.x, y0.y, y0.z, y0.w}, yw1[4] = {y1.x, y1.y, y1.z, y1.w};
        Tiles ys;                       // lo[q] = silu of word q (pixel g), hi[q] = (pixel g + 8): channels 8t+2q, 8t+2q+1
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 a = unpack_bf2(yw0[q]), c = unpack_bf2(yw1[q]);
            ys.lo[q] = pack_bf2(silu_f(a.x), silu_f(a.y));
            ys.hi[q] = pack_bf2(silu_f(c.x), silu_f(c.y));
        }
        // As an A operand the chunk words are (k-step s = q / 2, half = q % 2): {lo[2s], hi[2s], lo[2s+1], hi[2s+1]} -- the
        // same rule as Tiles, so `ys` is passed to mma_t directly (the forward packs Wada with the matching K permutation).
        const Tiles ysT = transpose_tiles(ys);

        // ---------------- forward recompute
        float xr[4][4];
        Tiles e0;
        {
            const float* T0 = P.postab + (size_t)(ky * 16 + g) * 32;
            const float* T1 = T0 + 8 * 32;
            float x0[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = 8 * j + 2 * t;
                const float2 ta = __ldg(reinterpret_cast<const float2*>(T0 + col));
                const float2 tb = __ldg(reinterpret_cast<const float2*>(T1 + col));
                const float* wa = sV + kVRgb + col * 3;
                x0[j][0] = ta.x + wa[0] * rgb[0][0] + wa[1] * rgb[0][1] + wa[2] * rgb[0][2];
                x0[j][1] = ta.y + wa[3] * rgb[0][0] + wa[4] * rgb[0][1] + wa[5] * rgb[0][2];
                x0[j][2] = tb.x + wa[0] * rgb[1][0] + wa[1] * rgb[1][1] + wa[2] * rgb[1][2];
                x0[j][3] = tb.y + wa[3] * rgb[1][0] + wa[4] * rgb[1][1] + wa[5] * rgb[1][2];
            }
            e0 = pack_tiles(x0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(sV + kVBin + 8 * j + 2 * t);
                xr[j][0] = bv.x; xr[j][1] = bv.y; xr[j][2] = bv.x; xr[j][3] = bv.y;
            }
            mma_t<4>(xr, e0, sW, 0, lane);
        }
#pragma unroll 1
        for (int rb = 0; rb < R; ++rb) {
            const uint32_t* wAda = sW + kF32 + rb * kFBlock;
            const uint32_t* w0 = wAda + kF96;
            const uint32_t* w2 = w0 + kF32;
            const float* vB = sV + kVBlock0 + rb * kVPerBlock;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) hsave[((rb * 16) + j * 4 + e) * 32 + lane] = xr[j][e];
            float ss[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 8 * j + 2 * t);
                ss[j][0] = bv.x; ss[j][1] = bv.y; ss[j][2] = bv.x; ss[j][3] = bv.y;
            }
            mma_t<8>(ss, ys, wAda, 0, lane);
            float m0, r0, m1, r1;
            ln_stats_m(xr, m0, r0, m1, r1);
            float h[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 gm = *reinterpret_cast<const float2*>(vB + 96 + 8 * j + 2 * t);
                const float2 bt = *reinterpret_cast<const float2*>(vB + 128 + 8 * j + 2 * t);
                h[j][0] = fmaf(fmaf((xr[j][0] - m0) * r0, gm.x, bt.x), 1.0f + ss[4 + j][0], ss[j][0]);
                h[j][1] = fmaf(fmaf((xr[j][1] - m0) * r0, gm.y, bt.y), 1.0f + ss[4 + j][1], ss[j][1]);
                h[j][2] = fmaf(fmaf((xr[j][2] - m1) * r1, gm.x, bt.x), 1.0f + ss[4 + j][2], ss[j][2]);
                h[j][3] = fmaf(fmaf((xr[j][3] - m1) * r1, gm.y, bt.y), 1.0f + ss[4 + j][3], ss[j][3]);
            }
            Tiles ah = pack_tiles(h);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 160 + 8 * j + 2 * t);
                h[j][0] = bv.x; h[j][1] = bv.y; h[j][2] = bv.x; h[j][3] = bv.y;
            }
            mma_t<4>(h, ah, w0, 0, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) h[j][e] = silu_f(h[j][e]);
            ah = pack_tiles(h);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 192 + 8 * j + 2 * t);
                h[j][0] = bv.x; h[j][1] = bv.y; h[j][2] = bv.x; h[j][3] = bv.y;
            }
            mma_t<4>(h, ah, w2, 0, lane);
            float gt[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 64 + 8 * j + 2 * t);
                gt[j][0] = bv.x; gt[j][1] = bv.y; gt[j][2] = bv.x; gt[j][3] = bv.y;
            }
            mma_t<4>(gt, ys, wAda, 8, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) xr[j][e] = fmaf(gt[j][e], h[j][e], xr[j][e]);
        }

        // ---------------- backward: final LayerNorm (no affine) + Linear 32 -> 3
        float dh[4][4];
        {
            float m0, r0, m1, r1;
            ln_stats_m(xr, m0, r0, m1, r1);
            float hf[4][4], dhf[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hf[j][0] = (xr[j][0] - m0) * r0; hf[j][1] = (xr[j][1] - m0) * r0;
                hf[j][2] = (xr[j][2] - m1) * r1; hf[j][3] = (xr[j][3] - m1) * r1;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int ch = 8 * j + 2 * t + (e & 1), pr = e >> 1;
                    dhf[j][e] = dout[pr][0] * sWf[ch] + dout[pr][1] * sWf[32 + ch] + dout[pr][2] * sWf[64 + ch];
                }
            }
            // dWf[c][ch] += sum_p dout[p][c] hf[p][ch]: A fragment rows = output channel c (< 3), built from global loads
            uint32_t af[4] = {0u, 0u, 0u, 0u};
            if (g < 3) {
                const float2 d0 = __ldg(reinterpret_cast<const float2*>(db + g * plane + 2 * t));
                const float2 d1 = __ldg(reinterpret_cast<const float2*>(db + g * plane + 8 + 2 * t));
                af[0] = pack_bf2(d0.x, d0.y); af[2] = pack_bf2(d1.x, d1.y);
            }
            const Tiles hfT = transpose_tiles(pack_tiles(hf));
            float* gf = sG + mg_final(R);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t bb[2] = {hfT.lo[nt], hfT.hi[nt]};
                mma_bf16_16816(c, af, bb);
                if (g < 3) { atomicAdd(gf + g * kS + 8 * nt + 2 * t, c[0]); atomicAdd(gf + g * kS + 8 * nt + 2 * t + 1, c[1]); }
            }
            {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                const uint32_t one = g == 0 ? 0x3F803F80u : 0u;
                const uint32_t bb[2] = {one, one};
                mma_bf16_16816(c, af, bb);
                if (t == 0 && g < 3) atomicAdd(gf + 3 * kS + g, c[0]);
            }
            ln_bwd_m<true>(dhf, hf, r0, r1, dh);
        }

        float dys[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { dys[j][0] = 0.f; dys[j][1] = 0.f; dys[j][2] = 0.f; dys[j][3] = 0.f; }

        // ---------------- backward through the res-blocks
#pragma unroll 1
        for (int rb = R - 1; rb >= 0; --rb) {
            const uint32_t* wAda = sW + kF32 + rb * kFBlock;
            const uint32_t* w0 = wAda + kF96;
            const uint32_t* w2 = w0 + kF32;
            const uint32_t* wAdaT = sWT + kF32 + rb * kFBlock;
            const uint32_t* w0T = wAdaT + kF96;
            const uint32_t* w2T = w0T + kF32;
            const float* vB = sV + kVBlock0 + rb * kVPerBlock;
            float* gB = sG + kGBlock0 + rb * kGBlock;
            float xin[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) xin[j][e] = hsave[((rb * 16) + j * 4 + e) * 32 + lane];
            // recompute the block forward
            float ss[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 8 * j + 2 * t);
                ss[j][0] = bv.x; ss[j][1] = bv.y; ss[j][2] = bv.x; ss[j][3] = bv.y;
            }
            mma_t<8>(ss, ys, wAda, 0, lane);
            float m0, r0, m1, r1;
            ln_stats_m(xin, m0, r0, m1, r1);
            float hn[4][4], h[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 gm = *reinterpret_cast<const float2*>(vB + 96 + 8 * j + 2 * t);
                const float2 bt = *reinterpret_cast<const float2*>(vB + 128 + 8 * j + 2 * t);
                hn[j][0] = (xin[j][0] - m0) * r0; hn[j][1] = (xin[j][1] - m0) * r0;
                hn[j][2] = (xin[j][2] - m1) * r1; hn[j][3] = (xin[j][3] - m1) * r1;
                h[j][0] = fmaf(fmaf(hn[j][0], gm.x, bt.x), 1.0f + ss[4 + j][0], ss[j][0]);
                h[j][1] = fmaf(fmaf(hn[j][1], gm.y, bt.y), 1.0f + ss[4 + j][1], ss[j][1]);
                h[j][2] = fmaf(fmaf(hn[j][2], gm.x, bt.x), 1.0f + ss[4 + j][2], ss[j][2]);
                h[j][3] = fmaf(fmaf(hn[j][3], gm.y, bt.y), 1.0f + ss[4 + j][3], ss[j][3]);
            }
            const Tiles hmP = pack_tiles(h);
            float z[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 160 + 8 * j + 2 * t);
                z[j][0] = bv.x; z[j][1] = bv.y; z[j][2] = bv.x; z[j][3] = bv.y;
            }
            mma_t<4>(z, hmP, w0, 0, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) h[j][e] = silu_f(z[j][e]);
            const Tiles actP = pack_tiles(h);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 192 + 8 * j + 2 * t);
                h[j][0] = bv.x; h[j][1] = bv.y; h[j][2] = bv.x; h[j][3] = bv.y;
            }
            mma_t<4>(h, actP, w2, 0, lane);                 // h = mm
            float gt[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 64 + 8 * j + 2 * t);
                gt[j][0] = bv.x; gt[j][1] = bv.y; gt[j][2] = bv.x; gt[j][3] = bv.y;
            }
            mma_t<4>(gt, ys, wAda, 8, lane);

            // h_out = h_in + gate * mm
            Tiles dmodP[3];                                 // d shift, d scale, d gate
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float d = dh[j][e]; gt[j][e] *= d; h[j][e] *= d; }   // gt = d mm, h = d gate
            dmodP[2] = pack_tiles(h);
            const Tiles dmmP = pack_tiles(gt);
            wgrad_coop<false>(sStage, parity, gB + kHW2, gB + kHb2, 0, transpose_tiles(dmmP), transpose_tiles(actP), warp, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) { h[j][0] = 0.f; h[j][1] = 0.f; h[j][2] = 0.f; h[j][3] = 0.f; }
            mma_t<4>(h, dmmP, w2T, 0, lane);                // h = d act
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) h[j][e] *= dsilu_m(z[j][e]);                                  // h = d z
            const Tiles dzP = pack_tiles(h);
            wgrad_coop<false>(sStage, parity, gB + kHW0, gB + kHb0, 0, transpose_tiles(dzP), transpose_tiles(hmP), warp, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) { h[j][0] = 0.f; h[j][1] = 0.f; h[j][2] = 0.f; h[j][3] = 0.f; }
            mma_t<4>(h, dzP, w0T, 0, lane);                 // h = d hm
            dmodP[0] = pack_tiles(h);
            {
                float tmp[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 gm = *reinterpret_cast<const float2*>(vB + 96 + 8 * j + 2 * t);
                    const float2 bt = *reinterpret_cast<const float2*>(vB + 128 + 8 * j + 2 * t);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float gmv = (e & 1) ? gm.y : gm.x, btv = (e & 1) ? bt.y : bt.x;
                        tmp[j][e] = h[j][e] * fmaf(hn[j][e], gmv, btv);          // d scale = d hm * hl
                    }
                }
                dmodP[1] = pack_tiles(tmp);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) { h[j][e] *= 1.0f + ss[4 + j][e]; tmp[j][e] = h[j][e] * hn[j][e]; }   // h = d hl
                vecsum_coop(sStage, parity, gB + kHlng, gB + kHlnb, transpose_tiles(pack_tiles(tmp)),
                            transpose_tiles(pack_tiles(h)), warp, lane);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 gm = *reinterpret_cast<const float2*>(vB + 96 + 8 * j + 2 * t);
                    h[j][0] *= gm.x; h[j][1] *= gm.y; h[j][2] *= gm.x; h[j][3] *= gm.y;               // h = d hn
                }
            }
            ln_bwd_m<false>(h, hn, r0, r1, dh);
            // adaLN: mod = Wada . silu(y) + bada
#pragma unroll
            for (int part = 0; part < 3; ++part)
                wgrad_coop<true>(sStage, parity, gB + kHWada, gB + kHbada, 32 * part, transpose_tiles(dmodP[part]), ysT, warp, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int s = 0; s < 6; ++s) {
                    const Tiles& a = dmodP[s >> 1];
                    const int sl = s & 1;
                    const uint2 bq = *reinterpret_cast<const uint2*>(wAdaT + (j * 6 + s) * kFW + lane * 2);
                    const uint32_t af[4] = {a.lo[2 * sl], a.hi[2 * sl], a.lo[2 * sl + 1], a.hi[2 * sl + 1]};
                    const uint32_t bb[2] = {bq.x, bq.y};
                    mma_bf16_16816(dys[j], af, bb);
                }
            }
        }

        // ---------------- input_proj and NerfEmbedder
        {
            const Tiles dhP = pack_tiles(dh);
            const Tiles dhT = transpose_tiles(dhP);
            wgrad_coop<false>(sStage, parity, sG + kGWin, sG + kGbin, 0, dhT, transpose_tiles(e0), warp, lane);
            float de[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { de[j][0] = 0.f; de[j][1] = 0.f; de[j][2] = 0.f; de[j][3] = 0.f; }
            mma_t<4>(de, dhP, sWT, 0, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) dpos[j][e] += de[j][e];
            // dWrgb[i][c] += sum_p de[p][i] rgb[p][c]: B fragment n = colour c (< 3) from global loads
            uint32_t bb[2] = {0u, 0u};
            if (g < 3) {
                const float2 c0 = __ldg(reinterpret_cast<const float2*>(xb + g * plane + 2 * t));
                const float2 c1 = __ldg(reinterpret_cast<const float2*>(xb + g * plane + 8 + 2 * t));
                bb[0] = pack_bf2(c0.x, c0.y); bb[1] = pack_bf2(c1.x, c1.y);
            }
            const Tiles deT = transpose_tiles(pack_tiles(de));
            float* gr = sG + mg_final(R) + 3 * kS + 4;       // [32][4]
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const uint32_t af[4] = {deT.lo[2 * mt], deT.lo[2 * mt + 1], deT.hi[2 * mt], deT.hi[2 * mt + 1]};
                float c[4] = {0.f, 0.f, 0.f, 0.f};
                mma_bf16_16816(c, af, bb);
                // columns 2t, 2t+1 = colours; only t == 0 (c 0, 1) and t == 1 (c 2) are real
                if (t == 0) {
                    atomicAdd(gr + (16 * mt + g) * 4, c[0]); atomicAdd(gr + (16 * mt + g) * 4 + 1, c[1]);
                    atomicAdd(gr + (16 * mt + g + 8) * 4, c[2]); atomicAdd(gr + (16 * mt + g + 8) * 4 + 1, c[3]);
                } else if (t == 1) {
                    atomicAdd(gr + (16 * mt + g) * 4 + 2, c[0]); atomicAdd(gr + (16 * mt + g + 8) * 4 + 2, c[2]);
                }
            }
        }

        // ---------------- d ycond = d silu(y) * silu'(y): dys[q] columns are channels 8t+2q, 8t+2q+1 = word q of the chunk
        {
            uint32_t o0[4], o1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 a = unpack_bf2(yw0[q]), c = unpack_bf2(yw1[q]);
                o0[q] = pack_bf2(dys[q][0] * dsilu_m(a.x), dys[q][1] * dsilu_m(a.y));
                o1[q] = pack_bf2(dys[q][2] * dsilu_m(c.x), dys[q][3] * dsilu_m(c.y));
            }
            __nv_bfloat16* drow = P.dycond + (m * 256 + ky * 16) * 32;
            st_stream16(drow + g * 32 + 8 * t, make_uint4(o0[0], o0[1], o0[2], o0[3]));
            st_stream16(drow + (g + 8) * 32 + 8 * t, make_uint4(o1[0], o1[1], o1[2], o1[3]));
        }
    }

    // ---------------- positional-table gradient of this warp's patch row
    {
        float* gpos = P.gblob + (kTBlock0_ + R * kTBlock_ + 132) + (size_t)ky * 16 * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(gpos + g * 32 + 8 * j + 2 * t, dpos[j][0]); atomicAdd(gpos + g * 32 + 8 * j + 2 * t + 1, dpos[j][1]);
            atomicAdd(gpos + (g + 8) * 32 + 8 * j + 2 * t, dpos[j][2]); atomicAdd(gpos + (g + 8) * 32 + 8 * j + 2 * t + 1, dpos[j][3]);
        }
    }
    // ---------------- flush the CTA's accumulators (shared, row stride 36 -> global blob layout, row stride 32)
    __syncthreads();
    float* G = P.gblob;
    for (int i = threadIdx.x; i < 96; i += blockDim.x) atomicAdd(G + kTWrgb_ + i, sG[mg_final(R) + 3 * kS + 4 + (i / 3) * 4 + i % 3]);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) atomicAdd(G + kTWin_ + i, sG[kGWin + (i >> 5) * kS + (i & 31)]);
    for (int i = threadIdx.x; i < 32; i += blockDim.x) atomicAdd(G + kTbin_ + i, sG[kGbin + i]);
    for (int r = 0; r < R; ++r) {
        float* gg = G + kTBlock0_ + r * kTBlock_;
        const float* a = sG + kGBlock0 + r * kGBlock;
        for (int i = threadIdx.x; i < 3072; i += blockDim.x) atomicAdd(gg + kBWada_ + i, a[kHWada + (i >> 5) * kS + (i & 31)]);
        for (int i = threadIdx.x; i < 96; i += blockDim.x) atomicAdd(gg + kBbada_ + i, a[kHbada + i]);
        for (int i = threadIdx.x; i < 32; i += blockDim.x) {
            atomicAdd(gg + kBlng_ + i, a[kHlng + i]); atomicAdd(gg + kBlnb_ + i, a[kHlnb + i]);
            atomicAdd(gg + kBb0_ + i, a[kHb0 + i]); atomicAdd(gg + kBb2_ + i, a[kHb2 + i]);
        }
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
            atomicAdd(gg + kBW0_ + i, a[kHW0 + (i >> 5) * kS + (i & 31)]);
            atomicAdd(gg + kBW2_ + i, a[kHW2 + (i >> 5) * kS + (i & 31)]);
        }
    }
    {
        float* gg = G + kTBlock0_ + R * kTBlock_;
        const float* a = sG + mg_final(R);
        for (int i = threadIdx.x; i < 96; i += blockDim.x) atomicAdd(gg + i, a[(i >> 5) * kS + (i & 31)]);
        if (threadIdx.x < 3) atomicAdd(gg + 128 + threadIdx.x, a[3 * kS + threadIdx.x]);
    }
}

}  // namespace deco

extern "C" int deco_decoder_bwd_blob_bytes(int num_res_blocks) { return deco::mb_words(num_res_blocks) * 4; }

extern "C" int deco_pixel_decoder_bwd_tc(const float* x, const void* ycond_bf16, const float* dout, const void* fwd_blob,
                                         const void* bwd_blob, const float* postab, void* dycond_bf16, float* grad_accum,
                                         int B, int H, int W, int patch, int hidden_x, int num_res_blocks, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(x && ycond_bf16 && dout && fwd_blob && bwd_blob && postab && dycond_bf16 && grad_accum,
                   "pixel_decoder_bwd_tc: null pointer");
    if (patch != 16 || hidden_x != 32) {
        deco_set_error("pixel_decoder_bwd_tc: built for patch_size 16 and hidden_size_x 32 (got %d, %d)", patch, hidden_x);
        return DECO_ERR_UNSUPPORTED;
    }
    DECO_CHECK_ARG(B > 0 && H % 16 == 0 && W % 16 == 0 && num_res_blocks >= 1 && num_res_blocks <= 3,
                   "pixel_decoder_bwd_tc: bad shape B=%d H=%d W=%d R=%d (R <= 3)", B, H, W, num_res_blocks);
    DecBwdMmaParams P;
    P.x = x; P.ycond = (const __nv_bfloat16*)ycond_bf16; P.dout = dout; P.fblob = (const uint32_t*)fwd_blob;
    P.bblob = (const uint32_t*)bwd_blob; P.postab = postab; P.dycond = (__nv_bfloat16*)dycond_bf16; P.gblob = grad_accum;
    P.R = num_res_blocks; P.H = H; P.W = W; P.Hp = H / 16; P.Wp = W / 16;
    P.M = (long long)B * P.Hp * P.Wp;
    const int R = P.R;
    const int words = ((mf_frag_words(R) + mf_vec_floats(R) + 3) & ~3) + ((mb_words(R) + 3) & ~3) + ((mg_floats(R) + 3) & ~3) +
                      8 * R * 16 * 32 + 2 * kStageWords;
    const int smem_bytes = words * 4;
    cudaError_t e = cudaFuncSetAttribute(pixel_decoder_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { deco_set_error("pixel_decoder_bwd_tc attr: %s", cudaGetErrorString(e)); return (int)e; }
    // warps = 8 * grid must be a multiple of 16 (a warp keeps one patch row): even grid
    long long grid = (P.M * 16 + 7) / 8;
    if (grid > kNumSMs) grid = kNumSMs;
    grid = (grid + 1) & ~1LL;
    pixel_decoder_bwd_mma_kernel<<<(unsigned)grid, 256, smem_bytes, (cudaStream_t)stream>>>(P);
    DECO_CHECK_LAUNCH("pixel_decoder_bwd_mma_kernel");
    return DECO_OK;
}
