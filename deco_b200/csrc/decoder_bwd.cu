// Backward of the per-pixel AdaLN-MLP pixel decoder (training step): NerfEmbedder + input_proj + R x ResBlock + final
// LayerNorm / Linear of /root/reference/src/models/transformer/dit_c2i_DeCo.py:212-248, :313-317, :329-332, :395-415
// under autograd.  One launch produces the gradient of the per-pixel condition (dycond, bf16 [M, p*p*32], the operand of
// the cond_embed dgrad / wgrad GEMMs) and the gradients of every decoder parameter.
//
// First version, built for correctness: fp32 scalar math, one THREAD per pixel.  A warp takes one patch position k and 32
// consecutive tokens (so the positional-table gradient reduces over the warp), recomputes the forward chain from the saved
// inputs (nothing but x and ycond is stashed by the forward), then walks it backwards.  Weights live in shared memory
// (fp32, the same bf16-rounded values the forward MMAs use); per-thread activation vectors live in local memory and are
// pulled into registers one layer at a time.  Weight gradients are reduced per warp through a staging tile
// (lane i owns row i of the outer product summed over the warp's 32 pixels), then accumulated per CTA in shared memory and
// flushed with one atomic per element per CTA.
#include "common.cuh"

namespace deco {

constexpr int kTW = 32;                      // decoder width
// fp32 weight blob (host: deco_b200/autograd.py::pack_decoder_train), in floats
constexpr int kTWrgb = 0;                    // [32][3]
constexpr int kTWin = 96;                    // [32][32]
constexpr int kTbin = 1120;                  // [32]
constexpr int kTBlock0 = 1152;
constexpr int kTBlock = 5344;                // Wada[96][32] bada[96] lng[32] lnb[32] W0[32][32] b0[32] W2[32][32] b2[32]
constexpr int kBWada = 0, kBbada = 3072, kBlng = 3168, kBlnb = 3200, kBW0 = 3232, kBb0 = 4256, kBW2 = 4288, kBb2 = 5312;
constexpr int kTFinal = 132;                 // Wf[4][32] (row 3 zero), bf[4]
__host__ __device__ inline int dect_blob_floats(int R) { return kTBlock0 + R * kTBlock + kTFinal; }
// shared-memory gradient accumulators: same items, matrix rows padded to 33 floats (bank-conflict-free row-per-lane adds)
constexpr int kAWrgb = 0;                    // [3][33]  (channel-major)
constexpr int kAWin = 99;                    // [32][33]
constexpr int kAbin = 1155;                  // [32]
constexpr int kABlock0 = 1187;
constexpr int kABlock = 5504;                // Wada[96][33] bada[96] lng[32] lnb[32] W0[32][33] b0[32] W2[32][33] b2[32]
constexpr int kCWada = 0, kCbada = 3168, kClng = 3264, kClnb = 3296, kCW0 = 3328, kCb0 = 4384, kCW2 = 4416, kCb2 = 5472;
constexpr int kAFinal = 136;                 // Wf[4][33], bf[4]
__host__ __device__ inline int dect_acc_floats(int R) { return kABlock0 + R * kABlock + kAFinal; }
constexpr int kStage = 32 * 36;              // one staging tile (floats)

__device__ __forceinline__ float dsilu_dec(float x) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-x));
    return s * fmaf(x, 1.0f - s, 1.0f);
}

// y[i] = bias[i] + sum_j W[i][j] x[j], i < nout (multiple of 4); W in shared memory
__device__ __noinline__ void lin_fwd(const float* W, const float* bias, const float* x, float* y, int nout)
{
    float xr[kTW];
#pragma unroll
    for (int j = 0; j < kTW; ++j) xr[j] = x[j];
#pragma unroll 1
    for (int i0 = 0; i0 < nout; i0 += 4) {
        float acc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = bias ? bias[i0 + e] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 w = *reinterpret_cast<const float4*>(W + (i0 + e) * kTW + 4 * q);
                acc[e] = fmaf(w.x, xr[4 * q], fmaf(w.y, xr[4 * q + 1], fmaf(w.z, xr[4 * q + 2], fmaf(w.w, xr[4 * q + 3], acc[e]))));
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) y[i0 + e] = acc[e];
    }
}

// dx[j] (+)= sum_i dy[i] W[i][j], i < nout
__device__ __noinline__ void lin_bwd(const float* W, const float* dy, float* dx, int nout, int accumulate)
{
#pragma unroll 1
    for (int ib = 0; ib < nout; ib += kTW) {
        float dr[kTW];
#pragma unroll
        for (int i = 0; i < kTW; ++i) dr[i] = (ib + i < nout) ? dy[ib + i] : 0.f;
        const int ni = nout - ib < kTW ? nout - ib : kTW;
#pragma unroll 1
        for (int j0 = 0; j0 < kTW; j0 += 8) {
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = (accumulate || ib > 0) ? dx[j0 + e] : 0.f;
#pragma unroll
            for (int i = 0; i < kTW; ++i) {
                if (i < ni) {
                    const float4 w0 = *reinterpret_cast<const float4*>(W + (ib + i) * kTW + j0);
                    const float4 w1 = *reinterpret_cast<const float4*>(W + (ib + i) * kTW + j0 + 4);
                    acc[0] = fmaf(dr[i], w0.x, acc[0]); acc[1] = fmaf(dr[i], w0.y, acc[1]);
                    acc[2] = fmaf(dr[i], w0.z, acc[2]); acc[3] = fmaf(dr[i], w0.w, acc[3]);
                    acc[4] = fmaf(dr[i], w1.x, acc[4]); acc[5] = fmaf(dr[i], w1.y, acc[5]);
                    acc[6] = fmaf(dr[i], w1.z, acc[6]); acc[7] = fmaf(dr[i], w1.w, acc[7]);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) dx[j0 + e] = acc[e];
        }
    }
}

// gW[(i)*33 + j] += sum_p a_p[i] * b_p[j] over the warp's 32 pixels p, i < nout (multiple of 32), j < 32;
// gb[i] += sum_p a_p[i].  sA / sB: this warp's staging tiles.
__device__ __noinline__ void outer_accum(float* sA, float* sB, const float* a, const float* b, int nout, float* gW, float* gb, int lane)
{
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(sB + lane * 36 + 4 * q) = make_float4(b[4 * q], b[4 * q + 1], b[4 * q + 2], b[4 * q + 3]);
#pragma unroll 1
    for (int base = 0; base < nout; base += kTW) {
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(sA + lane * 36 + 4 * q) = make_float4(a[base + 4 * q], a[base + 4 * q + 1], a[base + 4 * q + 2], a[base + 4 * q + 3]);
        __syncwarp();
        float acc[kTW];
#pragma unroll
        for (int j = 0; j < kTW; ++j) acc[j] = 0.f;
        float sb = 0.f;
#pragma unroll 2
        for (int p = 0; p < 32; ++p) {
            const float av = sA[p * 36 + lane];
            sb += av;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b4 = *reinterpret_cast<const float4*>(sB + p * 36 + 4 * q);
                acc[4 * q] = fmaf(av, b4.x, acc[4 * q]); acc[4 * q + 1] = fmaf(av, b4.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(av, b4.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(av, b4.w, acc[4 * q + 3]);
            }
        }
        float* row = gW + (base + lane) * 33;
#pragma unroll
        for (int j = 0; j < kTW; ++j) atomicAdd(row + j, acc[j]);
        if (gb) atomicAdd(gb + base + lane, sb);
    }
    __syncwarp();
}

// g[c * 33 + lane] += sum_p v_p[lane] * s_p[c], c < 3 (s has 4 entries, the last is ignored)
__device__ __noinline__ void outer_accum4(float* sA, float* sB, const float* v, float s0, float s1, float s2, float* g, int lane)
{
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(sA + lane * 36 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    *reinterpret_cast<float4*>(sB + lane * 36) = make_float4(s0, s1, s2, 0.f);
    __syncwarp();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
    for (int p = 0; p < 32; ++p) {
        const float av = sA[p * 36 + lane];
        const float4 b4 = *reinterpret_cast<const float4*>(sB + p * 36);
        a0 = fmaf(av, b4.x, a0); a1 = fmaf(av, b4.y, a1); a2 = fmaf(av, b4.z, a2);
    }
    atomicAdd(g + lane, a0); atomicAdd(g + 33 + lane, a1); atomicAdd(g + 66 + lane, a2);
    __syncwarp();
}

// out0[lane] += sum_p u_p[lane], out1[lane] += sum_p w_p[lane]  (either output may be NULL); out* may be global memory
__device__ __noinline__ void vec_accum2(float* sA, float* sB, const float* u, const float* w, float* out0, float* out1, int lane)
{
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        *reinterpret_cast<float4*>(sA + lane * 36 + 4 * q) = make_float4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
        if (w) *reinterpret_cast<float4*>(sB + lane * 36 + 4 * q) = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
    __syncwarp();
    float a = 0.f, c = 0.f;
#pragma unroll 4
    for (int p = 0; p < 32; ++p) {
        a += sA[p * 36 + lane];
        if (w) c += sB[p * 36 + lane];
    }
    if (out0) atomicAdd(out0 + lane, a);
    if (w && out1) atomicAdd(out1 + lane, c);
    __syncwarp();
}

__device__ __forceinline__ void ln_fwd(const float* h, float* hn, float& rstd)
{
    float s = 0.f;
#pragma unroll 8
    for (int j = 0; j < kTW; ++j) s += h[j];
    const float mean = s * (1.0f / kTW);
    float q = 0.f;
#pragma unroll 8
    for (int j = 0; j < kTW; ++j) { const float d = h[j] - mean; q = fmaf(d, d, q); }
    rstd = rsqrtf(q * (1.0f / kTW) + 1e-6f);
#pragma unroll 8
    for (int j = 0; j < kTW; ++j) hn[j] = (h[j] - mean) * rstd;
}

// dh[j] += rstd * (dhn[j] - mean(dhn) - hn[j] * mean(dhn * hn))   (set != 0: dh[j] = ...)
__device__ __forceinline__ void ln_bwd(const float* dhn, const float* hn, float rstd, float* dh, int set)
{
    float s = 0.f, q = 0.f;
#pragma unroll 8
    for (int j = 0; j < kTW; ++j) { s += dhn[j]; q = fmaf(dhn[j], hn[j], q); }
    s *= (1.0f / kTW); q *= (1.0f / kTW);
#pragma unroll 8
    for (int j = 0; j < kTW; ++j) {
        const float v = rstd * (dhn[j] - s - hn[j] * q);
        dh[j] = set ? v : dh[j] + v;
    }
}

struct DecBwdParams {
    const float* x;              // [B, 3, H, W] fp32 (decoder input image)
    const __nv_bfloat16* ycond;  // [M, p*p*32]
    const float* dout;           // [B, 3, H, W] fp32
    const float* blob;           // fp32 weights, dect_blob_floats(R)
    const float* postab;         // [p*p][32]
    __nv_bfloat16* dycond;       // [M, p*p*32]
    float* gblob;                // gradient accumulators (zeroed by the caller): dect_blob_floats(R) then dpostab [p*p][32]
    int R, H, W, Hp, Wp;
    long long M;
};

__global__ void __launch_bounds__(256, 1) pixel_decoder_bwd_kernel(DecBwdParams P)
{
    extern __shared__ __align__(16) float dsm[];
    const int nW = dect_blob_floats(P.R), nA = dect_acc_floats(P.R);
    float* sWt = dsm;
    float* sAcc = dsm + ((nW + 3) & ~3);
    float* sStage = sAcc + ((nA + 3) & ~3);
    for (int i = threadIdx.x; i < nW; i += blockDim.x) sWt[i] = __ldg(P.blob + i);
    for (int i = threadIdx.x; i < nA; i += blockDim.x) sAcc[i] = 0.f;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sA = sStage + warp * 2 * kStage;
    float* sB = sA + kStage;
    const int L = P.Hp * P.Wp;
    const long long groups = (P.M + 31) / 32;
    const long long items = groups * 256;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    const size_t plane = (size_t)P.H * P.W;
    float* gpos = P.gblob + nW;

    for (long long it = (long long)blockIdx.x * (blockDim.x >> 5) + warp; it < items; it += wstride) {
        const int k = (int)(it & 255);
        const long long m = (it >> 8) * 32 + lane;
        const bool valid = m < P.M;
        const long long b = valid ? m / L : 0;
        const int tok = valid ? (int)(m % L) : 0;
        const int py = tok / P.Wp, px = tok % P.Wp;
        const int ky = k >> 4, kx = k & 15;
        const size_t pix = (size_t)(py * 16 + ky) * P.W + px * 16 + kx;
        float rgb[3], dout[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            rgb[c] = valid ? round_bf(__ldg(P.x + ((size_t)b * 3 + c) * plane + pix)) : 0.f;
            dout[c] = valid ? __ldg(P.dout + ((size_t)b * 3 + c) * plane + pix) : 0.f;
        }
        float yv[kTW], ys[kTW];
        if (valid) {
            const __nv_bfloat16* yp = P.ycond + ((size_t)m * 256 + k) * kTW;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 v = ld_stream16(yp + 8 * q);
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = unpack_bf2(w4[e]);
                    yv[8 * q + 2 * e] = f.x; yv[8 * q + 2 * e + 1] = f.y;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < kTW; ++j) yv[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < kTW; ++j) ys[j] = silu_f(yv[j]);

        // ---------------- forward recompute, keeping the input of every res-block
        float e0[kTW], h[kTW], hs[3][kTW];
        {
            const float* T = P.postab + k * kTW;
#pragma unroll 8
            for (int i = 0; i < kTW; ++i) {
                const float* wr = sWt + kTWrgb + i * 3;
                e0[i] = __ldg(T + i) + wr[0] * rgb[0] + wr[1] * rgb[1] + wr[2] * rgb[2];
            }
        }
        lin_fwd(sWt + kTWin, sWt + kTbin, e0, h, kTW);
        float mod[96], hn[kTW], hm[kTW], z[kTW], act[kTW], mm[kTW];
        float rstd;
#pragma unroll 1
        for (int r = 0; r < P.R; ++r) {
            const float* wb = sWt + kTBlock0 + r * kTBlock;
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) hs[r][j] = h[j];
            lin_fwd(wb + kBWada, wb + kBbada, ys, mod, 96);
            ln_fwd(h, hn, rstd);
#pragma unroll 8
            for (int j = 0; j < kTW; ++j)
                hm[j] = fmaf(fmaf(hn[j], wb[kBlng + j], wb[kBlnb + j]), 1.0f + mod[32 + j], mod[j]);
            lin_fwd(wb + kBW0, wb + kBb0, hm, z, kTW);
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) act[j] = silu_f(z[j]);
            lin_fwd(wb + kBW2, wb + kBb2, act, mm, kTW);
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) h[j] = fmaf(mod[64 + j], mm[j], h[j]);
        }

        // ---------------- backward: final LayerNorm (no affine) + Linear 32 -> 3
        float dh[kTW], dys[kTW], tmp[kTW], tmp2[kTW];
#pragma unroll 8
        for (int j = 0; j < kTW; ++j) dys[j] = 0.f;
        {
            const float* wf = sWt + kTBlock0 + P.R * kTBlock;
            float* af = sAcc + kABlock0 + P.R * kABlock;
            ln_fwd(h, hn, rstd);
            outer_accum4(sA, sB, hn, dout[0], dout[1], dout[2], af, lane);
            {   // bias: sum of dout over the warp
                float d0 = warp_sum(dout[0]), d1 = warp_sum(dout[1]), d2 = warp_sum(dout[2]);
                if (lane == 0) { atomicAdd(af + 132, d0); atomicAdd(af + 133, d1); atomicAdd(af + 134, d2); }
            }
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) tmp[j] = dout[0] * wf[j] + dout[1] * wf[32 + j] + dout[2] * wf[64 + j];
            ln_bwd(tmp, hn, rstd, dh, 1);
        }

        // ---------------- backward through the res-blocks (forward of the block recomputed from its saved input)
#pragma unroll 1
        for (int r = P.R - 1; r >= 0; --r) {
            const float* wb = sWt + kTBlock0 + r * kTBlock;
            float* ab = sAcc + kABlock0 + r * kABlock;
            float hl[kTW], dmod[96];
            lin_fwd(wb + kBWada, wb + kBbada, ys, mod, 96);
            ln_fwd(hs[r], hn, rstd);
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) {
                hl[j] = fmaf(hn[j], wb[kBlng + j], wb[kBlnb + j]);
                hm[j] = fmaf(hl[j], 1.0f + mod[32 + j], mod[j]);
            }
            lin_fwd(wb + kBW0, wb + kBb0, hm, z, kTW);
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) act[j] = silu_f(z[j]);
            lin_fwd(wb + kBW2, wb + kBb2, act, mm, kTW);
            // h_out = h_in + gate * mm
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) { dmod[64 + j] = dh[j] * mm[j]; tmp[j] = dh[j] * mod[64 + j]; }   // tmp = dmm
            outer_accum(sA, sB, tmp, act, kTW, ab + kCW2, ab + kCb2, lane);
            lin_bwd(wb + kBW2, tmp, tmp2, kTW, 0);                                                          // tmp2 = dact
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) tmp[j] = tmp2[j] * dsilu_dec(z[j]);                               // tmp = dz
            outer_accum(sA, sB, tmp, hm, kTW, ab + kCW0, ab + kCb0, lane);
            lin_bwd(wb + kBW0, tmp, tmp2, kTW, 0);                                                          // tmp2 = dhm
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) {
                dmod[j] = tmp2[j];
                dmod[32 + j] = tmp2[j] * hl[j];
                tmp[j] = tmp2[j] * (1.0f + mod[32 + j]);            // dhl
                tmp2[j] = tmp[j] * hn[j];                           // dhl * hn -> d ln weight
            }
            vec_accum2(sA, sB, tmp2, tmp, ab + kClng, ab + kClnb, lane);
#pragma unroll 8
            for (int j = 0; j < kTW; ++j) tmp[j] *= wb[kBlng + j];  // dhn
            ln_bwd(tmp, hn, rstd, dh, 0);
            outer_accum(sA, sB, dmod, ys, 96, ab + kCWada, ab + kCbada, lane);
            lin_bwd(wb + kBWada, dmod, dys, 96, 1);
        }

        // ---------------- input_proj, NerfEmbedder
        outer_accum(sA, sB, dh, e0, kTW, sAcc + kAWin, sAcc + kAbin, lane);
        lin_bwd(sWt + kTWin, dh, tmp, kTW, 0);                      // tmp = de
        outer_accum4(sA, sB, tmp, rgb[0], rgb[1], rgb[2], sAcc + kAWrgb, lane);
        vec_accum2(sA, sB, tmp, nullptr, gpos + k * kTW, nullptr, lane);

        // ---------------- d ycond = d silu(y) * silu'(y)
        if (valid) {
            __nv_bfloat16* dp = P.dycond + ((size_t)m * 256 + k) * kTW;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 8 * q + 2 * e;
                    o[e] = pack_bf2(dys[j] * dsilu_dec(yv[j]), dys[j + 1] * dsilu_dec(yv[j + 1]));
                }
                *reinterpret_cast<uint4*>(dp + 8 * q) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }

    // ---------------- flush the CTA's accumulators
    __syncthreads();
    float* G = P.gblob;
    for (int i = threadIdx.x; i < 96; i += blockDim.x) atomicAdd(G + kTWrgb + i, sAcc[kAWrgb + (i % 3) * 33 + i / 3]);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) atomicAdd(G + kTWin + i, sAcc[kAWin + (i >> 5) * 33 + (i & 31)]);
    for (int i = threadIdx.x; i < 32; i += blockDim.x) atomicAdd(G + kTbin + i, sAcc[kAbin + i]);
    for (int r = 0; r < P.R; ++r) {
        float* g = G + kTBlock0 + r * kTBlock;
        const float* a = sAcc + kABlock0 + r * kABlock;
        for (int i = threadIdx.x; i < 3072; i += blockDim.x) atomicAdd(g + kBWada + i, a[kCWada + (i >> 5) * 33 + (i & 31)]);
        for (int i = threadIdx.x; i < 96; i += blockDim.x) atomicAdd(g + kBbada + i, a[kCbada + i]);
        for (int i = threadIdx.x; i < 32; i += blockDim.x) {
            atomicAdd(g + kBlng + i, a[kClng + i]); atomicAdd(g + kBlnb + i, a[kClnb + i]);
            atomicAdd(g + kBb0 + i, a[kCb0 + i]); atomicAdd(g + kBb2 + i, a[kCb2 + i]);
        }
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
            atomicAdd(g + kBW0 + i, a[kCW0 + (i >> 5) * 33 + (i & 31)]);
            atomicAdd(g + kBW2 + i, a[kCW2 + (i >> 5) * 33 + (i & 31)]);
        }
    }
    {
        float* g = G + kTBlock0 + P.R * kTBlock;
        const float* a = sAcc + kABlock0 + P.R * kABlock;
        for (int i = threadIdx.x; i < 96; i += blockDim.x) atomicAdd(g + i, a[(i >> 5) * 33 + (i & 31)]);
        if (threadIdx.x < 3) atomicAdd(g + 128 + threadIdx.x, a[132 + threadIdx.x]);
    }
}

}  // namespace deco

extern "C" int deco_decoder_train_blob_floats(int num_res_blocks) { return deco::dect_blob_floats(num_res_blocks); }

extern "C" int deco_pixel_decoder_bwd(const float* x, const void* ycond_bf16, const float* dout, const float* blob_f32,
                                      const float* postab, void* dycond_bf16, float* grad_accum, int B, int H, int W,
                                      int patch, int hidden_x, int num_res_blocks, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(x && ycond_bf16 && dout && blob_f32 && postab && dycond_bf16 && grad_accum, "pixel_decoder_bwd: null pointer");
    if (patch != 16 || hidden_x != kTW) {
        deco_set_error("pixel_decoder_bwd: built for patch_size 16 and hidden_size_x 32 (got %d, %d)", patch, hidden_x);
        return DECO_ERR_UNSUPPORTED;
    }
    DECO_CHECK_ARG(B > 0 && H % 16 == 0 && W % 16 == 0 && num_res_blocks >= 1 && num_res_blocks <= 3,
                   "pixel_decoder_bwd: bad shape B=%d H=%d W=%d R=%d (R <= 3)", B, H, W, num_res_blocks);
    DecBwdParams P;
    P.x = x; P.ycond = (const __nv_bfloat16*)ycond_bf16; P.dout = dout; P.blob = blob_f32; P.postab = postab;
    P.dycond = (__nv_bfloat16*)dycond_bf16; P.gblob = grad_accum;
    P.R = num_res_blocks; P.H = H; P.W = W; P.Hp = H / 16; P.Wp = W / 16;
    P.M = (long long)B * P.Hp * P.Wp;
    const int nW = (dect_blob_floats(P.R) + 3) & ~3, nA = (dect_acc_floats(P.R) + 3) & ~3;
    const int smem_bytes = (nW + nA + 8 * 2 * kStage) * 4;
    cudaError_t e = cudaFuncSetAttribute(pixel_decoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { deco_set_error("pixel_decoder_bwd attr: %s", cudaGetErrorString(e)); return (int)e; }
    const long long items = ((P.M + 31) / 32) * 256;
    long long grid = (items + 7) / 8;
    if (grid > kNumSMs) grid = kNumSMs;
    pixel_decoder_bwd_kernel<<<(unsigned)grid, 256, smem_bytes, (cudaStream_t)stream>>>(P);
    DECO_CHECK_LAUNCH("pixel_decoder_bwd_kernel");
    return DECO_OK;
}
