// Frequency-aware flow-matching loss: fused RGB->YCbCr, 8x8 block DCT, JPEG-weighted MSE, plus the plain
// FM MSE, forward and backward in ONE pass over `out` and `v_t`.
//
// Replaces (reference, /root/reference): src/diffusion/flow_matching/training_repa_DeCo.py
//   :106-114 _rgb2ycbcr, :116-136 _dct, :138-195 _build_freq_weight (weights are passed in), :273-285 loss.
//
// Linear algebra: dct(ycbcr(out)) - dct(ycbcr(v_t)) == dct(ycbcr(out - v_t)), so the kernel transforms the
// difference d = out - v_t once.  Bound: HBM.  Algorithmic bytes / pixel-channel: fwd 8 B (read out, v_t fp32),
// fused fwd+bwd 12 B (+ grad write).  ~90 FLOP / pixel-channel -> ~8 FLOP/B, far below the fp32 ridge.
//
// Tiling: one CTA = 128 threads = an 8-row x 128-column strip, all 3 channels.
//   pass 1: thread c owns image column c: 24 coalesced scalar loads per input (a warp reads 128 contiguous
//           bytes per load), colour transform in registers, vertical 8-point DCT in registers.
//   smem  : Z[ch][k][c] with row stride 129 floats (conflict-free for both access patterns).
//   pass 2: thread (blk=tid/8, k=tid%8) owns the 8 horizontally adjacent values of vertical frequency k of
//           8x8 block blk: horizontal 8-point DCT in registers, weighted square, warp-shuffle reduction.
//   backward retraces the same two passes with the transposed transforms and writes grad coalesced.
// Ragged H/W (not multiples of 8) follow the reference's reflect padding: loads are index-mirrored and the
// gradient of the padded copies is scattered back with atomics (rare path; the hot path uses plain stores).
#include "common.cuh"

namespace deco {

struct DctConsts {
    float c[8][8];  // orthonormal DCT-II matrix C[k][n]
};

__constant__ DctConsts g_dct;

// 8-point forward DCT y[k] = sum_n C[k][n] x[n], even/odd split (40 FMA instead of 64)
__device__ __forceinline__ void dct8_fwd(const float (&x)[8], float (&y)[8]) {
    float s[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[i] = x[i] + x[7 - i]; d[i] = x[i] - x[7 - i]; }
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) a = fmaf(g_dct.c[k][i], s[i], a);
        y[k] = a;
    }
#pragma unroll
    for (int k = 1; k < 8; k += 2) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) a = fmaf(g_dct.c[k][i], d[i], a);
        y[k] = a;
    }
}
// transposed transform x[n] = sum_k C[k][n] y[k]
__device__ __forceinline__ void dct8_bwd(const float (&y)[8], float (&x)[8]) {
    float e[4], o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k += 2) a = fmaf(g_dct.c[k][i], y[k], a);
#pragma unroll
        for (int k = 1; k < 8; k += 2) b = fmaf(g_dct.c[k][i], y[k], b);
        e[i] = a; o[i] = b;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = e[i] + o[i]; x[7 - i] = e[i] - o[i]; }
}

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < n ? i : 2 * n - 2 - i; }


constexpr int kTileW = 128;
constexpr int kZStride = 129;

template <typename TOut, bool kRagged, bool kLoss, bool kGrad>
__global__ void __launch_bounds__(kTileW) dct_fm_loss_kernel(
    const TOut* __restrict__ out, const float* __restrict__ vt, const float* __restrict__ freq_w,
    TOut* __restrict__ grad, double* __restrict__ accum, const float* __restrict__ upstream,
    int H, int W, int H2, int W2, float fm_scale, float freq_scale, float freq_loss_weight)
{
    __shared__ float Z[3][8][kZStride];
    __shared__ float red[2][kTileW / 32];

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 8;
    const int c = blockIdx.x * kTileW + tid;
    const bool col_ok = c < W2;
    const int cs = kRagged ? reflect_idx(min(c, W2 - 1), W) : c;   // source column
    const size_t plane = (size_t)H * W;
    const TOut* ob = out + (size_t)b * 3 * plane;
    const float* vb = vt + (size_t)b * 3 * plane;

    float d[3][8];
    float fm_part = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int rs = kRagged ? reflect_idx(r0 + r, H) : (r0 + r);
        float dr = 0.f, dg = 0.f, db = 0.f;
        if (col_ok) {
            const size_t o = (size_t)rs * W + cs;
            dr = ldf(ob + o) - __ldg(vb + o);
            dg = ldf(ob + plane + o) - __ldg(vb + plane + o);
            db = ldf(ob + 2 * plane + o) - __ldg(vb + 2 * plane + o);
            // FM term counts real pixels only (padded copies are not part of `out`)
            if (!kRagged || (r0 + r < H && c < W)) fm_part += dr * dr + dg * dg + db * db;
        }
        d[0][r] = 0.299f * dr + 0.587f * dg + 0.114f * db;
        d[1][r] = -0.168736f * dr - 0.331264f * dg + 0.5f * db;
        d[2][r] = 0.5f * dr - 0.418688f * dg - 0.081312f * db;
    }
    // vertical DCT, stage to smem
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float y[8];
        dct8_fwd(d[ch], y);
#pragma unroll
        for (int k = 0; k < 8; ++k) Z[ch][k][tid] = y[k];
    }
    __syncthreads();

    const int k = tid & 7, blk = tid >> 3;
    float fq_part = 0.f;
    float gcoef[3][8];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = Z[ch][k][blk * 8 + j];
        dct8_fwd(x, y);
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            const float w = __ldg(freq_w + ch * 64 + k * 8 + l);
            fq_part = fmaf(w * y[l], y[l], fq_part);
            if (kGrad) gcoef[ch][l] = w * y[l];
        }
    }

    if (kLoss) {
        fm_part = warp_sum(fm_part);
        fq_part = warp_sum(fq_part);
        if ((tid & 31) == 0) { red[0][tid >> 5] = fm_part; red[1][tid >> 5] = fq_part; }
    }
    if (kGrad) {
        __syncthreads();   // everyone finished reading Z
        // d loss / d coef = upstream * freq_loss_weight * 2 * w * coef / N_freq ; fold the scalar in at the end
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float x[8];
            dct8_bwd(gcoef[ch], x);
#pragma unroll
            for (int j = 0; j < 8; ++j) Z[ch][k][blk * 8 + j] = x[j];
        }
    }
    __syncthreads();
    if (kLoss && tid == 0) {
        float a = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < kTileW / 32; ++i) { a += red[0][i]; q += red[1][i]; }
        atomicAdd(accum + 0, (double)a);
        atomicAdd(accum + 1, (double)q);
    }
    if (kGrad) {
        const float up = upstream ? __ldg(upstream) : 1.0f;
        const float kf = up * freq_loss_weight * 2.0f * freq_scale;   // freq_scale = 1/N_freq
        const float km = up * 2.0f * fm_scale;                        // fm_scale = 1/N_fm
        float gy[3][8];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float y[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) y[kk] = Z[ch][kk][tid];
            dct8_bwd(y, gy[ch]);
        }
        if (col_ok) {
            TOut* gb = grad + (size_t)b * 3 * plane;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float a0 = gy[0][r] * kf, a1 = gy[1][r] * kf, a2 = gy[2][r] * kf;
                // transpose of the colour matrix
                float gr = 0.299f * a0 - 0.168736f * a1 + 0.5f * a2;
                float gg = 0.587f * a0 - 0.331264f * a1 - 0.418688f * a2;
                float gbv = 0.114f * a0 + 0.5f * a1 - 0.081312f * a2;
                if (!kRagged) {
                    // recover d in RGB: cheaper to reload than to keep 24 more registers live
                    const size_t o = (size_t)(r0 + r) * W + c;
                    const float dr = ldf(ob + o) - __ldg(vb + o);
                    const float dg = ldf(ob + plane + o) - __ldg(vb + plane + o);
                    const float db = ldf(ob + 2 * plane + o) - __ldg(vb + 2 * plane + o);
                    stf(gb + o, fmaf(km, dr, gr));
                    stf(gb + plane + o, fmaf(km, dg, gg));
                    stf(gb + 2 * plane + o, fmaf(km, db, gbv));
                } else {
                    // ragged: grad is fp32-accumulated with atomics (several padded copies map to one pixel)
                    const int rs = reflect_idx(r0 + r, H);
                    const size_t o = (size_t)rs * W + cs;
                    if (r0 + r < H && c < W) {
                        const float dr = ldf(ob + o) - __ldg(vb + o);
                        const float dg = ldf(ob + plane + o) - __ldg(vb + plane + o);
                        const float db = ldf(ob + 2 * plane + o) - __ldg(vb + 2 * plane + o);
                        gr = fmaf(km, dr, gr); gg = fmaf(km, dg, gg); gbv = fmaf(km, db, gbv);
                    }
                    float* gf = reinterpret_cast<float*>(grad) + (size_t)b * 3 * plane;
                    atomicAdd(gf + o, gr);
                    atomicAdd(gf + plane + o, gg);
                    atomicAdd(gf + 2 * plane + o, gbv);
                }
            }
        }
    }
}

__global__ void dct_fm_finalize_kernel(const double* __restrict__ accum, float* __restrict__ losses,
                                       double inv_fm, double inv_fq, float freq_loss_weight) {
    const double fm = accum[0] * inv_fm, fq = accum[1] * inv_fq;
    losses[0] = (float)fm;
    losses[1] = (float)fq;
    losses[2] = (float)(fm + (double)freq_loss_weight * fq);
}

static bool g_dct_init = false;
static int ensure_dct_consts() {
    // per-device constant upload; cheap enough to redo whenever the current device changes
    static int dev_done = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (g_dct_init && dev_done == dev) return 0;
    DctConsts h;
    for (int k = 0; k < 8; ++k)
        for (int n = 0; n < 8; ++n) {
            // float32 evaluation order of training_repa_DeCo.py:95-104: alpha * cos(pi*(2n+1)*k/(2N))
            float ang = 3.14159265358979323846f * (2.0f * n + 1.0f) * k / 16.0f;
            float alpha = (k == 0) ? sqrtf(1.0f / 8.0f) : sqrtf(2.0f / 8.0f);
            h.c[k][n] = alpha * cosf(ang);
        }
    cudaError_t e = cudaMemcpyToSymbol(g_dct, &h, sizeof(h));
    if (e != cudaSuccess) { deco_set_error("dct const upload: %s", cudaGetErrorString(e)); return (int)e; }
    g_dct_init = true;
    dev_done = dev;
    return 0;
}

template <typename TOut>
static int launch_dct(const TOut* out, const float* vt, const float* freq_w, TOut* grad, double* accum,
                      float* losses, const float* upstream, int B, int H, int W, float flw,
                      bool want_loss, bool want_grad, cudaStream_t st)
{
    const int H2 = (H + 7) / 8 * 8, W2 = (W + 7) / 8 * 8;
    const bool ragged = (H2 != H) || (W2 != W);
    const double n_fm = (double)B * 3 * H * W, n_fq = (double)B * 3 * H2 * W2;
    dim3 grid((W2 + kTileW - 1) / kTileW, H2 / 8, B), block(kTileW);
    if (want_loss) cudaMemsetAsync(accum, 0, 2 * sizeof(double), st);
    const float fms = (float)(1.0 / n_fm), fqs = (float)(1.0 / n_fq);
#define DCT_LAUNCH(R, L, G) dct_fm_loss_kernel<TOut, R, L, G><<<grid, block, 0, st>>>( \
        out, vt, freq_w, grad, accum, upstream, H, W, H2, W2, fms, fqs, flw)
    if (ragged) {
        if (want_grad) cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)B * 3 * H * W, st);
        if (want_loss && want_grad) DCT_LAUNCH(true, true, true);
        else if (want_grad) DCT_LAUNCH(true, false, true);
        else DCT_LAUNCH(true, true, false);
    } else {
        if (want_loss && want_grad) DCT_LAUNCH(false, true, true);
        else if (want_grad) DCT_LAUNCH(false, false, true);
        else DCT_LAUNCH(false, true, false);
    }
#undef DCT_LAUNCH
    DECO_CHECK_LAUNCH("dct_fm_loss_kernel");
    if (want_loss) {
        dct_fm_finalize_kernel<<<1, 1, 0, st>>>(accum, losses, 1.0 / n_fm, 1.0 / n_fq, flw);
        DECO_CHECK_LAUNCH("dct_fm_finalize_kernel");
    }
    return DECO_OK;
}

}  // namespace deco

extern "C" int deco_dct_fm_loss(const void* out, int out_is_bf16, const float* v_t, const float* freq_w,
                                int B, int H, int W, float freq_loss_weight,
                                float* losses, void* grad, const float* upstream,
                                double* accum, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(out && v_t && freq_w && accum, "dct_fm_loss: null pointer");
    DECO_CHECK_ARG(B > 0 && H >= 2 && W >= 2, "dct_fm_loss: bad shape B=%d H=%d W=%d", B, H, W);
    DECO_CHECK_ARG(losses || grad, "dct_fm_loss: nothing to compute");
    const bool ragged = (H % 8) || (W % 8);
    DECO_CHECK_ARG(!(ragged && grad && out_is_bf16), "dct_fm_loss: ragged sizes need an fp32 grad buffer");
    int rc = ensure_dct_consts();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_is_bf16)
        return launch_dct<__nv_bfloat16>((const __nv_bfloat16*)out, v_t, freq_w, (__nv_bfloat16*)grad, accum, losses,
                                         upstream, B, H, W, freq_loss_weight, losses != nullptr, grad != nullptr, st);
    return launch_dct<float>((const float*)out, v_t, freq_w, (float*)grad, accum, losses, upstream, B, H, W,
                             freq_loss_weight, losses != nullptr, grad != nullptr, st);
}
