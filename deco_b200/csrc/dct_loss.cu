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

// ---------------------------------------------------------------- vectorised kernel for H, W multiples of 8
// Same 8 x 128 strip per CTA, but the HORIZONTAL transform runs first so that global traffic is 128-bit:
//   pass 1: thread (r = tid / 16, bx = tid % 16) owns the 8 consecutive pixels of image row r0 + r in 8x8 block bx:
//           two float4 (one uint4 for bf16) per plane = 12 LDG.128 per thread, all in flight before the first use;
//           colour transform + horizontal 8-point DCT in registers -> Z[ch][r][8 bx + l].
//   pass 2: thread c owns column c: vertical DCT, weighted square (loss), gradient coefficients, transposed vertical
//           DCT back into Z.
//   pass 3: thread (r, bx) again: transposed horizontal DCT, transposed colour matrix, + FM gradient from the d it
//           still holds in registers (nothing is re-read), 6 STG.128 per thread.
// The last CTA to finish turns the two double accumulators into the three losses and re-zeroes the scratch, so the
// whole loss (forward + backward) is ONE launch.
constexpr int kZs = 132;   // row stride of Z in floats: 16-byte aligned rows, conflict-free column reads

template <typename T> __device__ __forceinline__ void load_row8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load_row8<float>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load_row8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const float2 a = unpack_bf2(q.x), b = unpack_bf2(q.y), c = unpack_bf2(q.z), d = unpack_bf2(q.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <typename T> __device__ __forceinline__ void store_row8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store_row8<float>(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store_row8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf2(v[0], v[1]), pack_bf2(v[2], v[3]), pack_bf2(v[4], v[5]), pack_bf2(v[6], v[7]));
}

constexpr int kDctMaxCtas = 148 * 8 * 2;          // per-CTA partial-sum slots in the scratch (grid never exceeds 8 CTAs per SM)

// Persistent, balanced grid: the 8 x 128 strips are dealt to the CTAs in equal shares (launch_dct sizes the grid so that
// every CTA gets the same number of strips whenever the strip count allows it: no partial last wave), a CTA accumulates its
// partial sums in registers over all its strips and leaves them in ITS slot of the scratch -- no floating-point atomics, so
// the loss is bit-reproducible run to run -- and takes a ticket; the last CTA adds the slots up in a fixed order,
// publishes the three losses and clears the ticket: forward + backward stay ONE launch.  (The first version did two double
// atomicAdds and a ticket per strip on the same three addresses: at 32 images the 6144 serialised L2 atomics, not HBM,
// set the kernel's duration.)
template <typename TOut, bool kLoss, bool kGrad>
__global__ void __launch_bounds__(kTileW, 8) dct_fm_loss_vec_kernel(
    const TOut* __restrict__ out, const float* __restrict__ vt, const float* __restrict__ freq_w,
    TOut* __restrict__ grad, double* __restrict__ accum, float* __restrict__ losses, const float* __restrict__ upstream,
    int H, int W, int tiles_x, int tiles_y, int ntiles, int tiles_per_cta,
    float fm_scale, float freq_scale, float freq_loss_weight, double inv_fm, double inv_fq)
{
    __shared__ __align__(16) float Z[3][8][kZs];
    __shared__ float red[2][kTileW / 32];
    __shared__ float fw[3 * 64];
    __shared__ int is_last;

    const int tid = threadIdx.x;
    const int r = tid >> 4, bx = tid & 15;
    const size_t plane = (size_t)H * W;
    for (int i = tid; i < 3 * 64; i += kTileW) fw[i] = __ldg(freq_w + i);
    float fm_part = 0.f, fq_part = 0.f;
    const int t_begin = blockIdx.x * tiles_per_cta;
    const int t_end = min(ntiles, t_begin + tiles_per_cta);

    for (int tile = t_begin; tile < t_end; ++tile) {
    const int txy = tiles_x * tiles_y;
    const int b = tile / txy;
    const int rem = tile - b * txy;
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    const int r0 = ty * 8;
    const int c0 = tx * kTileW + bx * 8;
    const bool ok = c0 < W;                       // W % 8 == 0: a block is inside or outside as a whole
    const size_t o = (size_t)b * 3 * plane + (size_t)(r0 + r) * W + c0;

    float d[3][8];                                // out - v_t per RGB plane, kept for the FM gradient
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int j = 0; j < 8; ++j) d[ch][j] = 0.f;
    if (ok) {
        float a[3][8], v[3][8];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) { load_row8<TOut>(out + o + ch * plane, a[ch]); load_row8<float>(vt + o + ch * plane, v[ch]); }
        asm volatile("" ::: "memory");            // compiler barrier: all 12 loads are issued before the first use
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int j = 0; j < 8; ++j) d[ch][j] = a[ch][j] - v[ch][j];
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
#pragma unroll
        for (int j = 0; j < 8; ++j) fm_part = fmaf(d[ch][j], d[ch][j], fm_part);
    __syncthreads();                              // the previous strip's pass 3 has finished reading Z (and fw is staged)
    {
        float ycc[3][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ycc[0][j] = 0.299f * d[0][j] + 0.587f * d[1][j] + 0.114f * d[2][j];
            ycc[1][j] = -0.168736f * d[0][j] - 0.331264f * d[1][j] + 0.5f * d[2][j];
            ycc[2][j] = 0.5f * d[0][j] - 0.418688f * d[1][j] - 0.081312f * d[2][j];
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float y[8];
            dct8_fwd(ycc[ch], y);                 // horizontal: y[l], l = horizontal frequency
            float4* zp = reinterpret_cast<float4*>(&Z[ch][r][bx * 8]);
            zp[0] = make_float4(y[0], y[1], y[2], y[3]);
            zp[1] = make_float4(y[4], y[5], y[6], y[7]);
        }
    }
    __syncthreads();
    // ---- pass 2: vertical transform on column tid (horizontal frequency l = tid % 8)
    {
        const int l = tid & 7;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float x[8], y[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) x[rr] = Z[ch][rr][tid];
            dct8_fwd(x, y);                       // y[k], k = vertical frequency
            float g[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float w = fw[ch * 64 + k * 8 + l];
                fq_part = fmaf(w * y[k], y[k], fq_part);
                g[k] = w * y[k];
            }
            if (kGrad) {
                float xb[8];
                dct8_bwd(g, xb);
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) Z[ch][rr][tid] = xb[rr];   // own column: no hazard with other threads
            }
        }
    }
    if (kGrad) {
        __syncthreads();
        if (ok) {
            const float up = upstream ? __ldg(upstream) : 1.0f;
            const float kf = up * freq_loss_weight * 2.0f * freq_scale;
            const float km = up * 2.0f * fm_scale;
            float gy[3][8];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float4* zp = reinterpret_cast<const float4*>(&Z[ch][r][bx * 8]);
                const float4 z0 = zp[0], z1 = zp[1];
                const float y[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                dct8_bwd(y, gy[ch]);
            }
            float gr[8], gg[8], gb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a0 = gy[0][j] * kf, a1 = gy[1][j] * kf, a2 = gy[2][j] * kf;
                gr[j] = fmaf(km, d[0][j], 0.299f * a0 - 0.168736f * a1 + 0.5f * a2);
                gg[j] = fmaf(km, d[1][j], 0.587f * a0 - 0.331264f * a1 - 0.418688f * a2);
                gb[j] = fmaf(km, d[2][j], 0.114f * a0 + 0.5f * a1 - 0.081312f * a2);
            }
            store_row8<TOut>(grad + o, gr);
            store_row8<TOut>(grad + o + plane, gg);
            store_row8<TOut>(grad + o + 2 * plane, gb);
        }
    }
    }   // strips of this CTA

    if (kLoss) {
        fm_part = warp_sum(fm_part);
        fq_part = warp_sum(fq_part);
        if ((tid & 31) == 0) { red[0][tid >> 5] = fm_part; red[1][tid >> 5] = fq_part; }
        __syncthreads();
        double* slots = accum + 4;                // [gridDim.x][2]
        if (tid == 0) {
            float a = 0.f, q = 0.f;
#pragma unroll
            for (int i = 0; i < kTileW / 32; ++i) { a += red[0][i]; q += red[1][i]; }
            slots[2 * blockIdx.x] = (double)a;
            slots[2 * blockIdx.x + 1] = (double)q;
            __threadfence();
            unsigned int* ticket = reinterpret_cast<unsigned int*>(accum + 2);
            is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (is_last) {
            // every CTA's slot is visible (fence + atomic ticket); fixed summation order: thread-strided, then a tree
            __threadfence();
            __shared__ double part[2][kTileW];
            double a = 0.0, q = 0.0;
            const volatile double* vs = slots;
            for (int i = tid; i < (int)gridDim.x; i += kTileW) { a += vs[2 * i]; q += vs[2 * i + 1]; }
            part[0][tid] = a; part[1][tid] = q;
            __syncthreads();
            for (int sft = kTileW / 2; sft > 0; sft >>= 1) {
                if (tid < sft) { part[0][tid] += part[0][tid + sft]; part[1][tid] += part[1][tid + sft]; }
                __syncthreads();
            }
            if (tid == 0) {
                const double fm = part[0][0] * inv_fm, fq = part[1][0] * inv_fq;
                losses[0] = (float)fm;
                losses[1] = (float)fq;
                losses[2] = (float)(fm + (double)freq_loss_weight * fq);
                *reinterpret_cast<unsigned int*>(accum + 2) = 0u;      // scratch contract: zero again on exit
            }
        }
    }
}

__global__ void dct_fm_finalize_kernel(double* __restrict__ accum, float* __restrict__ losses,
                                       double inv_fm, double inv_fq, float freq_loss_weight) {
    const double fm = accum[0] * inv_fm, fq = accum[1] * inv_fq;
    losses[0] = (float)fm;
    losses[1] = (float)fq;
    losses[2] = (float)(fm + (double)freq_loss_weight * fq);
    accum[0] = 0.0; accum[1] = 0.0; accum[2] = 0.0;   // same scratch contract as the one-launch path: zero on exit
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
    if (want_loss && ragged) cudaMemsetAsync(accum, 0, 3 * sizeof(double), st);
    const float fms = (float)(1.0 / n_fm), fqs = (float)(1.0 / n_fq);
#define DCT_LAUNCH(R, L, G) dct_fm_loss_kernel<TOut, R, L, G><<<grid, block, 0, st>>>( \
        out, vt, freq_w, grad, accum, upstream, H, W, H2, W2, fms, fqs, flw)
    if (ragged) {
        if (want_grad) cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)B * 3 * H * W, st);
        if (want_loss && want_grad) DCT_LAUNCH(true, true, true);
        else if (want_grad) DCT_LAUNCH(true, false, true);
        else DCT_LAUNCH(true, true, false);
    } else {
        // scratch contract of the one-launch path: accum[0..2] are zero on entry and zero again on exit
        // equal shares: t strips per CTA with t = ceil(strips / resident CTA slots); grid = ceil(strips / t)
        const int tiles_x = (W2 + kTileW - 1) / kTileW, tiles_y = H2 / 8;
        const long long ntiles_ll = (long long)tiles_x * tiles_y * B;
        if (ntiles_ll > 0x7fffffffLL) { deco_set_error("dct_fm_loss: too many strips"); return DECO_ERR_ARG; }
        const int ntiles = (int)ntiles_ll;
        const int resident = device_sm_count() * 8;
        const int per = (ntiles + resident - 1) / resident;
        const int nctas = (ntiles + per - 1) / per;               // <= resident <= kDctMaxCtas
        if (nctas > kDctMaxCtas) { deco_set_error("dct_fm_loss: grid exceeds the scratch slots"); return DECO_ERR_ARG; }
#define DCT_VEC(L, G) dct_fm_loss_vec_kernel<TOut, L, G><<<nctas, block, 0, st>>>( \
        out, vt, freq_w, grad, accum, losses, upstream, H, W, tiles_x, tiles_y, ntiles, per, fms, fqs, flw, 1.0 / n_fm, 1.0 / n_fq)
        if (want_loss && want_grad) DCT_VEC(true, true);
        else if (want_grad) DCT_VEC(false, true);
        else DCT_VEC(true, false);
#undef DCT_VEC
        DECO_CHECK_LAUNCH("dct_fm_loss_vec_kernel");
        return DECO_OK;
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

// doubles of scratch deco_dct_fm_loss needs: {fm, freq} accumulators of the ragged path, the ticket, padding, then one
// {fm, freq} slot per CTA of the persistent kernel
extern "C" int deco_dct_scratch_doubles(void) { return 4 + 2 * deco::kDctMaxCtas; }

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
