// Memory-bound glue kernels of the denoiser forward (all HBM-bound; 128-bit accesses, one pass each).
//
// Replaces (reference, /root/reference/src/models/transformer/dit_c2i_DeCo.py):
//   :491 F.unfold + transpose            -> patchify_kernel        (fp32 NCHW -> bf16 [B*L, C*p*p])
//   :43-53 timestep_embedding            -> timestep_freq_kernel   (cos || sin, max_period 10)
//   :493-494 y_embedder + silu(t + y)    -> cond_combine_kernel
//   :94-99 RMSNorm + :11-12 modulate     -> rmsnorm_modulate_kernel
//   :178-180 q_norm / k_norm + RoPE      -> qknorm_rope_kernel     (in place on the QKV GEMM output)
//   :499 silu(t + s)                     -> silu_add_rows_kernel
#include "common.cuh"

namespace deco {

// ---------------------------------------------------------------- patchify
// out[(b*L + py*Wp + px)][c*p*p + ky*p + kx] = x[b][c][py*p+ky][px*p+kx]; p % 8 == 0
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                       int C, int H, int W, int p, long long total8)
{
    const int Wp = W / p, Hp = H / p;
    const int p8 = p / 8;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
        long long r = i;
        const int kx8 = (int)(r % p8); r /= p8;
        const int ky = (int)(r % p); r /= p;
        const int c = (int)(r % C); r /= C;
        const int px = (int)(r % Wp); r /= Wp;
        const int py = (int)(r % Hp); r /= Hp;
        const long long b = r;
        const float* src = x + (((b * C + c) * H + (py * p + ky)) * (long long)W + px * p + kx8 * 8);
        const float4 a = __ldg(reinterpret_cast<const float4*>(src));
        const float4 d = __ldg(reinterpret_cast<const float4*>(src) + 1);
        uint4 o;
        o.x = pack_bf2(a.x, a.y); o.y = pack_bf2(a.z, a.w); o.z = pack_bf2(d.x, d.y); o.w = pack_bf2(d.z, d.w);
        reinterpret_cast<uint4*>(out)[i] = o;   // i enumerates output 16-byte chunks in order
    }
}

// ---------------------------------------------------------------- timestep sinusoid
__global__ void timestep_freq_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int B, int dim,
                                     float max_period)
{
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, k = i % half;
    // fp32 evaluation order of the reference: exp(-log(max_period) * k / half), then t * freq
    const float freq = expf(-logf(max_period) * (float)k / (float)half);
    const float arg = t[b] * freq;
    out[(size_t)b * dim + k] = f2bf(cosf(arg));
    out[(size_t)b * dim + half + k] = f2bf(sinf(arg));
}

// ---------------------------------------------------------------- c = silu(t_emb + y_emb[label])
__global__ void cond_combine_kernel(const __nv_bfloat16* __restrict__ temb, const float* __restrict__ table,
                                    const long long* __restrict__ labels, __nv_bfloat16* __restrict__ c,
                                    int B, int Hd, int num_rows)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Hd) return;
    const int b = i / Hd, h = i % Hd;
    long long lab = labels[b];
    if (lab < 0 || lab >= num_rows) lab = num_rows - 1;   // clamp: the reference would raise on the host
    const float v = bf2f(temb[i]) + __ldg(table + lab * Hd + h);
    c[i] = f2bf(silu_f(v));
}

// ---------------------------------------------------------------- h = rms(x) * w * (1 + scale) + shift
// One warp per row; the row lives in registers between the two passes.  Hd % 8 == 0, Hd <= 2048.
// TIn = float: fp32 residual stream (no intermediate rounding); TIn = bf16: the reference's rounding points
// (normalised value cast back to bf16, dit_c2i_DeCo.py:99; (1 + scale) evaluated as a bf16 op).
template <typename TIn> __device__ __forceinline__ void load8(const TIn* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 q = ld_stream16(p);
    const float2 a = unpack_bf2(q.x), b = unpack_bf2(q.y), c = unpack_bf2(q.z), d = unpack_bf2(q.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const uint4 a = ld_stream16(p), b = ld_stream16(p + 4);
    v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
    v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
}

template <typename TIn, int kMaxChunks>
__global__ void __launch_bounds__(256) rmsnorm_modulate_kernel(
    const TIn* __restrict__ x, const float* __restrict__ w,
    const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale, long long mod_row_stride,
    int rows_per_mod, __nv_bfloat16* __restrict__ out, long long M, int Hd, float eps)
{
    constexpr bool kRefRounding = sizeof(TIn) == 2;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nch = Hd >> 3;
    const TIn* xr = x + row * Hd;
    float v[kMaxChunks][8];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) {
            load8<TIn>(xr + ch * 8, v[j]);
#pragma unroll
            for (int e = 0; e < 8; ++e) ss = fmaf(v[j][e], v[j][e], ss);
        }
    }
    ss = warp_sum(ss);
    const float rs = rsqrtf(ss / (float)Hd + eps);
    const long long mrow = row / rows_per_mod;
    const uint4* shr = reinterpret_cast<const uint4*>(shift + mrow * mod_row_stride);
    const uint4* scr = reinterpret_cast<const uint4*>(scale + mrow * mod_row_stride);
    uint4* orow = reinterpret_cast<uint4*>(out + row * Hd);
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + ch * 2);
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w) + ch * 2 + 1);
            const uint4 sh = __ldg(shr + ch), sc = __ldg(scr + ch);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const uint32_t shw[4] = {sh.x, sh.y, sh.z, sh.w}, scw[4] = {sc.x, sc.y, sc.z, sc.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 s2 = unpack_bf2(shw[e]), c2 = unpack_bf2(scw[e]);
                float n0 = v[j][2 * e] * rs, n1 = v[j][2 * e + 1] * rs;
                float m0 = 1.0f + c2.x, m1 = 1.0f + c2.y;
                if (kRefRounding) { n0 = round_bf(n0); n1 = round_bf(n1); m0 = round_bf(m0); m1 = round_bf(m1); }
                o[e] = pack_bf2(fmaf(wv[2 * e] * n0, m0, s2.x), fmaf(wv[2 * e + 1] * n1, m1, s2.y));
            }
            orow[ch] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---------------------------------------------------------------- gated residual + the next RMSNorm / modulation in one pass
//   s_out = s + gate[b] * a   (fp32 stream)   and   h = rms(s_out) * w * (1 + scale[b]) + shift[b]   (bf16)
// (training forward: dit_c2i_DeCo.py:236-244, the residual add of one branch followed by the norm that feeds the next
// GEMM).  A warp owns a row and keeps it in registers between the two halves, so s_out is written once and never re-read.
// Same arithmetic, in the same order, as gate_residual_kernel followed by rmsnorm_modulate_kernel<float>.
template <int kMaxChunks>
__global__ void __launch_bounds__(256) gate_residual_norm_kernel(
    const float* s, const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ gate, long long gate_stride,
    float* s_out, const float* __restrict__ w, const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
    long long mod_row_stride, int rows_per_mod, __nv_bfloat16* __restrict__ h_out, long long M, int Hd, float eps)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nch = Hd >> 3;
    const long long mrow = row / rows_per_mod;
    const uint4* grow = reinterpret_cast<const uint4*>(gate + mrow * gate_stride);
    float v[kMaxChunks][8];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) {
            float av[8];
            load8<float>(s + row * Hd + ch * 8, v[j]);
            load8<__nv_bfloat16>(a + row * Hd + ch * 8, av);
            const uint4 g4 = __ldg(grow + ch);
            const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 g2 = unpack_bf2(gw[e]);
                v[j][2 * e] = fmaf(g2.x, av[2 * e], v[j][2 * e]);
                v[j][2 * e + 1] = fmaf(g2.y, av[2 * e + 1], v[j][2 * e + 1]);
            }
            float4* op = reinterpret_cast<float4*>(s_out + row * Hd + ch * 8);
            op[0] = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
            op[1] = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
#pragma unroll
            for (int e = 0; e < 8; ++e) ss = fmaf(v[j][e], v[j][e], ss);
        }
    }
    ss = warp_sum(ss);
    const float rs = rsqrtf(ss / (float)Hd + eps);
    const uint4* shr = reinterpret_cast<const uint4*>(shift + mrow * mod_row_stride);
    const uint4* scr = reinterpret_cast<const uint4*>(scale + mrow * mod_row_stride);
    uint4* orow = reinterpret_cast<uint4*>(h_out + row * Hd);
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + ch * 2);
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w) + ch * 2 + 1);
            const uint4 sh = __ldg(shr + ch), sc = __ldg(scr + ch);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const uint32_t shw[4] = {sh.x, sh.y, sh.z, sh.w}, scw[4] = {sc.x, sc.y, sc.z, sc.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 s2 = unpack_bf2(shw[e]), c2 = unpack_bf2(scw[e]);
                const float n0 = v[j][2 * e] * rs, n1 = v[j][2 * e + 1] * rs;
                o[e] = pack_bf2(fmaf(wv[2 * e] * n0, 1.0f + c2.x, s2.x), fmaf(wv[2 * e + 1] * n1, 1.0f + c2.y, s2.y));
            }
            orow[ch] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// ---------------------------------------------------------------- per-head RMSNorm + 2-D RoPE, in place
// buf: [M, row_stride]; segment 0 starts at column col0 (weights qw), optional segment 1 at col1 (weights kw); each
// segment is heads x D.  rope: [L, D/2] float2 (cos, sin) or NULL (norm only -- the t2i text keys); token position = row % L.
// G lanes share one (token, segment, head) vector (3 x 24 elements for D = 72, 4 x 16 for D = 64): a lane holds its 16-byte
// chunks packed, the row's sum of squares is a G-lane shuffle reduction.  (The first version gave a thread the whole vector
// in 72 fp32 registers: 0.53-0.57 of the HBM peak; with 3-4x the warps in flight the same traffic hides its latency.)
template <int D>
__global__ void __launch_bounds__(256) qknorm_rope_kernel(const __nv_bfloat16* src, __nv_bfloat16* qkv, long long row_stride,
                                                          int nseg, int col0, int col1, const float* __restrict__ qw,
                                                          const float* __restrict__ kw, const float2* __restrict__ rope,
                                                          long long M, int heads, int L, float eps)
{
    constexpr int G = (D == 72) ? 3 : 4;
    constexpr int CPL = D / 8 / G;            // 16-byte chunks per lane
    constexpr int VPW = 32 / G;               // vectors per warp (D = 72: lanes 30, 31 idle)
    static_assert(D % (8 * G) == 0, "head_dim must split into 16-byte chunks per lane");
    __shared__ __align__(16) float sw[2][D];
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) sw[i / D][i % D] = (i < D) ? qw[i] : (nseg > 1 ? kw[i - D] : 0.f);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G, sub = lane % G, base = grp * G;
    const int per_tok = nseg * heads;
    const long long total = M * per_tok;
    const long long item = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * VPW + grp;
    const bool live = grp < VPW && item < total;
    const long long tok = live ? item / per_tok : 0;
    const int r = live ? (int)(item % per_tok) : 0;
    const int is_k = r / heads, head = r % heads;
    const long long off = tok * row_stride + (is_k ? col1 : col0) + (long long)head * D;
    const float2* rp = rope ? rope + (long long)(tok % L) * (D / 2) : nullptr;
    uint4 rq[CPL];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        rq[k] = live ? *reinterpret_cast<const uint4*>(src + off + (sub + G * k) * 8) : make_uint4(0u, 0u, 0u, 0u);   // src == qkv: in place
        const uint32_t w4[4] = {rq[k].x, rq[k].y, rq[k].z, rq[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 a = unpack_bf2(w4[e]); ss = fmaf(a.x, a.x, fmaf(a.y, a.y, ss)); }
    }
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < G; ++j) tot += __shfl_sync(0xffffffffu, ss, (base + j) & 31);
    const float rs = rsqrtf(tot / (float)D + eps);
    if (!live) return;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int ch = sub + G * k;
        const uint32_t w4[4] = {rq[k].x, rq[k].y, rq[k].z, rq[k].w};
        float4 t01 = make_float4(1.f, 0.f, 1.f, 0.f), t23 = t01;
        if (rp) { t01 = __ldg(reinterpret_cast<const float4*>(rp + ch * 4)); t23 = __ldg(reinterpret_cast<const float4*>(rp + ch * 4 + 2)); }
        const float4 w03 = *reinterpret_cast<const float4*>(&sw[is_k][ch * 8]), w47 = *reinterpret_cast<const float4*>(&sw[is_k][ch * 8 + 4]);
        const float cc[4] = {t01.x, t01.z, t23.x, t23.z}, sn[4] = {t01.y, t01.w, t23.y, t23.w};
        const float ww[8] = {w03.x, w03.y, w03.z, w03.w, w47.x, w47.y, w47.z, w47.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 v = unpack_bf2(w4[e]);
            const float a = ww[2 * e] * round_bf(v.x * rs), b = ww[2 * e + 1] * round_bf(v.y * rs);
            o[e] = pack_bf2(a * cc[e] - b * sn[e], a * sn[e] + b * cc[e]);
        }
        *reinterpret_cast<uint4*>(qkv + off + ch * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------- out[m, :] = silu(x[m, :] + row[m / rows_per][:])
template <typename TIn>
__global__ void __launch_bounds__(256) silu_add_rows_kernel(const TIn* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ rowv,
                                                            __nv_bfloat16* __restrict__ out, long long M, int Hd,
                                                            int rows_per)
{
    const int nch = Hd >> 3;
    const long long total = M * nch;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long m = i / nch;
        const int ch = (int)(i % nch);
        float a[8];
        if (sizeof(TIn) == 2) {
            const uint4 q = reinterpret_cast<const uint4*>(x)[i];   // plain load: out may alias x
            const float2 p0 = unpack_bf2(q.x), p1 = unpack_bf2(q.y), p2 = unpack_bf2(q.z), p3 = unpack_bf2(q.w);
            a[0] = p0.x; a[1] = p0.y; a[2] = p1.x; a[3] = p1.y; a[4] = p2.x; a[5] = p2.y; a[6] = p3.x; a[7] = p3.y;
        } else {
            const float4 q0 = reinterpret_cast<const float4*>(x)[2 * i], q1 = reinterpret_cast<const float4*>(x)[2 * i + 1];
            a[0] = q0.x; a[1] = q0.y; a[2] = q0.z; a[3] = q0.w; a[4] = q1.x; a[5] = q1.y; a[6] = q1.z; a[7] = q1.w;
        }
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(rowv + (m / rows_per) * Hd) + ch);
        const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 q = unpack_bf2(bw[e]);
            o[e] = pack_bf2(silu_f(a[2 * e] + q.x), silu_f(a[2 * e + 1] + q.y));
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------- text embed: y = w * rms(x) + pos[m % T]
// Embed(norm_layer=RMSNorm) + learned position (dit_t2i_pixnerd.py:280; layers/patch_embed.py:19-22).  fp32 in/out.
template <int kMaxChunks>
__global__ void __launch_bounds__(256) rmsnorm_addpos_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ pos, int T,
                                                             float* __restrict__ out, long long M, int Hd, float eps)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nch = Hd >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + row * Hd);
    float4 v[kMaxChunks];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) {
            v[j] = xr[ch];
            ss = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, fmaf(v[j].z, v[j].z, fmaf(v[j].w, v[j].w, ss))));
        }
    }
    ss = warp_sum(ss);
    const float rs = rsqrtf(ss / (float)Hd + eps);
    const float4* pr = reinterpret_cast<const float4*>(pos + (row % T) * Hd);
    float4* orow = reinterpret_cast<float4*>(out + row * Hd);
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) {
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + ch), pp = __ldg(pr + ch);
            orow[ch] = make_float4(fmaf(ww.x, v[j].x * rs, pp.x), fmaf(ww.y, v[j].y * rs, pp.y),
                                   fmaf(ww.z, v[j].z * rs, pp.z), fmaf(ww.w, v[j].w * rs, pp.w));
        }
    }
}

// ---------------------------------------------------------------- fp32 -> bf16 copy (text stream -> kv_y GEMM operand)
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                            long long total8)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
        const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
        reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf2(a.x, a.y), pack_bf2(a.z, a.w), pack_bf2(b.x, b.y),
                                                      pack_bf2(b.z, b.w));
    }
}

static inline unsigned grid_for(long long work, int threads, int per_sm = 16) {
    long long b = (work + threads - 1) / threads;
    const long long cap = (long long)kNumSMs * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace deco

extern "C" int deco_patchify(const float* x, void* out_bf16, int B, int C, int H, int W, int p, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(x && out_bf16, "patchify: null pointer");
    DECO_CHECK_ARG(B > 0 && C > 0 && p > 0 && p % 8 == 0 && H % p == 0 && W % p == 0,
                   "patchify: unsupported shape B=%d C=%d H=%d W=%d p=%d (need p%%8==0, H,W%%p==0)", B, C, H, W, p);
    const long long total8 = (long long)B * C * H * W / 8;
    patchify_kernel<<<grid_for(total8, 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, C, H, W, p, total8);
    DECO_CHECK_LAUNCH("patchify_kernel");
    return DECO_OK;
}

extern "C" int deco_timestep_freq(const float* t, void* out_bf16, int B, int dim, float max_period, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(t && out_bf16 && B > 0 && dim > 0 && dim % 2 == 0, "timestep_freq: bad arguments");
    const int n = B * (dim / 2);
    timestep_freq_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, (__nv_bfloat16*)out_bf16, B, dim, max_period);
    DECO_CHECK_LAUNCH("timestep_freq_kernel");
    return DECO_OK;
}

extern "C" int deco_cond_combine(const void* temb_bf16, const float* table, const long long* labels, void* c_bf16,
                                 int B, int hidden, int num_rows, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(temb_bf16 && table && labels && c_bf16 && B > 0 && hidden > 0 && num_rows > 0,
                   "cond_combine: bad arguments");
    const int n = B * hidden;
    cond_combine_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)temb_bf16, table, labels, (__nv_bfloat16*)c_bf16, B, hidden, num_rows);
    DECO_CHECK_LAUNCH("cond_combine_kernel");
    return DECO_OK;
}

extern "C" int deco_rmsnorm_modulate(const void* x, int x_is_f32, const float* weight, const void* shift_bf16,
                                     const void* scale_bf16, long long mod_row_stride, int rows_per_mod,
                                     void* out_bf16, long long M, int hidden, float eps, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(x && weight && shift_bf16 && scale_bf16 && out_bf16, "rmsnorm_modulate: null pointer");
    DECO_CHECK_ARG(M > 0 && hidden % 8 == 0 && hidden <= 2048 && rows_per_mod > 0 && mod_row_stride % 8 == 0,
                   "rmsnorm_modulate: unsupported M=%lld hidden=%d", M, hidden);
    const int warps = 8;
    const unsigned grid = (unsigned)((M + warps - 1) / warps);
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* sh = (const __nv_bfloat16*)shift_bf16;
    const __nv_bfloat16* sc = (const __nv_bfloat16*)scale_bf16;
    __nv_bfloat16* o = (__nv_bfloat16*)out_bf16;
#define RMS_LAUNCH(T, C) rmsnorm_modulate_kernel<T, C><<<grid, warps * 32, 0, st>>>( \
        (const T*)x, weight, sh, sc, mod_row_stride, rows_per_mod, o, M, hidden, eps)
    if (x_is_f32) { if (hidden <= 1280) RMS_LAUNCH(float, 5); else RMS_LAUNCH(float, 8); }
    else { if (hidden <= 1280) RMS_LAUNCH(__nv_bfloat16, 5); else RMS_LAUNCH(__nv_bfloat16, 8); }
#undef RMS_LAUNCH
    DECO_CHECK_LAUNCH("rmsnorm_modulate_kernel");
    return DECO_OK;
}

extern "C" int deco_gate_residual_norm(const float* s, const void* a_bf16, const void* gate_bf16, long long gate_stride,
                                       float* s_out, const float* weight, const void* shift_bf16, const void* scale_bf16,
                                       long long mod_row_stride, int rows_per_image, void* h_out_bf16, long long M, int hidden,
                                       float eps, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(s && a_bf16 && gate_bf16 && s_out && weight && shift_bf16 && scale_bf16 && h_out_bf16, "gate_residual_norm: null pointer");
    DECO_CHECK_ARG(M > 0 && hidden % 8 == 0 && hidden <= 2048 && rows_per_image > 0 && mod_row_stride % 8 == 0 && gate_stride % 8 == 0,
                   "gate_residual_norm: unsupported M=%lld hidden=%d", M, hidden);
    DECO_CHECK_ARG((((uintptr_t)s | (uintptr_t)a_bf16 | (uintptr_t)gate_bf16 | (uintptr_t)s_out | (uintptr_t)weight | (uintptr_t)shift_bf16 |
                     (uintptr_t)scale_bf16 | (uintptr_t)h_out_bf16) & 15) == 0, "gate_residual_norm: pointers must be 16-byte aligned");
    const int warps = 8;
    const unsigned grid = (unsigned)((M + warps - 1) / warps);
    cudaStream_t st = (cudaStream_t)stream;
#define GRN_LAUNCH(C) gate_residual_norm_kernel<C><<<grid, warps * 32, 0, st>>>( \
        s, (const __nv_bfloat16*)a_bf16, (const __nv_bfloat16*)gate_bf16, gate_stride, s_out, weight, (const __nv_bfloat16*)shift_bf16, \
        (const __nv_bfloat16*)scale_bf16, mod_row_stride, rows_per_image, (__nv_bfloat16*)h_out_bf16, M, hidden, eps)
    if (hidden <= 1280) GRN_LAUNCH(5); else GRN_LAUNCH(8);
#undef GRN_LAUNCH
    DECO_CHECK_LAUNCH("gate_residual_norm_kernel");
    return DECO_OK;
}

static int headnorm_rope_launch(const void* src_bf16, void* buf_bf16, long long row_stride, int nseg, int col0, int col1,
                                const float* w0, const float* w1, const float* rope_cos_sin,
                                long long M, int heads, int head_dim, int L, float eps, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(src_bf16 && buf_bf16 && w0 && (nseg == 1 || (nseg == 2 && w1)), "headnorm_rope: null pointer / bad nseg");
    DECO_CHECK_ARG(M > 0 && heads > 0 && L > 0 && row_stride % 8 == 0 && col0 % 8 == 0 && col1 % 8 == 0 && col0 >= 0 && col1 >= 0,
                   "headnorm_rope: bad shape");
    const long long items = M * nseg * heads;
    const int vpb = 8 * (32 / (head_dim == 72 ? 3 : 4));       // vectors per 256-thread block
    const unsigned grid = (unsigned)((items + vpb - 1) / vpb);
    DECO_CHECK_ARG(!rope_cos_sin || ((uintptr_t)rope_cos_sin & 15) == 0, "headnorm_rope: rope table must be 16-byte aligned");
    const __nv_bfloat16* sp = (const __nv_bfloat16*)src_bf16;
    if (head_dim == 72)
        qknorm_rope_kernel<72><<<grid, 256, 0, (cudaStream_t)stream>>>(sp, (__nv_bfloat16*)buf_bf16, row_stride, nseg, col0, col1,
                                                                      w0, w1, (const float2*)rope_cos_sin, M, heads, L, eps);
    else if (head_dim == 64)
        qknorm_rope_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(sp, (__nv_bfloat16*)buf_bf16, row_stride, nseg, col0, col1,
                                                                      w0, w1, (const float2*)rope_cos_sin, M, heads, L, eps);
    else {
        deco_set_error("headnorm_rope: head_dim %d not built (64, 72)", head_dim);
        return DECO_ERR_UNSUPPORTED;
    }
    DECO_CHECK_LAUNCH("qknorm_rope_kernel");
    return DECO_OK;
}

extern "C" int deco_headnorm_rope(void* buf_bf16, long long row_stride, int nseg, int col0, int col1,
                                  const float* w0, const float* w1, const float* rope_cos_sin,
                                  long long M, int heads, int head_dim, int L, float eps, void* stream) {
    return headnorm_rope_launch(buf_bf16, buf_bf16, row_stride, nseg, col0, col1, w0, w1, rope_cos_sin, M, heads, head_dim, L,
                                eps, stream);
}

// out-of-place form: reads the segments from src, writes them to dst (same row stride and columns); the training forward
// keeps the raw QKV GEMM output for the backward of the norm and gets the normalised copy without a memcpy
extern "C" int deco_headnorm_rope_to(const void* src_bf16, void* dst_bf16, long long row_stride, int nseg, int col0, int col1,
                                     const float* w0, const float* w1, const float* rope_cos_sin,
                                     long long M, int heads, int head_dim, int L, float eps, void* stream) {
    return headnorm_rope_launch(src_bf16, dst_bf16, row_stride, nseg, col0, col1, w0, w1, rope_cos_sin, M, heads, head_dim, L,
                                eps, stream);
}

extern "C" int deco_qknorm_rope(void* qkv_bf16, const float* q_weight, const float* k_weight, const float* rope_cos_sin,
                                long long M, int heads, int head_dim, int L, float eps, void* stream) {
    if (!rope_cos_sin) { deco_set_error("qknorm_rope: null pointer"); return DECO_ERR_ARG; }
    return deco_headnorm_rope(qkv_bf16, 3LL * heads * head_dim, 2, 0, heads * head_dim, q_weight, k_weight, rope_cos_sin,
                              M, heads, head_dim, L, eps, stream);
}

extern "C" int deco_rmsnorm_addpos(const float* x, const float* weight, const float* pos, int T, float* out,
                                   long long M, int hidden, float eps, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(x && weight && pos && out && M > 0 && T > 0 && hidden % 4 == 0 && hidden <= 2048,
                   "rmsnorm_addpos: bad arguments (hidden %% 4 == 0, <= 2048)");
    const int warps = 8;
    const unsigned grid = (unsigned)((M + warps - 1) / warps);
    rmsnorm_addpos_kernel<16><<<grid, warps * 32, 0, (cudaStream_t)stream>>>(x, weight, pos, T, out, M, hidden, eps);
    DECO_CHECK_LAUNCH("rmsnorm_addpos_kernel");
    return DECO_OK;
}

extern "C" int deco_cast_f32_bf16(const float* x, void* out_bf16, long long n, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(x && out_bf16 && n > 0 && n % 8 == 0, "cast_f32_bf16: n must be a positive multiple of 8");
    cast_f32_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, n / 8);
    DECO_CHECK_LAUNCH("cast_f32_bf16_kernel");
    return DECO_OK;
}

extern "C" int deco_silu_add_rows(const void* x, int x_is_f32, const void* row_bf16, void* out_bf16, long long M,
                                  int hidden, int rows_per, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(x && row_bf16 && out_bf16 && M > 0 && hidden % 8 == 0 && rows_per > 0, "silu_add_rows: bad arguments");
    const long long total = M * (hidden / 8);
    if (x_is_f32)
        silu_add_rows_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
            (const float*)x, (const __nv_bfloat16*)row_bf16, (__nv_bfloat16*)out_bf16, M, hidden, rows_per);
    else
        silu_add_rows_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)x, (const __nv_bfloat16*)row_bf16, (__nv_bfloat16*)out_bf16, M, hidden, rows_per);
    DECO_CHECK_LAUNCH("silu_add_rows_kernel");
    return DECO_OK;
}
