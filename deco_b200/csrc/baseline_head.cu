// Output head of the patch-linear baseline denoiser (FlattenDiT): AdaLN final layer and fold.  HBM-bound glue, one pass each,
// 128-bit accesses.  The blocks in front of it are the same fused tcgen05 GEMM / attention kernels as the DeCo denoiser.
//
// Replaces (reference, /root/reference/src/models/transformer/dit_c2i_baseline.py):
//   :70-83  FinalLayer: LayerNorm(no affine, eps 1e-6) + modulate(shift, scale)   -> layernorm_modulate_kernel
//           (the Linear(H, C p^2) that follows is the tcgen05 GEMM with the bias epilogue)
//   :378    F.fold(kernel = stride = p)                                          -> unpatchify_kernel
#include "common.cuh"

namespace deco {

// h = (x - mean) * rsqrt(var + eps) * (1 + scale) + shift, one warp per row of the fp32 stream, the row kept in registers
// between the statistics and the output pass (two-pass variance: no E[x^2] - E[x]^2 cancellation).  Hd % 8 == 0.
template <int kMaxChunks>
__global__ void __launch_bounds__(256) layernorm_modulate_kernel(
    const float* __restrict__ x, const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
    long long mod_row_stride, int rows_per_mod, __nv_bfloat16* __restrict__ out, long long M, int Hd, float eps)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nch = Hd >> 3;
    const float* xr = x + row * Hd;
    float v[kMaxChunks][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) {
            const uint4 a = ld_stream16(xr + ch * 8), b = ld_stream16(xr + ch * 8 + 4);
            v[j][0] = __uint_as_float(a.x); v[j][1] = __uint_as_float(a.y); v[j][2] = __uint_as_float(a.z); v[j][3] = __uint_as_float(a.w);
            v[j][4] = __uint_as_float(b.x); v[j][5] = __uint_as_float(b.y); v[j][6] = __uint_as_float(b.z); v[j][7] = __uint_as_float(b.w);
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[j][e];
        }
    }
    const float mean = warp_sum(sum) / (float)Hd;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { v[j][e] -= mean; ss = fmaf(v[j][e], v[j][e], ss); }
        }
    }
    const float rs = rsqrtf(warp_sum(ss) / (float)Hd + eps);
    const long long mrow = row / rows_per_mod;
    const uint4* shr = reinterpret_cast<const uint4*>(shift + mrow * mod_row_stride);
    const uint4* scr = reinterpret_cast<const uint4*>(scale + mrow * mod_row_stride);
    uint4* orow = reinterpret_cast<uint4*>(out + row * Hd);
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) {
            const uint4 sh = __ldg(shr + ch), sc = __ldg(scr + ch);
            const uint32_t shw[4] = {sh.x, sh.y, sh.z, sh.w}, scw[4] = {sc.x, sc.y, sc.z, sc.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 s2 = unpack_bf2(shw[e]), c2 = unpack_bf2(scw[e]);
                o[e] = pack_bf2(fmaf(v[j][2 * e] * rs, 1.0f + c2.x, s2.x), fmaf(v[j][2 * e + 1] * rs, 1.0f + c2.y, s2.y));
            }
            orow[ch] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// out = x - mean(x) per row (fp32, may run in place): LayerNorm without affine is RMSNorm of the centred row, so the
// TRAINING path of the final layer is centre -> rmsnorm_modulate (unit weight), and its backward is rmsnorm_modulate_bwd ->
// centre (the projection is symmetric).  One warp per row, the row kept in registers.
template <int kMaxChunks>
__global__ void __launch_bounds__(256) center_rows_kernel(const float* x, float* out, long long M, int Hd)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nch = Hd >> 2;
    const float4* xr = reinterpret_cast<const float4*>(x + row * Hd);
    float4 v[kMaxChunks];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) { v[j] = xr[ch]; sum += (v[j].x + v[j].y) + (v[j].z + v[j].w); }
    }
    const float mean = warp_sum(sum) / (float)Hd;
    float4* orow = reinterpret_cast<float4*>(out + row * Hd);
#pragma unroll
    for (int j = 0; j < kMaxChunks; ++j) {
        const int ch = lane + j * 32;
        if (ch < nch) orow[ch] = make_float4(v[j].x - mean, v[j].y - mean, v[j].z - mean, v[j].w - mean);
    }
}

// out[b][c][py*p+ky][px*p+kx] = tok[(b*L + py*Wp + px)][c*p*p + ky*p + kx]; p % 8 == 0.  Thread i owns output chunk i (8
// pixels of one image row): stores are fully coalesced, loads are 32-byte runs (p = 16) of a token row.
__global__ void __launch_bounds__(256) unpatchify_kernel(const __nv_bfloat16* __restrict__ tok, __nv_bfloat16* __restrict__ out,
                                                         int C, int H, int W, int p, long long total8)
{
    const int Wp = W / p, W8 = W / 8, p8 = p / 8;
    const long long L = (long long)(H / p) * Wp;
    const int feat = C * p * p;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
        long long r = i;
        const int x8 = (int)(r % W8); r /= W8;
        const int y = (int)(r % H); r /= H;
        const int c = (int)(r % C); r /= C;
        const long long b = r;
        const int px = x8 / p8, kx8 = x8 % p8, py = y / p, ky = y % p;
        const __nv_bfloat16* src = tok + (b * L + (long long)py * Wp + px) * feat + (c * p + ky) * p + kx8 * 8;
        reinterpret_cast<uint4*>(out)[i] = ld_stream16(src);
    }
}

}  // namespace deco

extern "C" int deco_layernorm_modulate(const float* x, const void* shift_bf16, const void* scale_bf16,
                                       long long mod_row_stride, int rows_per_mod, void* out_bf16, long long M, int hidden,
                                       float eps, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(x && shift_bf16 && scale_bf16 && out_bf16, "layernorm_modulate: null pointer");
    DECO_CHECK_ARG(M > 0 && hidden > 0 && hidden % 8 == 0 && hidden <= 2048 && rows_per_mod > 0 && mod_row_stride % 8 == 0,
                   "layernorm_modulate: unsupported M=%lld hidden=%d", M, hidden);
    const int warps = 8;
    const unsigned grid = (unsigned)((M + warps - 1) / warps);
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* sh = (const __nv_bfloat16*)shift_bf16;
    const __nv_bfloat16* sc = (const __nv_bfloat16*)scale_bf16;
    if (hidden <= 1280)
        layernorm_modulate_kernel<5><<<grid, warps * 32, 0, st>>>(x, sh, sc, mod_row_stride, rows_per_mod,
                                                                  (__nv_bfloat16*)out_bf16, M, hidden, eps);
    else
        layernorm_modulate_kernel<8><<<grid, warps * 32, 0, st>>>(x, sh, sc, mod_row_stride, rows_per_mod,
                                                                  (__nv_bfloat16*)out_bf16, M, hidden, eps);
    DECO_CHECK_LAUNCH("layernorm_modulate_kernel");
    return DECO_OK;
}

extern "C" int deco_center_rows(const float* x, float* out, long long M, int hidden, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(x && out && M > 0 && hidden > 0 && hidden % 4 == 0 && hidden <= 2048, "center_rows: unsupported M=%lld hidden=%d", M, hidden);
    const int warps = 8;
    const unsigned grid = (unsigned)((M + warps - 1) / warps);
    if (hidden <= 1280) center_rows_kernel<10><<<grid, warps * 32, 0, (cudaStream_t)stream>>>(x, out, M, hidden);
    else center_rows_kernel<16><<<grid, warps * 32, 0, (cudaStream_t)stream>>>(x, out, M, hidden);
    DECO_CHECK_LAUNCH("center_rows_kernel");
    return DECO_OK;
}

extern "C" int deco_unpatchify(const void* tok_bf16, void* out_bf16, int B, int C, int H, int W, int p, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(tok_bf16 && out_bf16, "unpatchify: null pointer");
    DECO_CHECK_ARG(B > 0 && C > 0 && p > 0 && p % 8 == 0 && H % p == 0 && W % p == 0,
                   "unpatchify: unsupported shape B=%d C=%d H=%d W=%d p=%d (need p%%8==0, H,W%%p==0)", B, C, H, W, p);
    const long long total8 = (long long)B * C * H * W / 8;
    long long blocks = (total8 + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    unpatchify_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)tok_bf16, (__nv_bfloat16*)out_bf16,
                                                                         C, H, W, p, total8);
    DECO_CHECK_LAUNCH("unpatchify_kernel");
    return DECO_OK;
}
