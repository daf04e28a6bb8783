// Per-pixel AdaLN-MLP pixel decoder: NerfEmbedder + input_proj + R x ResBlock + final LayerNorm/Linear,
// patchify of the fp32 input and fold of the output fused in.  The 8192-wide per-token condition
// (cond_embed GEMM output) is read once with 128-bit coalesced loads; every activation of the MLP chain
// stays in registers: the m16n8k16 accumulator layout of one layer is the A-operand layout of the next,
// LayerNorm statistics are quad shuffles.
//
// Replaces (reference, /root/reference/src/models/transformer/dit_c2i_DeCo.py):
//   :212-248 NerfEmbedder (constant pos table folded into T = W_pos . table + b on the host)
//   :395-415 SimpleMLPAdaLN.forward, :313-317 ResBlock.forward, :329-332 decoder FinalLayer
//   :501-509 pixel reshape / transpose / F.fold
//
// One warp = one 16-pixel patch row (ky) of one token; lane (g = lane/4, t = lane%4) owns pixel rows g and g+8
// of the MMA tile and channels {8j + 2t, 8j + 2t + 1}, j = 0..3.
// Weights live in shared memory in B-fragment order (one conflict-free 8-byte load per MMA); the host packs
// them (deco_b200/denoiser.py::_pack_decoder).  The adaLN weights use a K permutation so that a lane's
// 16-byte chunk of the condition (channels 8t..8t+7) is directly its A fragment.
//
// Bound (fused): the 64 B/pixel condition read, 12 B/pixel input, 6 B/pixel output vs 37 kFLOP/pixel of
// K=32 MMAs -> tensor(legacy-pipe)/issue bound; see DESIGN.md.
#include "common.cuh"

namespace deco {

constexpr int kHx = 32;                  // decoder width
constexpr int kFragWords = 64;           // uint2 per (n-tile, k-step): 32 lanes -> 64 uint32
constexpr int kFrag32 = 4 * 2 * kFragWords;    // 32->32 layer: 4 n-tiles x 2 k-steps, in uint32
constexpr int kFrag96 = 12 * 2 * kFragWords;   // 32->96 layer
constexpr int kFrag8 = 1 * 2 * kFragWords;     // 32->8 (final, 3 valid outputs)
constexpr int kBlockFrag = kFrag96 + 2 * kFrag32;
constexpr int kVecRgb = 0;               // [32][3]
constexpr int kVecBin = 96;              // [32]
constexpr int kVecBlock0 = 128;          // per block: bada[96] lng[32] lnb[32] b0[32] b2[32]
constexpr int kVecPerBlock = 224;

__host__ __device__ inline int dec_frag_words(int R) { return kFrag32 + R * kBlockFrag + kFrag8; }
__host__ __device__ inline int dec_vec_floats(int R) { return kVecBlock0 + R * kVecPerBlock + 8; }

__device__ __forceinline__ void quad_sum2(float& a, float& b) {
    a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2);
}

// acc[j][.] (+)= A(16x32) . W^T for n-tiles [j0, j0+NT)
template <int NT>
__device__ __forceinline__ void mlp_mma(float (&acc)[NT][4], const uint32_t (&a)[2][4], const uint32_t* sW, int j0, int lane) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const uint2 b = *reinterpret_cast<const uint2*>(sW + ((j0 + j) * 2 + s) * kFragWords + lane * 2);
            const uint32_t bb[2] = {b.x, b.y};
            mma_bf16_16816(acc[j], a[s], bb);
        }
    }
}

// accumulator tile (16x32 fp32, 4 n-tiles) -> A fragments of the next layer
__device__ __forceinline__ void acc_to_afrag(const float (&v)[4][4], uint32_t (&a)[2][4]) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        a[s][0] = pack_bf2(v[2 * s][0], v[2 * s][1]);
        a[s][1] = pack_bf2(v[2 * s][2], v[2 * s][3]);
        a[s][2] = pack_bf2(v[2 * s + 1][0], v[2 * s + 1][1]);
        a[s][3] = pack_bf2(v[2 * s + 1][2], v[2 * s + 1][3]);
    }
}

// LayerNorm statistics over the 32 channels of pixel rows g (e = 0,1) and g+8 (e = 2,3)
__device__ __forceinline__ void ln_stats(const float (&x)[4][4], float& m0, float& r0, float& m1, float& r1) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { s0 += x[j][0] + x[j][1]; s1 += x[j][2] + x[j][3]; }
    quad_sum2(s0, s1);
    m0 = s0 * (1.0f / kHx); m1 = s1 * (1.0f / kHx);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float d;
        d = x[j][0] - m0; q0 = fmaf(d, d, q0); d = x[j][1] - m0; q0 = fmaf(d, d, q0);
        d = x[j][2] - m1; q1 = fmaf(d, d, q1); d = x[j][3] - m1; q1 = fmaf(d, d, q1);
    }
    quad_sum2(q0, q1);
    r0 = rsqrtf(q0 * (1.0f / kHx) + 1e-6f);
    r1 = rsqrtf(q1 * (1.0f / kHx) + 1e-6f);
}

struct DecParams {
    const float* x;              // [B, 3, H, W] fp32
    const __nv_bfloat16* ycond;  // [M, p*p*32]
    const uint32_t* blob;        // frag words then vec floats
    const float* postab;         // [p*p][32] fp32 : W_pos . dct + bias
    void* out;                   // [B, 3, H, W]
    int R, H, W, Hp, Wp;
    long long M;                 // tokens
};

template <typename TOut>
__global__ void __launch_bounds__(256, 2) pixel_decoder_kernel(DecParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int nfrag = dec_frag_words(P.R), nvec = dec_vec_floats(P.R);
    {
        const uint4* src = reinterpret_cast<const uint4*>(P.blob);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        const int n16 = (nfrag + nvec) / 4;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);   // weights: constants, not a predecessor's output
    }
    __syncthreads();
    pdl_wait();                   // the weight staging above overlapped the cond_embed GEMM's tail; ycond is visible from here
    pdl_launch_dependents();
    const uint32_t* sW = smem;
    const float* sV = reinterpret_cast<const float*>(smem + nfrag);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int L = P.Hp * P.Wp;
    const long long total = P.M * 16;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    const size_t plane = (size_t)P.H * P.W;

    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp; w < total; w += wstride) {
        const long long m = w >> 4;
        const int ky = (int)(w & 15);
        const long long b = m / L;
        const int tok = (int)(m % L);
        const int py = tok / P.Wp, px = tok % P.Wp;

        // ---- loads: condition chunks (16 B each) and the three colour samples of pixels g and g+8
        const __nv_bfloat16* yrow = P.ycond + (m * 256 + ky * 16) * kHx;
        const uint4 y0 = ld_stream16(yrow + g * kHx + 8 * t);
        const uint4 y1 = ld_stream16(yrow + (g + 8) * kHx + 8 * t);
        const size_t pix0 = (size_t)(py * 16 + ky) * P.W + px * 16;
        const float* xb = P.x + (size_t)b * 3 * plane + pix0;
        float rgb[2][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            rgb[0][c] = round_bf(__ldg(xb + c * plane + g));
            rgb[1][c] = round_bf(__ldg(xb + c * plane + g + 8));
        }

        // ---- silu(y) as A fragments (K-permuted: word q of the chunk <-> k-step q/2, half q%2)
        uint32_t ay[2][4];
        {
            const uint32_t w0[4] = {y0.x, y0.y, y0.z, y0.w}, w1[4] = {y1.x, y1.y, y1.z, y1.w};
            uint32_t s0[4], s1[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float2 a = unpack_bf2(w0[q]), c = unpack_bf2(w1[q]);
                s0[q] = pack_bf2(silu_f(a.x), silu_f(a.y));
                s1[q] = pack_bf2(silu_f(c.x), silu_f(c.y));
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) { ay[s][0] = s0[2 * s]; ay[s][1] = s1[2 * s]; ay[s][2] = s0[2 * s + 1]; ay[s][3] = s1[2 * s + 1]; }
        }

        // ---- NerfEmbedder: x0 = W_rgb . rgb + T[pixel]  (Linear output rounded to bf16 by the A-fragment pack)
        float xr[4][4];
        {
            const float* T0 = P.postab + (size_t)(ky * 16 + g) * kHx;
            const float* T1 = T0 + 8 * kHx;
            float x0[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = 8 * j + 2 * t;
                const float2 ta = __ldg(reinterpret_cast<const float2*>(T0 + col));
                const float2 tb = __ldg(reinterpret_cast<const float2*>(T1 + col));
                const float* wa = sV + kVecRgb + col * 3;
                x0[j][0] = ta.x + wa[0] * rgb[0][0] + wa[1] * rgb[0][1] + wa[2] * rgb[0][2];
                x0[j][1] = ta.y + wa[3] * rgb[0][0] + wa[4] * rgb[0][1] + wa[5] * rgb[0][2];
                x0[j][2] = tb.x + wa[0] * rgb[1][0] + wa[1] * rgb[1][1] + wa[2] * rgb[1][2];
                x0[j][3] = tb.y + wa[3] * rgb[1][0] + wa[4] * rgb[1][1] + wa[5] * rgb[1][2];
            }
            uint32_t a0[2][4];
            acc_to_afrag(x0, a0);
            // ---- input_proj
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(sV + kVecBin + 8 * j + 2 * t);
                xr[j][0] = bv.x; xr[j][1] = bv.y; xr[j][2] = bv.x; xr[j][3] = bv.y;
            }
            mlp_mma<4>(xr, a0, sW, 0, lane);
        }

        // ---- residual AdaLN-MLP blocks
        for (int rb = 0; rb < P.R; ++rb) {
            const uint32_t* wAda = sW + kFrag32 + rb * kBlockFrag;
            const uint32_t* w0 = wAda + kFrag96;
            const uint32_t* w2 = w0 + kFrag32;
            const float* vB = sV + kVecBlock0 + rb * kVecPerBlock;   // bada[96] lng[32] lnb[32] b0[32] b2[32]

            float ss[8][4];   // shift (n-tiles 0-3), scale (4-7)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 8 * j + 2 * t);
                ss[j][0] = bv.x; ss[j][1] = bv.y; ss[j][2] = bv.x; ss[j][3] = bv.y;
            }
            mlp_mma<8>(ss, ay, wAda, 0, lane);

            float m0, r0, m1, r1;
            ln_stats(xr, m0, r0, m1, r1);
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
            uint32_t ah[2][4];
            acc_to_afrag(h, ah);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 160 + 8 * j + 2 * t);
                h[j][0] = bv.x; h[j][1] = bv.y; h[j][2] = bv.x; h[j][3] = bv.y;
            }
            mlp_mma<4>(h, ah, w0, 0, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) h[j][e] = silu_f(h[j][e]);
            acc_to_afrag(h, ah);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 192 + 8 * j + 2 * t);
                h[j][0] = bv.x; h[j][1] = bv.y; h[j][2] = bv.x; h[j][3] = bv.y;
            }
            mlp_mma<4>(h, ah, w2, 0, lane);

            float gt[4][4];   // gate (n-tiles 8-11)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 bv = *reinterpret_cast<const float2*>(vB + 64 + 8 * j + 2 * t);
                gt[j][0] = bv.x; gt[j][1] = bv.y; gt[j][2] = bv.x; gt[j][3] = bv.y;
            }
            mlp_mma<4>(gt, ay, wAda, 8, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) xr[j][e] = fmaf(gt[j][e], h[j][e], xr[j][e]);
        }

        // ---- final LayerNorm (no affine) + Linear 32 -> 3
        {
            float m0, r0, m1, r1;
            ln_stats(xr, m0, r0, m1, r1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                xr[j][0] = (xr[j][0] - m0) * r0; xr[j][1] = (xr[j][1] - m0) * r0;
                xr[j][2] = (xr[j][2] - m1) * r1; xr[j][3] = (xr[j][3] - m1) * r1;
            }
            uint32_t af[2][4];
            acc_to_afrag(xr, af);
            const float* bf = sV + kVecBlock0 + P.R * kVecPerBlock;
            float o[1][4];
            o[0][0] = bf[2 * t]; o[0][1] = bf[2 * t + 1]; o[0][2] = o[0][0]; o[0][3] = o[0][1];
            mlp_mma<1>(o, af, sW + kFrag32 + P.R * kBlockFrag, 0, lane);
            // lane t=0 holds channels 0,1; t=1 holds channel 2 (cols 3..7 are padding)
            TOut* ob = reinterpret_cast<TOut*>(P.out) + (size_t)b * 3 * plane + pix0;
            if (t == 0) {
                stf(ob + g, o[0][0]); stf(ob + g + 8, o[0][2]);
                stf(ob + plane + g, o[0][1]); stf(ob + plane + g + 8, o[0][3]);
            } else if (t == 1) {
                stf(ob + 2 * plane + g, o[0][0]); stf(ob + 2 * plane + g + 8, o[0][2]);
            }
        }
    }
}

}  // namespace deco

extern "C" int deco_decoder_blob_bytes(int num_res_blocks) {
    return (deco::dec_frag_words(num_res_blocks) + deco::dec_vec_floats(num_res_blocks)) * 4;
}

extern "C" int deco_pixel_decoder(const float* x, const void* ycond_bf16, const void* blob, const float* postab,
                                  void* out, int out_is_bf16, int B, int H, int W, int patch, int hidden_x,
                                  int num_res_blocks, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(x && ycond_bf16 && blob && postab && out, "pixel_decoder: null pointer");
    if (patch != 16 || hidden_x != kHx) {
        deco_set_error("pixel_decoder: built for patch_size 16 and hidden_size_x 32 (got %d, %d)", patch, hidden_x);
        return DECO_ERR_UNSUPPORTED;
    }
    DECO_CHECK_ARG(B > 0 && H % 16 == 0 && W % 16 == 0 && num_res_blocks >= 1 && num_res_blocks <= 6,
                   "pixel_decoder: bad shape B=%d H=%d W=%d R=%d", B, H, W, num_res_blocks);
    DecParams P;
    P.x = x; P.ycond = (const __nv_bfloat16*)ycond_bf16; P.blob = (const uint32_t*)blob; P.postab = postab;
    P.out = out; P.R = num_res_blocks; P.H = H; P.W = W; P.Hp = H / 16; P.Wp = W / 16;
    P.M = (long long)B * P.Hp * P.Wp;
    const int smem_bytes = deco_decoder_blob_bytes(num_res_blocks);
    const long long warps_needed = P.M * 16;
    long long grid = (warps_needed + 7) / 8;
    if (grid > 2LL * kNumSMs) grid = 2LL * kNumSMs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = deco_pdl_enabled() ? 1 : 0;
    cudaError_t e;
    if (out_is_bf16) {
        e = cudaFuncSetAttribute(pixel_decoder_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { deco_set_error("pixel_decoder attr: %s", cudaGetErrorString(e)); return (int)e; }
        e = cudaLaunchKernelEx(&cfg, pixel_decoder_kernel<__nv_bfloat16>, P);
    } else {
        e = cudaFuncSetAttribute(pixel_decoder_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) { deco_set_error("pixel_decoder attr: %s", cudaGetErrorString(e)); return (int)e; }
        e = cudaLaunchKernelEx(&cfg, pixel_decoder_kernel<float>, P);
    }
    if (e != cudaSuccess) { deco_set_error("pixel_decoder launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    DECO_CHECK_LAUNCH("pixel_decoder_kernel");
    return DECO_OK;
}
