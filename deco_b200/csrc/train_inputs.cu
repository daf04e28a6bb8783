// Inputs of one flow-matching training step, fused: the 90/10 timestep mixture + time shift + scheduler coefficients,
// the (x_t, v_t) interpolation pair, and the classifier-free label dropout.
//
// Replaces (reference, /root/reference):
//   src/diffusion/flow_matching/training_repa_DeCo.py:222-229   t = where(u_sel <= 0.9, sigmoid(randn), u_uni), time_shift_fn
//   src/diffusion/flow_matching/training_repa_DeCo.py:231-237   alpha / sigma / dalpha / dsigma, x_t = alpha x + sigma eps,
//                                                               v_t = dalpha x + dsigma eps
//   src/diffusion/flow_matching/scheduling.py:6-14              LinearScheduler (alpha = t, sigma = 1 - t, dalpha = 1, dsigma = -1)
//   src/diffusion/base/training.py:14-20                        label dropout: condition*(1-mask) + uncondition*mask
// The random draws themselves stay torch's (randn / rand / randn_like on the CUDA generator, in the reference's order), so
// a seeded run consumes the same Philox stream as the reference; everything that FOLLOWS the draws -- ~15 eager
// element-wise kernels and their temporaries in the reference -- is the three kernels below.
// Bound: HBM.  flow_pair: 8 B read + 8 B written per element (x, eps -> x_t, v_t), the minimum for two outputs.
#include "common.cuh"

namespace deco {

// one thread per image: t and the four scheduler coefficients (alpha, sigma, dalpha, dsigma)
__global__ void train_timesteps_kernel(const float* __restrict__ nt, const float* __restrict__ u_uniform,
                                       const float* __restrict__ u_select, float timeshift, int linear,
                                       float* __restrict__ t_out, float4* __restrict__ coef_out, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float t_lognorm = 1.0f / (1.0f + expf(-nt[i]));
    const float base = u_select[i] <= 0.9f ? t_lognorm : u_uniform[i];
    const float t = base / (base + (1.0f - base) * timeshift);
    t_out[i] = t;
    if (linear) coef_out[i] = make_float4(t, 1.0f - t, 1.0f, -1.0f);
}

// x_t = alpha x + eps sigma ; v_t = dalpha x + dsigma eps, coefficients per image; 128-bit accesses, per_image % 4 == 0
__global__ void __launch_bounds__(256)
flow_pair_kernel(const float4* __restrict__ x, const float4* __restrict__ eps, const float4* __restrict__ coef,
                 float4* __restrict__ x_t, float4* __restrict__ v_t, long long n4, long long per_image4)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 c = __ldg(coef + i / per_image4);
        const uint4 xr = ld_stream16(x + i), er = ld_stream16(eps + i);
        const float4 a = make_float4(__uint_as_float(xr.x), __uint_as_float(xr.y), __uint_as_float(xr.z), __uint_as_float(xr.w));
        const float4 e = make_float4(__uint_as_float(er.x), __uint_as_float(er.y), __uint_as_float(er.z), __uint_as_float(er.w));
        // the reference's expression order (two products, one sum): no fused multiply-add contraction across the sum
        float4 xt, vt;
        xt.x = __fadd_rn(__fmul_rn(c.x, a.x), __fmul_rn(e.x, c.y)); vt.x = __fadd_rn(__fmul_rn(c.z, a.x), __fmul_rn(c.w, e.x));
        xt.y = __fadd_rn(__fmul_rn(c.x, a.y), __fmul_rn(e.y, c.y)); vt.y = __fadd_rn(__fmul_rn(c.z, a.y), __fmul_rn(c.w, e.y));
        xt.z = __fadd_rn(__fmul_rn(c.x, a.z), __fmul_rn(e.z, c.y)); vt.z = __fadd_rn(__fmul_rn(c.z, a.z), __fmul_rn(c.w, e.z));
        xt.w = __fadd_rn(__fmul_rn(c.x, a.w), __fmul_rn(e.w, c.y)); vt.w = __fadd_rn(__fmul_rn(c.z, a.w), __fmul_rn(c.w, e.w));
        st_stream16(x_t + i, make_uint4(__float_as_uint(xt.x), __float_as_uint(xt.y), __float_as_uint(xt.z), __float_as_uint(xt.w)));
        st_stream16(v_t + i, make_uint4(__float_as_uint(vt.x), __float_as_uint(vt.y), __float_as_uint(vt.z), __float_as_uint(vt.w)));
    }
}

__global__ void label_dropout_kernel(const long long* __restrict__ cond, const long long* __restrict__ uncond,
                                     const float* __restrict__ u, float p, long long* __restrict__ out, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) out[i] = u[i] < p ? uncond[i] : cond[i];
}

}  // namespace deco

using namespace deco;

extern "C" int deco_train_timesteps(const float* nt, const float* u_uniform, const float* u_select, float timeshift,
                                    int linear_scheduler, float* t_out, float* coef_out, int B, void* stream)
{
    DECO_CHECK_ARG(nt && u_uniform && u_select && t_out && B > 0, "train_timesteps: null pointer / empty batch");
    DECO_CHECK_ARG(!linear_scheduler || (coef_out && ((uintptr_t)coef_out & 15) == 0), "train_timesteps: bad coefficient buffer");
    train_timesteps_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nt, u_uniform, u_select, timeshift,
                                                                              linear_scheduler, t_out, (float4*)coef_out, B);
    DECO_CHECK_LAUNCH("train_timesteps_kernel");
    return DECO_OK;
}

extern "C" int deco_flow_pair(const float* x, const float* eps, const float* coef, float* x_t, float* v_t,
                              int B, long long per_image, void* stream)
{
    DECO_CHECK_ARG(x && eps && coef && x_t && v_t && B > 0 && per_image > 0, "flow_pair: null pointer / empty batch");
    DECO_CHECK_ARG(per_image % 4 == 0, "flow_pair: elements per image must be a multiple of 4 (got %lld)", per_image);
    DECO_CHECK_ARG((((uintptr_t)x | (uintptr_t)eps | (uintptr_t)coef | (uintptr_t)x_t | (uintptr_t)v_t) & 15) == 0,
                   "flow_pair: pointers must be 16-byte aligned");
    const long long n4 = (long long)B * per_image / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)device_sm_count() * 16;       // grid-stride: a few resident waves
    if (blocks > cap) blocks = cap;
    flow_pair_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)eps, (const float4*)coef,
                                                                   (float4*)x_t, (float4*)v_t, n4, per_image / 4);
    DECO_CHECK_LAUNCH("flow_pair_kernel");
    return DECO_OK;
}

extern "C" int deco_label_dropout(const long long* cond, const long long* uncond, const float* u, float p,
                                  long long* out, int B, void* stream)
{
    DECO_CHECK_ARG(cond && uncond && u && out && B > 0, "label_dropout: null pointer / empty batch");
    label_dropout_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cond, uncond, u, p, out, B);
    DECO_CHECK_LAUNCH("label_dropout_kernel");
    return DECO_OK;
}
