// CFG-batched flow sampler update: guidance combine + Euler / Heun / Adams-multistep state update, optional
// prediction stash and optional fp2uint8 of the new state, one pass.
//
// Replaces (reference, /root/reference):
//   src/diffusion/base/guidance.py:3-6          simple_guidance_fn  (rows [uncond || cond])
//   src/diffusion/flow_matching/sampling.py:14-15, :89-104  ode_step_fn + Euler loop body
//   src/diffusion/flow_matching/sampling.py:283-291        Heun corrector average
//   src/diffusion/flow_matching/adam_sampling.py:109-117   multistep combination
//   src/models/autoencoder/base.py:32-34                    fp2uint8
//
//   pred  = u + g * (c - u)
//   v     = c0 * pred + c1 * p1 + c2 * p2 + c3 * p3          (p_j = earlier predictions, fp32)
//   x_out = x + dt * v
// Euler: c0 = 1.  Heun corrector: c0 = c1 = 1/2 with p1 = predictor velocity.  Adams order k: c_j from the host.
//
// Extended form (EXT = true; deco_cfg_step_ex), the other step functions / samplers that share this pass:
//   src/diffusion/flow_matching/sampling.py:170            EulerSamplerJiT: the net predicts x, u' = (u - x) / den,
//                                                          c' = (c - x) / den with den = clamp_min(1 - t, 0.05)
//   src/diffusion/flow_matching/sampling.py:17-24, :98     sde_mean / sde / sde_preserve step functions:
//       s     = (kd * v - x) / sden          kd = 1 / dalpha_over_alpha(t), sden = sigma^2 - kd * dsigma_mul_sigma(t)
//       x_out = x + dt * v + a_s * s + a_n * noise        (a_s, a_n) = (w dt, 0) | (w dt, sqrt(2 w dt)) | (w dt / 2, sqrt(w dt))
//   noise = the caller's torch.randn_like(x) (same CUDA Philox stream as the reference's call at :21,:24)
// Bound: HBM.  Euler algorithmic bytes / element: 4 (x) + 2*sizeof(net out) + 4 (x_out) = 12 B with bf16 net output.
#include "common.cuh"

namespace deco {

struct StepArgs {
    const float* x;
    const void* net_out;     // [2B, ...] rows [uncond || cond]
    const float* p[3];       // optional earlier predictions
    float* x_out;
    float* pred_out;         // optional: store pred (fp32)
    float* v_out;            // optional: store combined v (fp32)
    uint8_t* u8_out;         // optional: fp2uint8(x_out)
    float g, dt, c0, c[3];
    const float* dev;        // optional: {g, dt, c0, c1, c2, c3} in DEVICE memory (CUDA-graph replays: the step's scalars
                             // change between replays without changing the kernel arguments)
    long long n;             // elements per CFG half (B*C*H*W)
    // extended form only
    float xpred_den;         // > 0: the net predicts x (EulerSamplerJiT); dev[7] when dev != NULL
    float kd, sden, a_s, a_n;
    const float* noise;      // required when a_n != 0
};

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ float4 load(const void* p, long long i) {
        return __ldg(reinterpret_cast<const float4*>(p) + i);
    }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load(const void* p, long long i) {
        uint2 r = __ldg(reinterpret_cast<const uint2*>(p) + i);
        float2 a = unpack_bf2(r.x), b = unpack_bf2(r.y);
        return make_float4(a.x, a.y, b.x, b.y);
    }
};

__device__ __forceinline__ uint8_t to_u8(float x) {
    // clamp((x + 1) * 127.5 + 0.5, 0, 255).to(uint8): truncation toward zero
    float v = fminf(fmaxf((x + 1.0f) * 127.5f + 0.5f, 0.0f), 255.0f);
    return (uint8_t)v;
}

// INPLACE = false: inputs are read through the non-coherent path (ld.global.nc), which lets the compiler hoist the loads of
// later grid-stride iterations above earlier stores (94 % of the HBM peak vs 76 % with plain loads); INPLACE = true (graph
// replays update the state in place: x_out == x, pred_out == p1) must use plain loads.
template <typename TNet, bool INPLACE, bool EXT = false>
__global__ void __launch_bounds__(256) cfg_step_kernel(StepArgs a) {
    if (a.dev) {
        a.g = __ldg(a.dev); a.dt = __ldg(a.dev + 1); a.c0 = __ldg(a.dev + 2);
        a.c[0] = __ldg(a.dev + 3); a.c[1] = __ldg(a.dev + 4); a.c[2] = __ldg(a.dev + 5);
        if (EXT) a.xpred_den = __ldg(a.dev + 7);
    }
    const long long n4 = a.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = INPLACE ? reinterpret_cast<const float4*>(a.x)[i] : __ldg(reinterpret_cast<const float4*>(a.x) + i);
        float4 u = Vec4<TNet>::load(a.net_out, i);
        float4 c = Vec4<TNet>::load(a.net_out, i + n4);
        if (EXT && a.xpred_den > 0.f) {       // x-prediction -> velocity, each CFG half against its own copy of x
            const float den = a.xpred_den;
            u.x = (u.x - x.x) / den; u.y = (u.y - x.y) / den; u.z = (u.z - x.z) / den; u.w = (u.w - x.w) / den;
            c.x = (c.x - x.x) / den; c.y = (c.y - x.y) / den; c.z = (c.z - x.z) / den; c.w = (c.w - x.w) / den;
        }
        float4 pr;
        pr.x = u.x + a.g * (c.x - u.x);
        pr.y = u.y + a.g * (c.y - u.y);
        pr.z = u.z + a.g * (c.z - u.z);
        pr.w = u.w + a.g * (c.w - u.w);
        float4 v = make_float4(a.c0 * pr.x, a.c0 * pr.y, a.c0 * pr.z, a.c0 * pr.w);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (a.p[j]) {
                const float4 q = INPLACE ? reinterpret_cast<const float4*>(a.p[j])[i] : __ldg(reinterpret_cast<const float4*>(a.p[j]) + i);
                v.x = fmaf(a.c[j], q.x, v.x); v.y = fmaf(a.c[j], q.y, v.y);
                v.z = fmaf(a.c[j], q.z, v.z); v.w = fmaf(a.c[j], q.w, v.w);
            }
        }
        float4 xo = make_float4(fmaf(a.dt, v.x, x.x), fmaf(a.dt, v.y, x.y), fmaf(a.dt, v.z, x.z), fmaf(a.dt, v.w, x.w));
        if (EXT && a.a_s != 0.f) {            // score term of the SDE step functions
            xo.x = fmaf(a.a_s, (a.kd * v.x - x.x) / a.sden, xo.x); xo.y = fmaf(a.a_s, (a.kd * v.y - x.y) / a.sden, xo.y);
            xo.z = fmaf(a.a_s, (a.kd * v.z - x.z) / a.sden, xo.z); xo.w = fmaf(a.a_s, (a.kd * v.w - x.w) / a.sden, xo.w);
        }
        if (EXT && a.a_n != 0.f) {
            const float4 z = __ldg(reinterpret_cast<const float4*>(a.noise) + i);
            xo.x = fmaf(a.a_n, z.x, xo.x); xo.y = fmaf(a.a_n, z.y, xo.y); xo.z = fmaf(a.a_n, z.z, xo.z); xo.w = fmaf(a.a_n, z.w, xo.w);
        }
        if (a.x_out) reinterpret_cast<float4*>(a.x_out)[i] = xo;
        if (a.pred_out) reinterpret_cast<float4*>(a.pred_out)[i] = pr;
        if (a.v_out) reinterpret_cast<float4*>(a.v_out)[i] = v;
        if (a.u8_out) {
            uchar4 q = make_uchar4(to_u8(xo.x), to_u8(xo.y), to_u8(xo.z), to_u8(xo.w));
            reinterpret_cast<uchar4*>(a.u8_out)[i] = q;
        }
    }
}

// One replay of a graphed sampling step starts here: row (*counter % rows) of the host-precomputed schedule table
// {g, dt, c0, c1, c2, c3, t, -} becomes the step's scalars, t is broadcast into the denoiser's timestep vector, and the
// counter moves on -- so consecutive replays walk the schedule with no host work in between.
__global__ void sampler_advance_kernel(const float* __restrict__ table, int rows, int* counter, float* cur, float* t_out, int nt) {
    const int idx = *counter % rows;
    const float* row = table + (size_t)idx * 8;
    if (threadIdx.x < 8) cur[threadIdx.x] = row[threadIdx.x];
    const float t = row[6];
    for (int i = threadIdx.x; i < nt; i += blockDim.x) t_out[i] = t;
    __syncthreads();
    if (threadIdx.x == 0) *counter = idx + 1;
}

__global__ void __launch_bounds__(256) fp2uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ o, long long n4) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        reinterpret_cast<uchar4*>(o)[i] = make_uchar4(to_u8(v.x), to_u8(v.y), to_u8(v.z), to_u8(v.w));
    }
}

// Heun predictor / corrector with the SDE step functions (src/diffusion/flow_matching/sampling.py:17-24, :266-293):
//   s      = s_in ? s_in : (kd v - x) / sden                       score at (x, t_cur), or the one kept from the last corrector
//   predictor:  x_out = x + dt v + a_s s + a_n z
//   corrector:  v_hat = u + g (c - u);  s_hat = (kdh v_hat - x_hat) / sdenh        (network evaluated at (x_hat, t_next))
//               x_out = x + dt (v + v_hat) / 2 + a_s (s + s_hat) / 2 + a_n z ;  v_hat, s_hat stored for the next predictor
// Bound: HBM (corrector: x, x_hat, v fp32 + 2 x bf16 net rows + z read, x_out / v_hat / s_hat written = 36 B per element).
struct HeunSdeArgs {
    const float* x; const float* v; const float* s_in; const void* net_out; const float* x_hat; const float* noise;
    float g, dt, kd, sden, kdh, sdenh, a_s, a_n;
    int corrector;
    float* x_out; float* v_hat_out; float* s_hat_out; float* v_avg_out; uint8_t* u8_out;
    long long n;
};

template <typename TNet>
__global__ void __launch_bounds__(256) heun_sde_step_kernel(HeunSdeArgs a) {
    const long long n4 = a.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x4 = reinterpret_cast<const float4*>(a.x)[i];
        const float4 v4 = reinterpret_cast<const float4*>(a.v)[i];
        const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w};
        float ss[4];
        if (a.s_in) {
            const float4 s4 = reinterpret_cast<const float4*>(a.s_in)[i];
            ss[0] = s4.x; ss[1] = s4.y; ss[2] = s4.z; ss[3] = s4.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) ss[e] = (a.kd * vs[e] - xs[e]) / a.sden;
        }
        float zs[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.a_n != 0.f) {
            const float4 z4 = __ldg(reinterpret_cast<const float4*>(a.noise) + i);
            zs[0] = z4.x; zs[1] = z4.y; zs[2] = z4.z; zs[3] = z4.w;
        }
        float ve[4], se[4], vh[4], sh[4];
        if (a.corrector) {
            const float4 u4 = Vec4<TNet>::load(a.net_out, i), c4 = Vec4<TNet>::load(a.net_out, i + n4);
            const float4 h4 = reinterpret_cast<const float4*>(a.x_hat)[i];
            const float us[4] = {u4.x, u4.y, u4.z, u4.w}, cs[4] = {c4.x, c4.y, c4.z, c4.w}, hs[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                vh[e] = us[e] + a.g * (cs[e] - us[e]);
                sh[e] = (a.kdh * vh[e] - hs[e]) / a.sdenh;
                ve[e] = (vs[e] + vh[e]) * 0.5f;
                se[e] = (ss[e] + sh[e]) * 0.5f;
            }
            if (a.v_hat_out) reinterpret_cast<float4*>(a.v_hat_out)[i] = make_float4(vh[0], vh[1], vh[2], vh[3]);
            if (a.s_hat_out) reinterpret_cast<float4*>(a.s_hat_out)[i] = make_float4(sh[0], sh[1], sh[2], sh[3]);
            if (a.v_avg_out) reinterpret_cast<float4*>(a.v_avg_out)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) { ve[e] = vs[e]; se[e] = ss[e]; }
        }
        float xo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) xo[e] = fmaf(a.a_n, zs[e], fmaf(a.a_s, se[e], fmaf(a.dt, ve[e], xs[e])));
        reinterpret_cast<float4*>(a.x_out)[i] = make_float4(xo[0], xo[1], xo[2], xo[3]);
        if (a.u8_out) reinterpret_cast<uchar4*>(a.u8_out)[i] = make_uchar4(to_u8(xo[0]), to_u8(xo[1]), to_u8(xo[2]), to_u8(xo[3]));
    }
}

}  // namespace deco

// Heun step with an SDE step function (see heun_sde_step_kernel).  corrector == 0: predictor / last step from (x, v, s);
// corrector != 0: net_out [2n] rows [uncond || cond] evaluated at (x_hat, t_next).  s_in NULL = score from (x, v, kd, sden).
// x_out must not alias x when the caller still needs x (the corrector of the same step does).
extern "C" int deco_heun_sde_step(const float* x, const float* v, const float* s_in, const void* net_out, int net_is_bf16,
                                  const float* x_hat, const float* noise, float g, float dt, float kd, float sden,
                                  float kdh, float sdenh, float a_s, float a_n, int corrector,
                                  float* x_out, float* v_hat_out, float* s_hat_out, float* v_avg_out, uint8_t* u8_out,
                                  long long n, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(x && v && x_out, "heun_sde_step: null input");
    DECO_CHECK_ARG(n > 0 && (n % 4) == 0, "heun_sde_step: element count %lld must be a positive multiple of 4", n);
    DECO_CHECK_ARG(!corrector || (net_out && x_hat && sdenh != 0.f), "heun_sde_step: the corrector needs net_out, x_hat, sdenh");
    DECO_CHECK_ARG(a_n == 0.f || noise, "heun_sde_step: a_n != 0 needs a noise tensor");
    DECO_CHECK_ARG(s_in || sden != 0.f, "heun_sde_step: zero score denominator");
    HeunSdeArgs a;
    a.x = x; a.v = v; a.s_in = s_in; a.net_out = net_out; a.x_hat = x_hat; a.noise = noise;
    a.g = g; a.dt = dt; a.kd = kd; a.sden = sden; a.kdh = kdh; a.sdenh = sdenh; a.a_s = a_s; a.a_n = a_n;
    a.corrector = corrector ? 1 : 0;
    a.x_out = x_out; a.v_hat_out = v_hat_out; a.s_hat_out = s_hat_out; a.v_avg_out = v_avg_out; a.u8_out = u8_out; a.n = n;
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (net_is_bf16) heun_sde_step_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    else heun_sde_step_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    DECO_CHECK_LAUNCH("heun_sde_step_kernel");
    return DECO_OK;
}

extern "C" int deco_cfg_step(const float* x, const void* net_out, int net_is_bf16,
                             const float* p1, const float* p2, const float* p3,
                             float g, float dt, float c0, float c1, float c2, float c3,
                             float* x_out, float* pred_out, float* v_out, uint8_t* u8_out,
                             long long n, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(x && net_out, "cfg_step: null input");
    DECO_CHECK_ARG(n > 0 && (n % 4) == 0, "cfg_step: element count %lld must be a positive multiple of 4", n);
    StepArgs a;
    a.x = x; a.net_out = net_out; a.p[0] = p1; a.p[1] = p2; a.p[2] = p3;
    a.x_out = x_out; a.pred_out = pred_out; a.v_out = v_out; a.u8_out = u8_out;
    a.g = g; a.dt = dt; a.c0 = c0; a.c[0] = c1; a.c[1] = c2; a.c[2] = c3; a.n = n; a.dev = nullptr;
    a.xpred_den = 0.f; a.kd = 0.f; a.sden = 1.f; a.a_s = 0.f; a.a_n = 0.f; a.noise = nullptr;
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;   // grid-stride: 16 CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    const bool inplace = x_out == x || (pred_out && (pred_out == p1 || pred_out == p2 || pred_out == p3)) ||
                         (v_out && (v_out == p1 || v_out == p2 || v_out == p3));
    cudaStream_t st = (cudaStream_t)stream;
    if (net_is_bf16) {
        if (inplace) cfg_step_kernel<__nv_bfloat16, true><<<(unsigned)blocks, 256, 0, st>>>(a);
        else cfg_step_kernel<__nv_bfloat16, false><<<(unsigned)blocks, 256, 0, st>>>(a);
    } else {
        if (inplace) cfg_step_kernel<float, true><<<(unsigned)blocks, 256, 0, st>>>(a);
        else cfg_step_kernel<float, false><<<(unsigned)blocks, 256, 0, st>>>(a);
    }
    DECO_CHECK_LAUNCH("cfg_step_kernel");
    return DECO_OK;
}

extern "C" int deco_fp2uint8(const float* x, uint8_t* out, long long n, void* stream) {
    using namespace deco;
    DECO_CHECK_ARG(x && out && n > 0 && (n % 4) == 0, "fp2uint8: bad arguments");
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    fp2uint8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, out, n4);
    DECO_CHECK_LAUNCH("fp2uint8_kernel");
    return DECO_OK;
}

// deco_cfg_step with the step scalars {g, dt, c0, c1, c2, c3} read from device memory (dev_params), for CUDA-graph replays.
// x_out may alias x and pred_out may alias p1 (element-wise, read before write).
extern "C" int deco_cfg_step_dev(const float* x, const void* net_out, int net_is_bf16,
                                 const float* p1, const float* p2, const float* p3, const float* dev_params,
                                 float* x_out, float* pred_out, float* v_out, uint8_t* u8_out, long long n, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(x && net_out && dev_params, "cfg_step_dev: null input");
    DECO_CHECK_ARG(n > 0 && (n % 4) == 0, "cfg_step_dev: element count %lld must be a positive multiple of 4", n);
    StepArgs a;
    a.x = x; a.net_out = net_out; a.p[0] = p1; a.p[1] = p2; a.p[2] = p3;
    a.x_out = x_out; a.pred_out = pred_out; a.v_out = v_out; a.u8_out = u8_out;
    a.g = 1.f; a.dt = 0.f; a.c0 = 1.f; a.c[0] = 0.f; a.c[1] = 0.f; a.c[2] = 0.f; a.n = n; a.dev = dev_params;
    a.xpred_den = 0.f; a.kd = 0.f; a.sden = 1.f; a.a_s = 0.f; a.a_n = 0.f; a.noise = nullptr;
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    const bool inplace = x_out == x || (pred_out && pred_out == p1);
    cudaStream_t st = (cudaStream_t)stream;
    if (net_is_bf16) {
        if (inplace) cfg_step_kernel<__nv_bfloat16, true><<<(unsigned)blocks, 256, 0, st>>>(a);
        else cfg_step_kernel<__nv_bfloat16, false><<<(unsigned)blocks, 256, 0, st>>>(a);
    } else {
        if (inplace) cfg_step_kernel<float, true><<<(unsigned)blocks, 256, 0, st>>>(a);
        else cfg_step_kernel<float, false><<<(unsigned)blocks, 256, 0, st>>>(a);
    }
    DECO_CHECK_LAUNCH("cfg_step_kernel");
    return DECO_OK;
}

// Extended update (see the top of the file): x-prediction nets and the SDE step functions.  dev_params (optional, device)
// = {g, dt, c0, c1, c2, c3, t, xpred_den} as written by deco_sampler_advance; NULL = the scalar arguments are used.
// noise fp32 [n] is read only when a_n != 0.  x_out may alias x, pred_out may alias p1.
extern "C" int deco_cfg_step_ex(const float* x, const void* net_out, int net_is_bf16,
                                const float* p1, const float* p2, const float* p3, const float* dev_params,
                                float g, float dt, float c0, float c1, float c2, float c3,
                                float xpred_den, float kd, float sden, float a_s, float a_n, const float* noise,
                                float* x_out, float* pred_out, float* v_out, uint8_t* u8_out, long long n, void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(x && net_out, "cfg_step_ex: null input");
    DECO_CHECK_ARG(n > 0 && (n % 4) == 0, "cfg_step_ex: element count %lld must be a positive multiple of 4", n);
    DECO_CHECK_ARG(a_n == 0.f || noise, "cfg_step_ex: a_n != 0 needs a noise tensor");
    DECO_CHECK_ARG(a_s == 0.f || sden != 0.f, "cfg_step_ex: zero score denominator");
    DECO_CHECK_ARG(xpred_den >= 0.f, "cfg_step_ex: negative x-prediction denominator");
    StepArgs a;
    a.x = x; a.net_out = net_out; a.p[0] = p1; a.p[1] = p2; a.p[2] = p3;
    a.x_out = x_out; a.pred_out = pred_out; a.v_out = v_out; a.u8_out = u8_out;
    a.g = g; a.dt = dt; a.c0 = c0; a.c[0] = c1; a.c[1] = c2; a.c[2] = c3; a.n = n; a.dev = dev_params;
    a.xpred_den = xpred_den; a.kd = kd; a.sden = sden; a.a_s = a_s; a.a_n = a_n; a.noise = noise;
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    const bool inplace = x_out == x || (pred_out && (pred_out == p1 || pred_out == p2 || pred_out == p3)) ||
                         (v_out && (v_out == p1 || v_out == p2 || v_out == p3));
    cudaStream_t st = (cudaStream_t)stream;
    if (net_is_bf16) {
        if (inplace) cfg_step_kernel<__nv_bfloat16, true, true><<<(unsigned)blocks, 256, 0, st>>>(a);
        else cfg_step_kernel<__nv_bfloat16, false, true><<<(unsigned)blocks, 256, 0, st>>>(a);
    } else {
        if (inplace) cfg_step_kernel<float, true, true><<<(unsigned)blocks, 256, 0, st>>>(a);
        else cfg_step_kernel<float, false, true><<<(unsigned)blocks, 256, 0, st>>>(a);
    }
    DECO_CHECK_LAUNCH("cfg_step_kernel<ext>");
    return DECO_OK;
}

// Head of a graphed sampling step: cur[0..8) = table[*counter % rows], t_out[0..nt) = that row's t, ++*counter.
extern "C" int deco_sampler_advance(const float* table, int rows, int* counter, float* cur_params, float* t_out, int nt,
                                    void* stream)
{
    using namespace deco;
    DECO_CHECK_ARG(table && counter && cur_params && t_out && rows > 0 && nt > 0, "sampler_advance: bad arguments");
    sampler_advance_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(table, rows, counter, cur_params, t_out, nt);
    DECO_CHECK_LAUNCH("sampler_advance_kernel");
    return DECO_OK;
}
