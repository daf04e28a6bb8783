// Shared device helpers for the deco_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define DECO_OK 0
#define DECO_ERR_ARG (-1)
#define DECO_ERR_UNSUPPORTED (-2)
#define DECO_ERR_DRIVER (-3)

void deco_set_error(const char* fmt, ...);

#define DECO_CHECK_ARG(cond, ...)                       \
    do {                                                \
        if (!(cond)) {                                  \
            deco_set_error(__VA_ARGS__);                \
            return DECO_ERR_ARG;                        \
        }                                               \
    } while (0)

#define DECO_CHECK_LAUNCH(name)                                               \
    do {                                                                      \
        cudaError_t e_ = cudaGetLastError();                                  \
        if (e_ != cudaSuccess) {                                              \
            deco_set_error("%s launch failed: %s", name, cudaGetErrorString(e_)); \
            return (int)e_;                                                   \
        }                                                                     \
    } while (0)

// process-wide switch for programmatic dependent launch (DECO_B200_PDL=0 disables it); defined in api.cu
bool deco_pdl_enabled();
// SMs the persistent GEMMs leave to concurrent kernels (deco_gemm_reserve_sms); defined in api.cu
int deco_reserved_sms();

namespace deco {

constexpr int kNumSMs = 148;

// One-time per-DEVICE host setup (cudaFuncSetAttribute, SM count): a process may drive several GPUs, and function
// attributes / device properties belong to the current device's context, not to the process.
inline int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev & 63;
}
inline bool device_setup_done(const unsigned long long& mask) { return (mask >> current_device_slot()) & 1ull; }
inline void mark_device_setup(unsigned long long& mask) { mask |= 1ull << current_device_slot(); }
inline int device_sm_count() {
    static int cached[64] = {0};
    const int slot = current_device_slot();
    if (!cached[slot]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, slot);
        cached[slot] = n > 0 ? n : kNumSMs;
    }
    return cached[slot];
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float round_bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf2(uint32_t v) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
    return __bfloat1622float2(t);
}

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return bf2f(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = f2bf(v); }

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// silu(x) = x sigmoid(x) = h + h tanh(h), h = x / 2: ONE MUFU op (tanh.approx, relative error ~2^-11) instead of two
// (ex2 + rcp); for outputs that are rounded to bf16 anyway (GEMM epilogues, tensor-core operands)
__device__ __forceinline__ float silu_fast(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 128-bit streaming loads/stores (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while
// its predecessor in the stream drains -- its prologue (barrier init, TMEM allocation, descriptor prefetch, constant
// weights) overlaps the predecessor's tail -- and must call pdl_wait() before touching anything the predecessor wrote or
// still reads; pdl_launch_dependents() lets the NEXT kernel be scheduled as soon as SMs free up.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// cp.async 16B with zero-fill when !pred
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    uint32_t s = smem_u32(smem);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// legacy warp-level tensor-core MMA (used by the register-resident per-pixel MLP and by attention v1)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
    uint32_t s = smem_u32(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], const void* smem) {
    uint32_t s = smem_u32(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(s));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
    uint32_t s = smem_u32(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem) {
    uint32_t s = smem_u32(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(s));
}

}  // namespace deco
