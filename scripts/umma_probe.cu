// Stand-alone probe for the shared-memory / tensor-memory operand layouts the tcgen05 attention kernel relies on
// (csrc/attention_tc.cu).  One CTA, real tcgen05.mma instructions, results checked against a host reference:
//   A  S = Q . K^T        Q, K bf16 [128 x 80] K-major, 32-byte swizzle, chunk-major storage (sw32_offset)
//   B  O = P . V          P bf16 [128 x 128] in TENSOR MEMORY (tcgen05.st, two K per 32-bit column),
//                         V bf16 [128 keys x 80] MN-major with the same storage function
//   C  O = P . V          P in shared memory (K-major SW32), V as in B
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ideco_b200/csrc scripts/umma_probe.cu -o build/umma_probe
#include "tcgen05.cuh"
#include <cstdlib>
#include <cmath>
#include <vector>

void deco_set_error(const char*, ...) {}

using namespace deco;

constexpr int R = 128, DP = 80, KV = 128;
constexpr int kTileBytes = (DP / 16) * R * 32;      // 20480
constexpr int kPBytes = (KV / 16) * R * 32;         // 32768

struct ProbeOut { float* S; float* O_ts; float* O_ss; };

__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* Q, const __nv_bfloat16* K,
                                                       const __nv_bfloat16* P, const __nv_bfloat16* V, ProbeOut out,
                                                       uint32_t v_lbo, uint32_t v_sbo)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sQ = base, sK = sQ + kTileBytes, sV = sK + kTileBytes, sP = sV + kTileBytes;
    const uint32_t bar = sP + kPBytes, slot = bar + 16;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < R * DP; i += 128) {
        const int r = i / DP, c = i % DP;
        *reinterpret_cast<__nv_bfloat16*>(gen + (sQ - base) + sw32_offset(r, c, R)) = Q[i];
        *reinterpret_cast<__nv_bfloat16*>(gen + (sK - base) + sw32_offset(r, c, R)) = K[i];
        *reinterpret_cast<__nv_bfloat16*>(gen + (sV - base) + sw32_offset(r, c, KV)) = V[i];
    }
    for (int i = tid; i < R * KV; i += 128) {
        const int r = i / KV, c = i % KV;
        *reinterpret_cast<__nv_bfloat16*>(gen + (sP - base) + sw32_offset(r, c, R)) = P[i];
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

    // P into tensor memory, columns [256, 320): thread = row, column c = (P[r][2c], P[r][2c+1])
    {
        const uint32_t* prow = reinterpret_cast<const uint32_t*>(P + (size_t)tid * KV);
#pragma unroll 1
        for (int c = 0; c < KV / 2; c += 16) {
            uint32_t v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = prow[c + i];
            tmem_st16(tmem + lane_base + 256u + (uint32_t)c, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (tid == 0) {
        // A: S -> columns [0, 128)
        const uint32_t id_s = make_idesc_major(128, 128, 0, 0);
        for (int kc = 0; kc < DP / 16; ++kc)
            umma_bf16(tmem + 0u, make_umma_desc(sQ + kc * R * 32, 16, 256, 6), make_umma_desc(sK + kc * R * 32, 16, 256, 6),
                      id_s, kc ? 1u : 0u);
        // B: O_ts -> columns [128, 208)
        const uint32_t id_o = make_idesc_major(128, DP, 0, 1);
        for (int ks = 0; ks < KV / 16; ++ks)
            umma_bf16_ts(tmem + 128u, tmem + 256u + (uint32_t)(ks * 8), make_umma_desc(sV + ks * 512, v_lbo, v_sbo, 6),
                         id_o, ks ? 1u : 0u);
        // C: O_ss -> columns [384, 464)
        for (int ks = 0; ks < KV / 16; ++ks)
            umma_bf16(tmem + 384u, make_umma_desc(sP + ks * R * 32, 16, 256, 6), make_umma_desc(sV + ks * 512, v_lbo, v_sbo, 6),
                      id_o, ks ? 1u : 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    {
        uint32_t v[32];
        for (int c = 0; c < 128; c += 32) {
            tmem_ld32(tmem + lane_base + (uint32_t)c, v);
            tmem_ld_wait();
            for (int i = 0; i < 32; ++i) out.S[tid * 128 + c + i] = __uint_as_float(v[i]);
        }
        uint32_t w[16];
        for (int c = 0; c < DP; c += 16) {
            tmem_ld16(tmem + lane_base + 128u + (uint32_t)c, w);
            tmem_ld_wait();
            for (int i = 0; i < 16; ++i) out.O_ts[tid * DP + c + i] = __uint_as_float(w[i]);
            tmem_ld16(tmem + lane_base + 384u + (uint32_t)c, w);
            tmem_ld_wait();
            for (int i = 0; i < 16; ++i) out.O_ss[tid * DP + c + i] = __uint_as_float(w[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    std::vector<float> q(R * DP), k(R * DP), p(R * KV), v(KV * DP);
    srand(1);
    auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto& x : q) x = bf(rnd());
    for (auto& x : k) x = bf(rnd());
    for (auto& x : p) x = bf(fabsf(rnd()));
    for (auto& x : v) x = bf(rnd());
    std::vector<__nv_bfloat16> qb(q.size()), kb(k.size()), pb(p.size()), vb(v.size());
    for (size_t i = 0; i < q.size(); ++i) qb[i] = __float2bfloat16(q[i]);
    for (size_t i = 0; i < k.size(); ++i) kb[i] = __float2bfloat16(k[i]);
    for (size_t i = 0; i < p.size(); ++i) pb[i] = __float2bfloat16(p[i]);
    for (size_t i = 0; i < v.size(); ++i) vb[i] = __float2bfloat16(v[i]);
    std::vector<double> Sref(R * 128), Oref(R * DP);
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < 128; ++j) {
            double a = 0;
            for (int c = 0; c < DP; ++c) a += (double)q[i * DP + c] * k[j * DP + c];
            Sref[i * 128 + j] = a;
        }
    for (int i = 0; i < R; ++i)
        for (int n = 0; n < DP; ++n) {
            double a = 0;
            for (int c = 0; c < KV; ++c) a += (double)p[i * KV + c] * v[c * DP + n];
            Oref[i * DP + n] = a;
        }
    __nv_bfloat16 *dq, *dk, *dp, *dv;
    cudaMalloc(&dq, qb.size() * 2); cudaMalloc(&dk, kb.size() * 2); cudaMalloc(&dp, pb.size() * 2); cudaMalloc(&dv, vb.size() * 2);
    cudaMemcpy(dq, qb.data(), qb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dk, kb.data(), kb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dp, pb.data(), pb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dv, vb.data(), vb.size() * 2, cudaMemcpyHostToDevice);
    ProbeOut out;
    cudaMalloc(&out.S, R * 128 * 4); cudaMalloc(&out.O_ts, R * DP * 4); cudaMalloc(&out.O_ss, R * DP * 4);
    const int smem = 3 * kTileBytes + kPBytes + 1024 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const uint32_t variants[2][2] = {{4096, 256}, {256, 4096}};
    int rc = 1;
    for (int vi = 0; vi < 2; ++vi) {
        cudaMemset(out.S, 0, R * 128 * 4); cudaMemset(out.O_ts, 0, R * DP * 4); cudaMemset(out.O_ss, 0, R * DP * 4);
        probe_kernel<<<1, 128, smem>>>(dq, dk, dp, dv, out, variants[vi][0], variants[vi][1]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", vi, cudaGetErrorString(e)); return 2; }
        std::vector<float> S(R * 128), Ots(R * DP), Oss(R * DP);
        cudaMemcpy(S.data(), out.S, S.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(Ots.data(), out.O_ts, Ots.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(Oss.data(), out.O_ss, Oss.size() * 4, cudaMemcpyDeviceToHost);
        double es = 0, ets = 0, ess = 0;
        for (size_t i = 0; i < S.size(); ++i) es = fmax(es, fabs(S[i] - Sref[i]));
        for (size_t i = 0; i < Ots.size(); ++i) { ets = fmax(ets, fabs(Ots[i] - Oref[i])); ess = fmax(ess, fabs(Oss[i] - Oref[i])); }
        printf("variant %d (V lbo=%u sbo=%u): max|S err|=%.3e  max|O_ts err|=%.3e  max|O_ss err|=%.3e\n",
               vi, variants[vi][0], variants[vi][1], es, ets, ess);
        if (es < 1e-3 && ets < 1e-3 && ess < 1e-3) rc = 0;
    }
    printf(rc == 0 ? "PROBE OK\n" : "PROBE FAILED\n");
    return rc;
}
