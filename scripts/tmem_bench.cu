// Micro-benchmark behind the pixel-decoder design (csrc/decoder_tc.cu): tensor-memory read / write bandwidth per SM as a
// function of the number of warps, and the issue rate of small-N tcgen05.mma (M = 128, N in {8, 32, 64, 96}, K = 16) with
// the A operand in tensor memory (TS) or shared memory (SS).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ideco_b200/csrc scripts/tmem_bench.cu -o build/tmem_bench
#include "tcgen05.cuh"
#include <cstdlib>
#include <vector>
#include <type_traits>

void deco_set_error(const char*, ...) {}
using namespace deco;

struct Res { long long ld[5], st[5], mma[24], multi[3]; };

template <int WARPS>
__device__ void bench_ld_st(uint32_t tmem, int warp, int lane, long long* ld_out, long long* st_out, int iters, float* sink) {
    // every warp reads / writes its own lane quarter (warp % 4); warps sharing a quarter use different columns
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = lane + i;
    __syncthreads();
    long long t0 = clock64();
    if (warp < WARPS) {
        for (int it = 0; it < iters; ++it) {
            tmem_st32(base + (uint32_t)((it & 1) * 32), v);
        }
        tmem_st_wait();
    }
    __syncthreads();
    long long t1 = clock64();
    float acc = 0.f;
    if (warp < WARPS) {
        for (int it = 0; it < iters; ++it) {
            uint32_t r[32];
            tmem_ld32(base + (uint32_t)((it & 1) * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 8) acc += __uint_as_float(r[i]);
        }
    }
    __syncthreads();
    long long t2 = clock64();
    if (threadIdx.x == 0) { *st_out = t1 - t0; *ld_out = t2 - t1; }
    if (acc == 123.456f) *sink = acc;
}

__global__ void __launch_bounds__(512, 1) bench_kernel(Res* out, int iters, float* sink)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sB = base, sA = base + 16384, bar = base + 65536, slot = bar + 64;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;   // small bf16 values
    if (tid == 0) { for (int i = 0; i < 6; ++i) mbar_init(bar + 8u * i, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));

    bench_ld_st<1>(tmem, warp, lane, &out->ld[0], &out->st[0], iters, sink);
    bench_ld_st<2>(tmem, warp, lane, &out->ld[1], &out->st[1], iters, sink);
    bench_ld_st<4>(tmem, warp, lane, &out->ld[2], &out->st[2], iters, sink);
    bench_ld_st<8>(tmem, warp, lane, &out->ld[3], &out->st[3], iters, sink);
    bench_ld_st<16>(tmem, warp, lane, &out->ld[4], &out->st[4], iters, sink);

    // ---- MMA issue rate: unrolled groups of 8 instructions with precomputed descriptors, (a) round-robin over independent
    // accumulators, (b) all accumulating into ONE accumulator (dependent chain), TS and SS forms
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        uint32_t phase = 0;
        const uint64_t db = make_umma_desc(sB, 16, 256, 6);
        const uint64_t da = make_umma_desc(sA, 16, 256, 6);
        auto run = [&](auto ntag, int mode, bool chain, long long* dst) {
            constexpr int N = decltype(ntag)::value;
            constexpr uint32_t idesc = make_idesc_major(128, N, 0, 0);
            constexpr int nacc = 256 / N > 8 ? 8 : (256 / N < 1 ? 1 : 256 / N);
            long long t0 = clock64();
            for (int it = 0; it < iters / 8; ++it) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t d = tmem + 256u + (chain ? 0u : (uint32_t)((u % nacc) * N));
                    if (mode == 0) umma_bf16_ts(d, tmem + (uint32_t)((u & 3) * 8), db, idesc, 1u);
                    else umma_bf16(d, da, db, idesc, 1u);
                }
            }
            umma_commit(bar);
            mbar_wait(bar, phase);
            phase ^= 1;
            *dst = clock64() - t0;
        };
        int o = 0;
        for (int mode = 0; mode < 2; ++mode)
            for (int chain = 0; chain < 2; ++chain) {
                run(std::integral_constant<int, 16>{}, mode, chain, &out->mma[o++]);
                run(std::integral_constant<int, 32>{}, mode, chain, &out->mma[o++]);
                run(std::integral_constant<int, 64>{}, mode, chain, &out->mma[o++]);
                run(std::integral_constant<int, 96>{}, mode, chain, &out->mma[o++]);
                run(std::integral_constant<int, 128>{}, mode, chain, &out->mma[o++]);
                run(std::integral_constant<int, 256>{}, mode, chain, &out->mma[o++]);
            }
    }
    // ---- is the ~94 clk per instruction a per-THREAD issue cost or a tensor-pipe cost?  1 / 2 / 4 warps issue N = 32 TS MMAs
    // at the same time (own accumulators, own barriers)
    int uses = 0;
    for (int nw = 1; nw <= 4; nw *= 2) {
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        long long t0 = clock64();
        if (warp < nw && lane == 0) {
            constexpr uint32_t idesc = make_idesc_major(128, 32, 0, 0);
            const uint64_t db = make_umma_desc(sB, 16, 256, 6);
            for (int it = 0; it < iters / 8; ++it) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    umma_bf16_ts(tmem + 256u + (uint32_t)(warp * 64 + (u & 1) * 32), tmem + (uint32_t)((u & 3) * 8), db, idesc, 1u);
            }
            umma_commit(bar + 8u * (1 + warp));
            mbar_wait(bar + 8u * (1 + warp), (uint32_t)(uses & 1));
            ++uses;
        }
        __syncthreads();
        if (tid == 0) out->multi[nw == 1 ? 0 : nw == 2 ? 1 : 2] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    Res* d; float* sink;
    cudaMalloc(&d, sizeof(Res)); cudaMalloc(&sink, 4);
    const int iters = 2000;
    cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    for (int rep = 0; rep < 2; ++rep) bench_kernel<<<1, 512, 80 * 1024>>>(d, iters, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    Res h; cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const int W[5] = {1, 2, 4, 8, 16};
    for (int i = 0; i < 5; ++i) {
        const double bytes = (double)iters * W[i] * 32 * 32 * 4;
        printf("warps %2d: tcgen05.ld 32x32b.x32  %7.1f B/clk/SM (%5.1f clk per instr per warp)   tcgen05.st %7.1f B/clk/SM\n",
               W[i], bytes / h.ld[i], (double)h.ld[i] / iters, bytes / h.st[i]);
    }
    const int Ns[6] = {16, 32, 64, 96, 128, 256};
    const char* names[4] = {"TS independent", "TS chained    ", "SS independent", "SS chained    "};
    for (int m = 0; m < 4; ++m)
        for (int i = 0; i < 6; ++i) {
            const double clk = (double)h.mma[m * 6 + i] / iters;
            printf("tcgen05.mma M128 N%-3d K16 %s: %6.1f clk/instr  (%5.0f MAC/clk; floor 128*N/256 = %d clk)\n", Ns[i], names[m], clk,
                   128.0 * Ns[i] * 16 / clk, 128 * Ns[i] / 256);
        }
    for (int i = 0; i < 3; ++i)
        printf("N32 TS MMAs issued by %d warps concurrently: %6.1f clk per instruction per warp, %6.1f clk per instruction overall\n",
               1 << i, (double)h.multi[i] / iters, (double)h.multi[i] / iters / (1 << i));
    return 0;
}
