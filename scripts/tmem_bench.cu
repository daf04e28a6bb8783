// Micro-benchmark behind the pixel-decoder design (csrc/decoder_tc.cu): tensor-memory read / write bandwidth per SM as a
// function of the number of warps, and the issue rate of small-N tcgen05.mma (M = 128, N in {8, 32, 64, 96}, K = 16) with
// the A operand in tensor memory (TS) or shared memory (SS).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ideco_b200/csrc scripts/tmem_bench.cu -o build/tmem_bench
#include "tcgen05.cuh"
#include <cstdlib>
#include <vector>

void deco_set_error(const char*, ...) {}
using namespace deco;

struct Res { long long ld[5], st[5], mma_ts[4], mma_ss[4]; };

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
    const uint32_t sB = base, sA = base + 16384, bar = base + 65536, slot = bar + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;   // small bf16 values
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
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

    // ---- MMA issue rate
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const int Ns[4] = {16, 32, 64, 96};
        uint32_t phase = 0;
        for (int mode = 0; mode < 2; ++mode)
            for (int ni = 0; ni < 4; ++ni) {
                const uint32_t idesc = make_idesc_major(128, Ns[ni], 0, 0);
                const uint64_t db = make_umma_desc(sB, 16, 256, 6);
                const uint64_t da = make_umma_desc(sA, 16, 256, 6);
                long long t0 = clock64();
                for (int it = 0; it < iters; ++it) {
                    // round-robin over independent accumulators (as many as fit in 256 columns): issue rate, not latency
                    const int nacc = 256 / Ns[ni];
                    const uint32_t d = tmem + 256u + (uint32_t)((it % nacc) * Ns[ni]);
                    if (mode == 0) umma_bf16_ts(d, tmem + (uint32_t)((it & 3) * 8), db, idesc, 1u);
                    else umma_bf16(d, da, db, idesc, 1u);
                }
                umma_commit(bar);
                mbar_wait(bar, phase);
                phase ^= 1;
                long long t1 = clock64();
                (mode == 0 ? out->mma_ts : out->mma_ss)[ni] = t1 - t0;
            }
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
    const int Ns[4] = {16, 32, 64, 96};
    for (int i = 0; i < 4; ++i)
        printf("tcgen05.mma M128 N%-2d K16: TS %6.1f clk/instr (%6.0f MAC/clk)   SS %6.1f clk/instr (%6.0f MAC/clk)\n", Ns[i],
               (double)h.mma_ts[i] / iters, 128.0 * Ns[i] * 16 * iters / h.mma_ts[i],
               (double)h.mma_ss[i] / iters, 128.0 * Ns[i] * 16 * iters / h.mma_ss[i]);
    return 0;
}
