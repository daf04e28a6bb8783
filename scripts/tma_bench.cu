// Micro-benchmark behind the attention tile layouts: how fast ONE SM's TMA unit fills shared memory from L2-resident
// data as a function of the box shape -- the attention kernels load [rows x head_dim] operand tiles from the strided
// [tokens, 3H] matrix as 16-column (32-byte, SWIZZLE_32B) boxes; a 64-byte or 128-byte inner box moves the same tile in
// 2x / 4x fewer box rows.  Per configuration every CTA (one per SM) keeps DEPTH loads in flight from one thread and
// reports bytes / clock / SM and clocks per box row; the loads are issued by 1 thread, by several lanes of one warp, or by
// one lane of several warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ideco_b200/csrc scripts/tma_bench.cu -o build/tma_bench
#include "tcgen05.cuh"
#include "tma_host.cuh"
#include <cstdlib>
#include <vector>

void deco_set_error(const char*, ...) {}
using namespace deco;


template <int kDepth>
__global__ void __launch_bounds__(128, 1) tma_bench_kernel(const __grid_constant__ CUtensorMap map, int box_cols, int box_rows,
                                                           int n_cols, int n_rows, int iters, int lanes, int warps,
                                                           long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t box_bytes = (uint32_t)box_cols * box_rows * 2;
    const uint32_t slot_bytes = (box_bytes + 1023u) & ~1023u;
    const int nthr = lanes * warps;                                 // issuing threads, each with its own ring of kDepth slots
    const uint32_t bars0 = base + nthr * kDepth * slot_bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nthr * kDepth; ++i) mbar_init(bars0 + 8 * i, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (w < warps && l < lanes) {
        const int me = w * lanes + l;
        const uint32_t bars = bars0 + 8 * kDepth * me;
        const uint32_t mybase = base + me * kDepth * slot_bytes;
        // power-of-two ranges: an integer division per load would dominate the loop
        const uint32_t col_mask = 31u, row_mask = (uint32_t)(n_rows / box_rows) - 1u;
        uint32_t rng = (blockIdx.x * 64 + me) * 2654435761u + 12345u;
        auto issue = [&](int i) {
            rng = rng * 1664525u + 1013904223u;
            const int cb = (int)((rng >> 8) & col_mask), rb = (int)((rng >> 16) & row_mask);
            const int s = i % kDepth;
            mbar_expect_tx(bars + 8 * s, box_bytes);
            tma_load_2d(mybase + s * slot_bytes, &map, bars + 8 * s, cb * box_cols, rb * box_rows);
        };
        for (int i = 0; i < kDepth; ++i) issue(i);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            mbar_wait(bars + 8 * (i % kDepth), (uint32_t)((i / kDepth) & 1));
            issue(i + kDepth);
        }
        const long long t1 = clock64();
        for (int i = iters; i < iters + kDepth; ++i) mbar_wait(bars + 8 * (i % kDepth), (uint32_t)((i / kDepth) & 1));
        if (me == 0) out[blockIdx.x] = t1 - t0;
    }
}

// The attention kernels' access pattern: 4-D map (d, token, head, image) over a [tokens, 3 * heads * pitch] matrix, boxes of
// 16 columns x ROWS tokens of one head; `lanes` x `warps` issuing threads, each with kDepth loads in flight.
template <int kDepth>
__global__ void __launch_bounds__(256, 1) tma_bench4d_kernel(const __grid_constant__ CUtensorMap map, int box_rows, int nch, int L, int heads,
                                                             int B, int iters, int lanes, int warps, long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t box_bytes = 16u * box_rows * 2;
    const uint32_t slot_bytes = (box_bytes + 1023u) & ~1023u;
    const int nthr = lanes * warps;
    const uint32_t bars0 = base + nthr * kDepth * slot_bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nthr * kDepth; ++i) mbar_init(bars0 + 8 * i, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (w < warps && l < lanes) {
        const int me = w * lanes + l;
        const uint32_t bars = bars0 + 8 * kDepth * me;
        const uint32_t mybase = base + me * kDepth * slot_bytes;
        uint32_t rng = (blockIdx.x * 64 + me) * 2654435761u + 12345u;
        const uint32_t rb_mask = (uint32_t)(L / box_rows) - 1u, h_mask = (uint32_t)heads - 1u, b_mask = (uint32_t)B - 1u;
        int c = l % nch;
        auto issue = [&](int i) {
            rng = rng * 1664525u + 1013904223u;
            const int s = i % kDepth;
            mbar_expect_tx(bars + 8 * s, box_bytes);
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                :: "r"(mybase + s * slot_bytes), "l"(&map), "r"(bars + 8 * s), "r"(16 * c), "r"((int)((rng >> 8) & rb_mask) * box_rows),
                   "r"((int)((rng >> 16) & h_mask)), "r"((int)((rng >> 20) & b_mask)) : "memory");
        };
        for (int i = 0; i < kDepth; ++i) issue(i);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            mbar_wait(bars + 8 * (i % kDepth), (uint32_t)((i / kDepth) & 1));
            issue(i + kDepth);
        }
        const long long t1 = clock64();
        for (int i = iters; i < iters + kDepth; ++i) mbar_wait(bars + 8 * (i % kDepth), (uint32_t)((i / kDepth) & 1));
        if (me == 0) out[blockIdx.x] = t1 - t0;
    }
}

int main() {
    const int n_rows = 8192, n_cols = 3456;       // 56 MB of bf16: L2-resident after the first pass
    __nv_bfloat16* src;
    cudaMalloc(&src, (size_t)n_rows * n_cols * 2);
    cudaMemset(src, 0, (size_t)n_rows * n_cols * 2);
    long long* out;
    cudaMalloc(&out, 148 * sizeof(long long));
    PFN_encodeTiled enc = get_tensormap_encoder();
    if (!enc) { printf("no encoder\n"); return 1; }
    cudaFuncSetAttribute(tma_bench_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(tma_bench_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { int cols, rows; CUtensorMapSwizzle sw; const char* name; };
    const Cfg cfgs[] = {
        {16, 128, CU_TENSOR_MAP_SWIZZLE_32B, "16 cols x 128 rows, SW32 "}, {16, 64, CU_TENSOR_MAP_SWIZZLE_32B, "16 cols x  64 rows, SW32 "},
        {32, 128, CU_TENSOR_MAP_SWIZZLE_64B, "32 cols x 128 rows, SW64 "}, {32, 64, CU_TENSOR_MAP_SWIZZLE_64B, "32 cols x  64 rows, SW64 "},
        {64, 128, CU_TENSOR_MAP_SWIZZLE_128B, "64 cols x 128 rows, SW128"}, {64, 64, CU_TENSOR_MAP_SWIZZLE_128B, "64 cols x  64 rows, SW128"},
        {64, 32, CU_TENSOR_MAP_SWIZZLE_128B, "64 cols x  32 rows, SW128"},
    };
    struct Mode { int depth, lanes, warps; };
    const Mode modes[] = {{8, 1, 1}, {24, 1, 1}, {8, 4, 1}, {8, 1, 4}, {8, 5, 2}};
    for (const Mode& m : modes) {
    const int kDepth = m.depth;
    printf("TMA fill rate of one SM from L2-resident data (148 CTAs; %d issuing lane(s) x %d warp(s), %d loads in flight per thread)\n",
           m.lanes, m.warps, kDepth);
    for (const Cfg& c : cfgs) {
        if ((size_t)kDepth * m.lanes * m.warps * c.cols * c.rows * 2 > 190 * 1024) continue;
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)n_cols, (cuuint64_t)n_rows};
        cuuint64_t strides[1] = {(cuuint64_t)n_cols * 2};
        cuuint32_t box[2] = {(cuuint32_t)c.cols, (cuuint32_t)c.rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int iters = 2000;
        const size_t smem = (size_t)kDepth * m.lanes * m.warps * (((size_t)c.cols * c.rows * 2 + 1023) & ~(size_t)1023) + 1024 + 8 * 512;
        for (int rep = 0; rep < 2; ++rep) {
            if (kDepth == 8) tma_bench_kernel<8><<<148, 128, smem>>>(map, c.cols, c.rows, n_cols, n_rows, iters, m.lanes, m.warps, out);
            else tma_bench_kernel<24><<<148, 128, smem>>>(map, c.cols, c.rows, n_cols, n_rows, iters, m.lanes, m.warps, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        }
        std::vector<long long> h(148);
        cudaMemcpy(h.data(), out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (long long v : h) avg += (double)v;
        avg /= 148;
        const double bytes = (double)iters * c.cols * c.rows * 2 * m.lanes * m.warps;
        printf("  box %s: %7.1f bytes/clk/SM, %6.2f clk per box row, %7.0f clk per box\n", c.name, bytes / avg,
               avg / ((double)iters * c.rows * m.lanes * m.warps), avg / ((double)iters * m.lanes * m.warps));
    }
    }
    {
        const int L = 256, heads = 16, Bn = 32, D = 72;
        for (int pitch : {72, 80}) {
            const long long rs = 3LL * heads * pitch;
            __nv_bfloat16* q;
            cudaMalloc(&q, (size_t)Bn * L * rs * 2);
            cudaMemset(q, 0, (size_t)Bn * L * rs * 2);
            cudaFuncSetAttribute(tma_bench4d_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            for (int rows : {64, 128}) for (int wl : {0, 1, 2}) {
                const int lanes = wl == 0 ? 1 : 5, warps = wl == 2 ? 4 : (wl == 1 ? 2 : 1);
                CUtensorMap map;
                cuuint64_t dims[4] = {(cuuint64_t)(pitch == 72 ? D : pitch), (cuuint64_t)L, (cuuint64_t)heads, (cuuint64_t)Bn};
                cuuint64_t strides[3] = {(cuuint64_t)rs * 2, (cuuint64_t)pitch * 2, (cuuint64_t)L * rs * 2};
                cuuint32_t box[4] = {16, (cuuint32_t)rows, 1, 1};
                cuuint32_t estr[4] = {1, 1, 1, 1};
                CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, q, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encode4d failed %d\n", (int)r); return 1; }
                const int iters = 1000, nthr = lanes * warps;
                const size_t smem = (size_t)8 * nthr * (((size_t)16 * rows * 2 + 1023) & ~(size_t)1023) + 1024 + 8 * 8 * 32;
                if (smem > 200 * 1024) continue;
                for (int rep = 0; rep < 2; ++rep) {
                    tma_bench4d_kernel<8><<<148, 256, smem>>>(map, rows, 5, L, heads, Bn, iters, lanes, warps, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("kernel4d failed: %s\n", cudaGetErrorString(e)); return 1; }
                }
                std::vector<long long> h(148);
                cudaMemcpy(h.data(), out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
                double avg = 0;
                for (long long v : h) avg += (double)v;
                avg /= 148;
                printf("  4-D map, head pitch %d, box 16 cols x %3d tokens, %d lane(s) x %d warp(s): %6.2f clk per box row, %6.0f clk per box\n",
                       pitch, rows, lanes, warps, avg / ((double)iters * rows * nthr), avg / ((double)iters * nthr));
            }
            cudaFree(q);
        }
    }
    return 0;
}
