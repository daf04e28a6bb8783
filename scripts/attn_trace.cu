// Per-event clock64 timeline of CTA 0 of the attention kernel (XL/16 256 px shape).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDECO_ATTN_TRACE -Ideco_b200/csrc scripts/attn_trace.cu -o build/attn_trace
#include "../deco_b200/csrc/attention_tc.cu"
#include <cstdarg>
#include <vector>
void deco_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }

int main() {
    const int B = 512, heads = 16, d = 72, L = 256, H = heads * d;
    const size_t n = (size_t)B * L * 3 * H;
    std::vector<__nv_bfloat16> h(n);
    srand(1);
    for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16((float)rand() / RAND_MAX * 2.f - 1.f);
    __nv_bfloat16 *qkv, *out;
    cudaMalloc(&qkv, n * 2); cudaMalloc(&out, (size_t)B * L * H * 2);
    cudaMemcpy(qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
    for (int it = 0; it < 3; ++it) {
        int rc = deco_attention_fwd(qkv, 3 * H, qkv + H, qkv + 2 * H, 3 * H, L, nullptr, nullptr, 0, 0, out, H, B, heads, L, d, 0.1178511f, nullptr);
        if (rc) return 1;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    }
    std::vector<long long> tr(3 * 2048);
    cudaMemcpyFromSymbol(tr.data(), deco::g_attn_trace, tr.size() * 8);
    long long t0 = tr[0];
    for (int reg = 0; reg < 3; ++reg) {
        printf("region %d\n", reg);
        for (int i = 0; i < 40; ++i) printf("  %8lld code %lld\n", tr[reg * 2048 + 2 * i] - t0, tr[reg * 2048 + 2 * i + 1]);
    }
    return 0;
}
