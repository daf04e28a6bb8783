// Error reporting and version for the C-ABI library (include/deco_b200.h).
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

static thread_local char g_err[512] = "";

void deco_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* deco_last_error(void) { return g_err; }
extern "C" int deco_abi_version(void) { return 1; }

// Number of kernels the library has launched since load is not tracked here; bench.py counts launches on the host side.

// Programmatic dependent launch for the big kernels of the sampling step (GEMMs, attention, pixel decoder): on by default,
// DECO_B200_PDL=0 turns it off (A/B measurements).
bool deco_pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DECO_B200_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

// SMs the persistent GEMM kernels leave free (process-wide, default 0): while NCCL's kernels average gradients on another
// stream (deco_b200.distributed.OverlappedGradientAverager) a persistent GEMM sized for every SM would wait, with its static
// tile order, for the SMs NCCL holds; sized for (SMs - reserved) the two run side by side.
static int g_reserved_sms = 0;
int deco_reserved_sms() { return g_reserved_sms; }
extern "C" int deco_gemm_reserve_sms(int n) {
    if (n < 0 || n > 64) { deco_set_error("gemm_reserve_sms: %d out of range [0, 64]", n); return DECO_ERR_ARG; }
    g_reserved_sms = n & ~1;
    return DECO_OK;
}
