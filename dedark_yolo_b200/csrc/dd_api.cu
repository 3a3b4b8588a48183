// dd_api.cu -- library-wide plumbing of libdedark_b200.so: version, error string, launch counter,
// workspace sizing.  The compute entry points live next to their kernels (dd_synth.cu,
// dd_predictor.cu, dd_recovery.cu).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "dd_common.cuh"
#include "dd_layout.cuh"

namespace dd {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DD_ERR_CUDA;
    }
    return DD_OK;
}

}  // namespace dd

extern "C" {

int dd_version(void) { return 100; }  // 0.1.0

const char* dd_last_error(void) { return dd::g_err; }

unsigned long long dd_launch_count(void) { return dd::g_launches.load(std::memory_order_relaxed); }

size_t dd_workspace_bytes(int kind, int B, int H, int W) {
    if (B <= 0) return 0;
    switch (kind) {
        case DD_WS_SYNTH: return dd::synth_ws_bytes(B);
        case DD_WS_PREDICTOR_ACTS: return dd::predictor_acts_bytes(B);
        case DD_WS_PREDICTOR_BWD: return dd::predictor_bwd_ws_bytes(B);
        case DD_WS_RECOVERY_BWD: return (H > 0 && W > 0) ? dd::recovery_bwd_ws_bytes(B, H, W) : 0;
        case DD_WS_DARK_PRIOR: return dd::prior_ws_bytes(B);
        default: return 0;
    }
}

}  // extern "C"
