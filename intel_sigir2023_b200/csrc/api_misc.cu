// Error plumbing of the C ABI and the small building-block entry points exposed for unit tests.
#include <stdarg.h>

#include "kernels.h"
#include "../../include/intel_b200.h"

namespace intel {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return INTEL_ERR_CUDA;
    }
    return INTEL_OK;
}

}  // namespace intel

using namespace intel;

extern "C" {

const char* intel_last_error(void) { return g_err; }
int intel_abi_version(void) { return INTEL_ABI_VERSION; }

int intel_gather_fwd(int64_t rows, int d, const float* table, const int64_t* idx, float* out, int ld_out, int relu,
                     intel_stream_t stream) {
    return gather_rows(rows, d, table, idx, out, ld_out, relu, (cudaStream_t)stream);
}

int intel_scatter_add_bwd(int64_t rows, int d, const float* d_out, int ld, const int64_t* idx, float* grad_table,
                          intel_stream_t stream) {
    return scatter_add_rows(rows, d, d_out, ld, idx, grad_table, nullptr, (cudaStream_t)stream);
}

int intel_linear_fwd(int64_t M, int64_t N, int64_t K, const float* A, const float* W, const float* bias, float* C,
                     intel_stream_t stream) {
    return linear(M, N, K, A, K, W, K, bias, C, N, (cudaStream_t)stream);
}

}  // extern "C"
