// Error plumbing of the C ABI and the small building-block entry points exposed for unit tests.
#include <stdarg.h>

#include <map>
#include <set>
#include <mutex>
#include <string>
#include <vector>

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

#ifndef INTEL_EMU
void ensure_smem_impl(const void* kernel, size_t smem) {
    static std::mutex mu;
    static std::map<const void*, size_t> done;
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find(kernel);
    if (it != done.end() && it->second >= smem) return;
    // opt in to the B200 maximum once (227 KB per CTA, shared between the kernel's static and dynamic parts)
    cudaFuncAttributes fa;
    size_t want = 227 * 1024;
    if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess && fa.sharedSizeBytes < want) want -= fa.sharedSizeBytes;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > want ? smem : want));
    // ask for the largest shared-memory carve-out so that as many CTAs as the request allows are co-resident
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    done[kernel] = smem > want ? smem : want;
}

struct ProfRec { cudaEvent_t a, b; const char* what; double bytes, flops; };
static bool g_prof_on = false, g_prof_pending = false, g_prof_detail = false;
static cudaEvent_t g_pa, g_pb;
static std::vector<ProfRec> g_recs;

void prof_before(cudaStream_t s) {
    if (!g_prof_on) return;
    cudaEventCreate(&g_pa);
    cudaEventCreate(&g_pb);
    cudaEventRecord(g_pa, s);
    g_prof_pending = true;
}
void prof_mark(cudaStream_t s) {
    if (g_prof_on && g_prof_pending) cudaEventRecord(g_pb, s);
}
static void prof_commit(const char* what, double bytes, double flops) {
    if (!g_prof_on || !g_prof_pending) return;
    g_recs.push_back({g_pa, g_pb, what, bytes, flops});
    g_prof_pending = false;
}
bool prof_detail() { return g_prof_on && g_prof_detail; }
const char* prof_intern(const char* name) {
    static std::set<std::string> names;
    return names.insert(name).first->c_str();
}
#else
void prof_before(cudaStream_t) {}
void prof_mark(cudaStream_t) {}
static void prof_commit(const char*, double, double) {}
bool prof_detail() { return false; }
const char* prof_intern(const char* name) { return name; }
#endif

int check_launch(const char* what, double bytes, double flops) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return INTEL_ERR_CUDA;
    }
    prof_commit(what, bytes, flops);
    return INTEL_OK;
}

}  // namespace intel

using namespace intel;

extern "C" {

const char* intel_last_error(void) { return g_err; }
int intel_abi_version(void) { return INTEL_ABI_VERSION; }

int intel_profile_enable(int on) {
#ifndef INTEL_EMU
    g_prof_on = on != 0;
    g_prof_detail = on == 2;       // 2: GEMM launches are reported per shape
    g_prof_pending = false;
#endif
    (void)on;
    return INTEL_OK;
}

// Synchronises the device (the only call of the library that does), aggregates the recorded launches by
// kernel name and writes one line per name: "name launches total_ms bytes flops".
int intel_profile_report(char* buf, size_t cap) {
    if (!buf || cap == 0) return INTEL_ERR_ARG;
    buf[0] = 0;
#ifndef INTEL_EMU
    cudaError_t e = cudaDeviceSynchronize();
    INTEL_REQUIRE(e == cudaSuccess, INTEL_ERR_CUDA, "profile_report: %s", cudaGetErrorString(e));
    struct Agg { double n = 0, ms = 0, bytes = 0, flops = 0; };
    std::map<std::string, Agg> agg;
    for (auto& r : g_recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        Agg& a = agg[r.what];
        a.n += 1; a.ms += ms; a.bytes += r.bytes; a.flops += r.flops;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_recs.clear();
    size_t off = 0;
    for (auto& kv : agg) {
        int w = snprintf(buf + off, cap - off, "%s %.0f %.6f %.0f %.0f\n", kv.first.c_str(), kv.second.n, kv.second.ms,
                         kv.second.bytes, kv.second.flops);
        if (w < 0 || (size_t)w >= cap - off) break;
        off += (size_t)w;
    }
#endif
    return INTEL_OK;
}

int intel_awelv_fwd(int64_t B, int64_t L, int K, int h, const float* user_table, const float* model_table, const int64_t* u_id,
                    const double* scores, float* weights, float* ens_score, float* w_user, intel_stream_t stream) {
    return awelv_fwd(B, L, K, h, user_table, model_table, u_id, scores, weights, ens_score, w_user, (cudaStream_t)stream);
}

int intel_awelv_bwd(int64_t B, int64_t L, int K, int h, const float* user_table, const float* model_table, const int64_t* u_id,
                    const double* scores, const float* w_user, const float* d_weights, const float* d_ens, float* g_user_table,
                    float* g_model_table, intel_stream_t stream) {
    return awelv_bwd(B, L, K, h, user_table, model_table, u_id, scores, w_user, d_weights, d_ens, g_user_table, g_model_table,
                     (cudaStream_t)stream);
}

int intel_pool_head_fwd(int64_t B, int64_t L, int K, const float* slot_weights, const double* scores, float* weights,
                        float* ens_score, float* p_sess, float* w_sess, intel_stream_t stream) {
    return pool_head_fwd(B, L, K, slot_weights, scores, weights, ens_score, p_sess, w_sess, (cudaStream_t)stream);
}

int intel_pool_head_bwd(int64_t B, int64_t L, int K, const double* scores, const float* p_sess, const float* w_sess,
                        const float* d_weights, const float* d_ens, float* d_slot_weights, intel_stream_t stream) {
    return pool_head_bwd(B, L, K, scores, p_sess, w_sess, d_weights, d_ens, d_slot_weights, (cudaStream_t)stream);
}

int intel_batch_validate(const intel_dims_t* d, const intel_batch_t* bt, int32_t* flags, intel_stream_t stream) {
    INTEL_REQUIRE(d && bt, INTEL_ERR_ARG, "intel_batch_validate: null argument");
    ValidateArgs a;
    a.B = d->B; a.L = d->L; a.H1 = d->H1; a.H2 = d->H2; a.I = d->I;
    a.item_rows = d->item_rows; a.class_rows = d->class_rows; a.user_rows = d->user_rows; a.ctx_rows = d->ctx_rows;
    a.u_id = bt->u_id; a.i_id = bt->i_id; a.i_class = bt->i_class; a.session_len = bt->session_len;
    a.context_mh = bt->context_mh; a.his_context = bt->his_context; a.history_len = bt->history_len;
    a.his_item_id = bt->his_item_id; a.history_item_len = bt->history_item_len;
    a.idx1 = bt->his_intents_idx; a.idx2 = bt->his_item_int_idx; a.nz1 = bt->nz1; a.nz2 = bt->nz2;
    return batch_validate(a, flags, (cudaStream_t)stream);
}

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

int intel_linear_dx(int64_t M, int64_t N, int64_t K, const float* dY, const float* W, float* dX, const float* relu_mask,
                    intel_stream_t stream) {
    return linear_dx(M, N, K, dY, N, W, K, dX, K, (cudaStream_t)stream, 0, relu_mask, K);
}

int intel_linear_dw(int64_t M, int64_t N, int64_t K, const float* dY, const float* X, float* dW, float* db,
                    intel_stream_t stream) {
    return linear_dw(M, N, K, dY, N, X, K, dW, K, db, (cudaStream_t)stream, false);
}

int intel_linear_fwd_ex(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* W, const float* bias,
                        float* C, int64_t ldc, int relu_a, intel_stream_t stream) {
    return linear(M, N, K, A, lda, W, K, bias, C, ldc, (cudaStream_t)stream, relu_a != 0, false);
}

int intel_linear_dx_ex(int64_t M, int64_t N, int64_t K, const float* dY, const float* W, float* dX, int64_t lddx,
                       const float* relu_mask, int64_t ldmask, intel_stream_t stream) {
    return linear_dx(M, N, K, dY, N, W, K, dX, lddx, (cudaStream_t)stream, 0, relu_mask, ldmask);
}

int intel_linear_dw_ex(int64_t M, int64_t N, int64_t K, const float* dY, const float* X, int64_t ldx, float* dW, float* db,
                       int relu_x, intel_stream_t stream) {
    return linear_dw(M, N, K, dY, N, X, ldx, dW, K, db, (cudaStream_t)stream, relu_x != 0);
}

int intel_softmax_rows_fwd(int64_t R, int64_t N, const float* Z, float* P, intel_stream_t stream) {
    return softmax_rows(R, N, Z, P, (cudaStream_t)stream);
}

int intel_softmax_rows_bwd(int64_t R, int64_t N, const float* P, const float* dP, float* dZ, intel_stream_t stream) {
    return softmax_rows_bwd(R, N, P, dP, nullptr, dZ, (cudaStream_t)stream);
}

int intel_mha_fwd(int64_t B, int64_t T, int d, int heads, const float* qkv, const int64_t* lens, float* out,
                  intel_stream_t stream) {
    return mha_fwd(B, T, d, heads, qkv, lens, out, (cudaStream_t)stream);
}

int intel_mha_bwd(int64_t B, int64_t T, int d, int heads, const float* qkv, const int64_t* lens, const float* d_out,
                  float* d_qkv, intel_stream_t stream) {
    return mha_bwd(B, T, d, heads, qkv, lens, d_out, d_qkv, (cudaStream_t)stream);
}

}  // extern "C"
