// Host-side input packing for the end-to-end path.  The reference hands the history intents to the model as dense
// float64 [B,H,I] tensors (collate_batch; 1.4 GB per 4096 sessions, >99 % zeros), which makes the host->device copy
// the slowest stage of a step by 4x.  This routine scans the dense rows with all host cores and emits the compact
// (index, value) form the kernels already accept (intel_batch_t.his_intents_idx / _val), so only the non-zeros
// cross PCIe.  Pure data movement: values are cast to fp32 exactly as dense_rows_fwd_kernel does on the device.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>
#if defined(__x86_64__) && !defined(INTEL_EMU)
#include <immintrin.h>
#define INTEL_HOST_AVX2 1
#endif

#include "common.cuh"

using namespace intel;

namespace {

#ifdef INTEL_HOST_AVX2
// first index c >= c0 (multiple of 8) whose 8-double block holds a non-zero, or the last multiple of 8 <= I
__attribute__((target("avx2"))) int64_t skip_zero_blocks_avx2(const uint64_t* row, int64_t c, int64_t I) {
    const __m256i absmask = _mm256_set1_epi64x(0x7fffffffffffffffLL);
    for (; c + 8 <= I; c += 8) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(row + c));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(row + c + 4));
        if (!_mm256_testz_si256(_mm256_or_si256(a, b), absmask)) break;
    }
    return c;
}
#endif

// rows [r0, r1): returns the largest non-zero count seen
int32_t pack_block(int64_t r0, int64_t r1, int64_t I, const double* dense, int32_t nz, int32_t* idx, float* val, int32_t* row_nnz,
                   int64_t group, const int64_t* group_len) {
    int32_t worst = 0;
    const uint64_t* bits = reinterpret_cast<const uint64_t*>(dense);
    for (int64_t r = r0; r < r1; ++r) {
        const uint64_t* row = bits + r * I;
        const double* drow = dense + r * I;
        int32_t* oi = idx + r * nz;
        float* ov = val + r * nz;
        int32_t n = 0;
        int64_t c = 0;
        if (group > 0 && (r % group) >= group_len[r / group]) c = I;     // padding row of a ragged group: not read at all
#ifdef INTEL_HOST_AVX2
        static const bool avx2 = __builtin_cpu_supports("avx2");
#endif
        for (; c + 8 <= I; c += 8) {            // +0.0 and -0.0 both vanish after the shift
#ifdef INTEL_HOST_AVX2
            if (avx2) {
                c = skip_zero_blocks_avx2(row, c, I);
                if (c + 8 > I) break;
            }
#endif
            const uint64_t any = (row[c] | row[c + 1] | row[c + 2] | row[c + 3] | row[c + 4] | row[c + 5] | row[c + 6] | row[c + 7]) << 1;
            if (any == 0) continue;
            for (int j = 0; j < 8; ++j) {
                if ((row[c + j] << 1) != 0) {
                    if (n < nz) { oi[n] = (int32_t)(c + j); ov[n] = (float)drow[c + j]; }
                    ++n;
                }
            }
        }
        for (; c < I; ++c) {
            if ((row[c] << 1) != 0) {
                if (n < nz) { oi[n] = (int32_t)c; ov[n] = (float)drow[c]; }
                ++n;
            }
        }
        for (int32_t j = n; j < nz; ++j) { oi[j] = 0; ov[j] = 0.f; }
        if (row_nnz) row_nnz[r] = n;
        if (n > worst) worst = n;
    }
    return worst;
}

}  // namespace

extern "C" {

// dense [rows, I] float64 (host) -> idx int32 [rows, nz], val float32 [rows, nz] (host, zero padded).
// Returns the largest number of non-zeros found in a row (>= 0): if it exceeds nz the output is truncated and the
// caller must retry with a larger nz (or keep the dense layout).  Negative: error.  threads <= 0: all cores.
// group > 0: the rows form groups of `group` consecutive rows (the H history slots of a session) of which only the
// first group_len[g] are real; the padding rows behind them (zeros by construction of collate_batch, and never read by
// the model: the encoders stop at history_len) are emitted empty without being scanned.
int64_t intel_host_pack_rows(int64_t rows, int64_t I, const double* dense, int32_t nz, int32_t* idx, float* val, int32_t* row_nnz,
                             int threads, int64_t group, const int64_t* group_len) {
    if (group > 0 && (!group_len || rows % group != 0)) {
        set_error("host_pack_rows: bad group arguments");
        return -1;
    }
    if (rows < 0 || I <= 0 || nz <= 0 || !dense || !idx || !val) {
        set_error("host_pack_rows: bad argument");
        return -1;
    }
    if (rows == 0) return 0;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if ((int64_t)nt > rows) nt = (int)rows;
    std::atomic<int32_t> worst(0);
    std::atomic<int64_t> next(0);
    const int64_t chunk = 256;                  // rows per grab: dynamic balancing, 2 MB of input at I = 1071
    auto work = [&]() {
        int32_t w = 0;
        for (;;) {
            const int64_t r0 = next.fetch_add(chunk);
            if (r0 >= rows) break;
            const int64_t r1 = r0 + chunk < rows ? r0 + chunk : rows;
            const int32_t b = pack_block(r0, r1, I, dense, nz, idx, val, row_nnz, group, group_len);
            if (b > w) w = b;
        }
        int32_t cur = worst.load();
        while (w > cur && !worst.compare_exchange_weak(cur, w)) {}
    };
    std::vector<std::thread> pool;
    pool.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return (int64_t)worst.load();
}

}  // extern "C"
