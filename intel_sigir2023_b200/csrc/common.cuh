// Shared device/host helpers for the IntEL sm_100a kernels.
#pragma once
#include <stdint.h>

#ifdef INTEL_EMU
#include "cuda_emu.h"
#define DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::S().dyn_smem)
#define LAUNCH(kfn, grid, block, smem, stream, ...) \
    emu::launch(grid, block, smem, [=]() { kfn(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define DYN_SMEM(T, name)                                          \
    extern __shared__ __align__(16) unsigned char name##_raw_[];   \
    T* name = reinterpret_cast<T*>(name##_raw_)
#define LAUNCH(kfn, grid, block, smem, stream, ...)            \
    do {                                                       \
        intel::prof_before(stream);                            \
        kfn<<<grid, block, smem, stream>>>(__VA_ARGS__);       \
        intel::prof_mark(stream);                              \
    } while (0)
#endif

#include <stdio.h>
#include <string.h>

namespace intel {

// optional per-launch device timing (bench.py's live roofline): CUDA events around every kernel,
// attributed by check_launch(name, algorithmic bytes, flops).  Off by default; no-ops in the emulator.
void prof_before(cudaStream_t s);
void prof_mark(cudaStream_t s);
bool prof_detail();                              // per-shape GEMM names requested (intel_profile_enable(2))
const char* prof_intern(const char* name);
bool prof_detail();                              // per-shape GEMM names requested (intel_profile_enable(2))
const char* prof_intern(const char* name);

// ---- error plumbing: no exceptions across the C ABI, int status + thread-local message ----
enum { INTEL_OK = 0, INTEL_ERR_ARG = 1, INTEL_ERR_WORKSPACE = 2, INTEL_ERR_CUDA = 3, INTEL_ERR_UNSUPPORTED = 4 };
void set_error(const char* fmt, ...);
int check_launch(const char* what, double bytes = 0.0, double flops = 0.0);

#define INTEL_REQUIRE(cond, code, ...)        \
    do {                                      \
        if (!(cond)) {                        \
            intel::set_error(__VA_ARGS__);    \
            return code;                      \
        }                                     \
    } while (0)
#define INTEL_TRY(expr)              \
    do {                             \
        int _st = (expr);            \
        if (_st != 0) return _st;    \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// B200: 148 SMs.  Streaming grids are sized in multiples of it (capped by the work).
static const int kNumSMs = 148;
static inline unsigned stream_grid(int64_t work_items, int per_sm) {
    int64_t cap = (int64_t)kNumSMs * per_sm;
    int64_t g = work_items < cap ? work_items : cap;
    return (unsigned)(g < 1 ? 1 : g);
}

// ---- bump allocator over the caller-provided workspace ----
struct Arena {
    unsigned char* base;
    size_t cap, off;
    bool dry;  // size query: only count
    Arena(void* p, size_t bytes) : base((unsigned char*)p), cap(bytes), off(0), dry(p == nullptr) {}
    template <class T> T* take(size_t n) {
        size_t start = align_up(off, 256);
        off = start + n * sizeof(T);
        if (dry) return nullptr;
        return reinterpret_cast<T*>(base + start);
    }
    bool ok() const { return dry || off <= cap; }
};

// Opt a kernel in to > 48 KB of dynamic shared memory ONCE (the attribute call is a context-wide operation
// that costs milliseconds; doing it per launch made the forward pass host-bound).
#ifndef INTEL_EMU
void ensure_smem_impl(const void* kernel, size_t smem);
template <class K> inline void ensure_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) ensure_smem_impl(reinterpret_cast<const void*>(kernel), smem);
}
#else
template <class K> inline void ensure_smem(K, size_t) {}
#endif

// counter-based uniform 32-bit hash (splitmix64 finaliser): BPR negative draw, dropout masks
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}
struct Dropout {
    float p, inv;          // drop probability, 1 / (1 - p)
    uint32_t thr;          // keep iff hash >= thr
    uint64_t seed;         // already mixed with (stream, layer)
};
static inline Dropout make_dropout(float p, uint64_t seed, int stream, int layer) {
    Dropout d;
    d.p = p;
    d.inv = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    double t = (double)p * 4294967296.0;
    d.thr = p > 0.f ? (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t) : 0u;
    d.seed = seed * 0x100000001B3ull + (uint64_t)(stream * 64 + layer + 1) * 0xD6E8FEB86659FD93ull;
    return d;
}
// multiplier of element (row, col) of a [rows, width] activation: 0 or 1/(1-p)
__device__ __forceinline__ float dropout_scale(const Dropout& d, int64_t row, int col, int width) {
    if (d.p <= 0.f) return 1.0f;
    return hash_u32(d.seed, (uint64_t)(row * width + col)) >= d.thr ? d.inv : 0.0f;
}

// ---- warp helpers ----
// barrier over a subset of the CTA's warps (bar.sync id, nthreads): id 1..15, nthreads a multiple of 32
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
#ifdef INTEL_EMU
    emu::named_barrier(id, nthreads);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace intel
