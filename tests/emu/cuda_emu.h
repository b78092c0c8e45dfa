// TEST INFRASTRUCTURE ONLY - a tiny single-OS-thread emulator of the CUDA execution model.
//
// The build container has no GPU and a gpurun round trip costs minutes, so the kernel
// *logic* (indexing, barriers, warp collectives, the C-ABI orchestration) is first run
// here: tests/emu/build_emu.py compiles the very same .cu sources with g++ and
// -DINTEL_EMU, every CUDA thread of a block becomes a fiber, __syncthreads() and the
// *_sync warp collectives become cooperative rendezvous points.  Blocks run one after
// another.  The product package never loads the emulated library (see _lib.py): it is
// a debugger for tests/, not a CPU fallback.
//
// Fidelity limits: no memory model (a missing barrier is only caught when the fiber
// order exposes it - run with INTEL_EMU_ORDER=reverse as well), full-mask warp
// collectives only, exited threads drop out of barriers like on hardware.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define warpSize 32

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) longlong2 { long long x, y; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline double2 make_double2(double a, double b) { return {a, b}; }
static inline int2 make_int2(int a, int b) { return {a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
enum { cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemset2DAsync(void* p, size_t pitch, int v, size_t w, size_t h, cudaStream_t) {
    for (size_t r = 0; r < h; ++r) memset((char*)p + r * pitch, v, w);
    return 0;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

namespace emu {

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    uint3 tid{0, 0, 0};
    int lin = 0;       // linear thread id in the block
    int state = 0;     // 0 runnable, 1 waiting block barrier, 2 waiting warp barrier, 3 done, 16+id waiting named barrier id
};

static const size_t kStack = 192 * 1024;
extern "C" void emu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

struct Sched {
    std::vector<Fiber> fibers;
    std::vector<char*> stacks;
    void* main_sp = nullptr;
    Fiber* cur = nullptr;
    int nthreads = 0, alive = 0, block_arrived = 0;
    int named_arrived[16] = {0}, named_expect[16] = {0};   // bar.sync id, count
    std::vector<int> warp_alive, warp_arrived;
    std::vector<uint64_t> slots;           // per-thread exchange slot for collectives
    std::function<void()> body;
    unsigned char* dyn_smem = nullptr;
    size_t dyn_cap = 0;
    dim3 gridDim_, blockDim_;
    uint3 blockIdx_{0, 0, 0};
};
inline Sched& S() { static Sched s; return s; }

inline void yield_to_main() { Sched& s = S(); emu_switch(&s.cur->sp, s.main_sp); }

static void fiber_main() {
    Sched& s = S();
    s.body();
    s.cur->state = 3;
    yield_to_main();
    abort();
}

inline void release_checks() {
    Sched& s = S();
    if (s.alive > 0 && s.block_arrived == s.alive) {
        for (auto& f : s.fibers) if (f.state == 1) f.state = 0;
        s.block_arrived = 0;
    }
    for (int id = 0; id < 16; ++id) {
        if (s.named_expect[id] > 0 && s.named_arrived[id] >= s.named_expect[id]) {
            for (auto& f : s.fibers) if (f.state == 16 + id) f.state = 0;
            s.named_arrived[id] = 0;
            s.named_expect[id] = 0;
        }
    }
    for (size_t w = 0; w < s.warp_alive.size(); ++w) {
        if (s.warp_alive[w] > 0 && s.warp_arrived[w] == s.warp_alive[w]) {
            for (int l = 0; l < 32; ++l) {
                size_t i = w * 32 + l;
                if (i < s.fibers.size() && s.fibers[i].state == 2) s.fibers[i].state = 0;
            }
            s.warp_arrived[w] = 0;
        }
    }
}

inline void run_block() {
    Sched& s = S();
    const int n = s.nthreads;
    if ((int)s.stacks.size() < n) {
        size_t old = s.stacks.size();
        s.stacks.resize(n);
        for (int i = (int)old; i < n; ++i) s.stacks[i] = (char*)aligned_alloc(64, kStack);
    }
    s.fibers.assign(n, Fiber());
    s.slots.assign(n, 0);
    int nw = (n + 31) / 32;
    s.warp_alive.assign(nw, 0);
    s.warp_arrived.assign(nw, 0);
    for (int i = 0; i < n; ++i) {
        Fiber& f = s.fibers[i];
        f.stack = s.stacks[i];
        f.lin = i;
        f.tid.x = i % s.blockDim_.x;
        f.tid.y = (i / s.blockDim_.x) % s.blockDim_.y;
        f.tid.z = i / (s.blockDim_.x * s.blockDim_.y);
        // initial frame: 6 callee-saved registers + return address into fiber_main
        uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)63;
        void** sp = (void**)top;
        *--sp = nullptr;                   // alignment pad: rsp % 16 == 8 at fiber_main entry
        *--sp = (void*)&fiber_main;
        for (int r = 0; r < 6; ++r) *--sp = nullptr;
        f.sp = sp;
        s.warp_alive[i / 32]++;
    }
    s.alive = n;
    s.block_arrived = 0;
    for (int id = 0; id < 16; ++id) s.named_arrived[id] = s.named_expect[id] = 0;
    static int order = -1;
    if (order < 0) {
        const char* e = getenv("INTEL_EMU_ORDER");
        order = (e && !strcmp(e, "reverse")) ? 1 : 0;
    }
    while (s.alive > 0) {
        bool progressed = false;
        for (int k = 0; k < n; ++k) {
            int i = order ? n - 1 - k : k;
            Fiber& f = s.fibers[i];
            if (f.state != 0) continue;
            s.cur = &f;
            emu_switch(&s.main_sp, f.sp);
            progressed = true;
            if (f.state == 3) { s.alive--; s.warp_alive[i / 32]--; }
            release_checks();
        }
        if (!progressed) { fprintf(stderr, "emu: deadlock (divergent barrier?)\n"); abort(); }
    }
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F&& f) {
    Sched& s = S();
    if (s.cur && s.cur->state != 3 && s.alive > 0) { fprintf(stderr, "emu: nested launch\n"); abort(); }
    s.body = std::forward<F>(f);
    s.gridDim_ = grid;
    s.blockDim_ = block;
    s.nthreads = block.x * block.y * block.z;
    if (smem + 64 > s.dyn_cap) { free(s.dyn_smem); s.dyn_cap = smem + 4096; s.dyn_smem = (unsigned char*)aligned_alloc(128, (s.dyn_cap + 127) / 128 * 128); }
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                s.blockIdx_ = {x, y, z};
                memset(s.dyn_smem, 0xCD, smem);   // poison: catches reads of unwritten shared memory
                run_block();
            }
    s.cur = nullptr;
}

inline void block_barrier() {
    Sched& s = S();
    s.cur->state = 1;
    s.block_arrived++;
    yield_to_main();
}
// bar.sync id, count: the first `count` threads to arrive at barrier `id` are released together
inline void named_barrier(int id, int count) {
    Sched& s = S();
    if (id < 0 || id >= 16 || (count & 31)) { fprintf(stderr, "emu: bad named barrier %d/%d\n", id, count); abort(); }
    if (s.named_expect[id] != 0 && s.named_expect[id] != count) { fprintf(stderr, "emu: named barrier %d count mismatch\n", id); abort(); }
    s.named_expect[id] = count;
    s.cur->state = 16 + id;
    s.named_arrived[id]++;
    yield_to_main();
}
inline void warp_barrier() {
    Sched& s = S();
    s.cur->state = 2;
    s.warp_arrived[s.cur->lin / 32]++;
    yield_to_main();
}
inline int lane() { return S().cur->lin & 31; }
inline int warp_base() { return S().cur->lin & ~31; }

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <class T> inline T exchange(T v, int src_lane) {
    Sched& s = S();
    s.slots[s.cur->lin] = to_bits(v);
    warp_barrier();
    int src = warp_base() + (src_lane & 31);
    T r = v;
    if (src < s.nthreads && s.fibers[src].state != 3) r = from_bits<T>(s.slots[src]);
    warp_barrier();
    return r;
}
}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().blockIdx_)
#define blockDim (emu::S().blockDim_)
#define gridDim (emu::S().gridDim_)

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline unsigned __activemask() { return 0xffffffffu; }

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::exchange(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu::exchange(v, emu::lane() ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    int src = emu::lane() + (int)d;
    T r = emu::exchange(v, src > 31 ? emu::lane() : src);
    return src > 31 ? v : r;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    int src = emu::lane() - (int)d;
    T r = emu::exchange(v, src < 0 ? emu::lane() : src);
    return src < 0 ? v : r;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Sched& s = emu::S();
    s.slots[s.cur->lin] = pred ? 1 : 0;
    emu::warp_barrier();
    unsigned r = 0;
    int base = emu::warp_base();
    for (int l = 0; l < 32; ++l) {
        int i = base + l;
        if (i < s.nthreads && s.fibers[i].state != 3 && s.slots[i]) r |= 1u << l;
    }
    emu::warp_barrier();
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) {
    emu::Sched& s = emu::S();
    unsigned b = __ballot_sync(m, p);
    unsigned live = 0;
    int base = emu::warp_base();
    for (int l = 0; l < 32; ++l) { int i = base + l; if (i < s.nthreads && s.fibers[i].state != 3) live |= 1u << l; }
    return (b & live) == live;
}

// ---- atomics (single OS thread: plain read-modify-write) ----
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { auto o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }

// ---- intrinsics ----
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __saturatef(float x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }
static inline int __float_as_int(float f) { return emu::from_bits<int>(emu::to_bits(f)); }
static inline unsigned __float_as_uint(float f) { return emu::from_bits<unsigned>(emu::to_bits(f)); }
static inline float __int_as_float(int i) { return emu::from_bits<float>(emu::to_bits(i)); }
static inline float __uint_as_float(unsigned i) { return emu::from_bits<float>(emu::to_bits(i)); }
static inline long long __double_as_longlong(double d) { return emu::from_bits<long long>(emu::to_bits(d)); }
static inline double __longlong_as_double(long long i) { return emu::from_bits<double>(emu::to_bits(i)); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (x & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
using std::max;
using std::min;
