// Hardware probe for the tcgen05 forms the fused stack kernels rely on (run on a B200: tests/hw/run_probe.sh).
// Every test multiplies small integer matrices (exact in TF32 and in the fp32 accumulator), so a correct form
// reproduces the CPU result bit for bit.  One test per process: an illegal form must not poison the others.
//   1  SS, both operands K-major, quad-slab planes [k-chunk][row][4]         (LBO = rows * 16, SBO = 128)
//   2  TS: A read from tensor memory (tcgen05.st 32x32b), B K-major
//   3  SS with a disable-output-lane mask (two MMAs fill the two halves of one accumulator)
//   4  TS with a disable-output-lane mask
//   5  SS, both operands MN-major from the same quad-slab planes (contraction over the plane rows), M = 64
//   6  TS + MN-major B (the input-gradient form dX = dY W)
//   7  SS MN-major A with M = 128, K-major B
//   8  cycles of one issue -> commit -> wait round trip (M = 128, N = 64, 12 MMAs)
//  14  K-major operands read from the SAME natural tiles (32-byte chunk XOR row & 3) that tests 9-11 read MN-major:
//      SWIZZLE_128B_BASE32B as a K-major layout (one tile = both views: S = Q K^T and dQ = dS K from one copy of K)
//  15  SS, quad-slab planes, operands with garbage in the low 13 mantissa bits: the tensor core truncates (ignores them)
//  16  TS, A in tensor memory with garbage low bits
//  17  mma.sync.m16n8k8 tf32 with garbage low bits (one warp)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return make_desc(saddr, lbo, sbo) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ss_mask(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t m0, uint32_t m1,
                                            uint32_t m2, uint32_t m3) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts_mask(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t m0, uint32_t m1,
                                            uint32_t m2, uint32_t m3) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n}" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
                 : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// P[rows][cols] (row-major, global) -> quad-slab plane [cols / 4][rows][4] in shared memory
__device__ void stage_plane(float* sm, const float* P, int rows, int cols) {
    for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
        const int r = e / cols, c = e % cols;
        sm[(c / 4) * rows * 4 + r * 4 + (c % 4)] = P[e];
    }
}

// P[rows][32] (row-major, global) -> natural tile with 128-byte rows; the 32-byte chunk index is XORed with (row & 3)
// (Swizzle<2,5,2>: byte-address bits [5,7) ^= bits [7,9)), the MN-major SWIZZLE_128B_BASE32B atom of 4 k-rows x 32 elements
__device__ void stage_nat(float* sm, const float* P, int ld, int c0, int rows) {
    for (int e = threadIdx.x; e < rows * 32; e += blockDim.x) {
        const int r = e / 32, c = e % 32;
        const int byte = r * 128 + c * 4;
        const int sw = byte ^ (((byte >> 7) & 3) << 5);
        sm[sw / 4] = P[r * ld + c0 + c];
    }
}

// natural tile, 128-byte rows, 16-byte chunk index XORed with (row & 7): the SWIZZLE_128B atom (8 rows x 128 bytes)
__device__ void stage_nat128(float* sm, const float* P, int ld, int c0, int rows) {
    for (int e = threadIdx.x; e < rows * 32; e += blockDim.x) {
        const int r = e / 32, c = e % 32;
        const int byte = r * 128 + c * 4;
        const int sw = byte ^ (((byte >> 7) & 7) << 4);
        sm[sw / 4] = P[r * ld + c0 + c];
    }
}

__device__ __forceinline__ float dirty(float x, int i) {      // same tf32 value, random low 13 bits
    return __uint_as_float(__float_as_uint(x) | ((uint32_t)(i * 2654435761u) >> 19));
}
__device__ void stage_plane_dirty(float* sm, const float* P, int rows, int cols) {
    for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
        const int r = e / cols, c = e % cols;
        sm[(c / 4) * rows * 4 + r * 4 + (c % 4)] = dirty(P[e], e + 7);
    }
}

struct Args {
    int test;
    const float *A, *B, *B2;    // row-major inputs
    float* D;                   // [128][64] dump of the accumulator lanes x columns
    long long* cycles;
};

__global__ void __launch_bounds__(128) probe(Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    float* sA = reinterpret_cast<float*>(smem);                 // up to 128 x 64 floats = 32 KB
    float* sB = sA + 128 * 64;                                   // 16 KB
    float* sB2 = sB + 128 * 32;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base_s;
    const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
    // sentinel in the accumulator columns 0..63
    {
        uint32_t s[32];
        for (int j = 0; j < 32; ++j) s[j] = __float_as_uint(-777.0f);
        tmem_st32(lane_addr + 0, s);
        tmem_st32(lane_addr + 32, s);
    }
    const int T = a.test;
    long long t0 = 0, t1 = 0;
    if (T == 1 || T == 3) {                                      // A [128][32], B [32][32], B2 [32][32]
        stage_plane(sA, a.A, 128, 32);
        stage_plane(sB, a.B, 32, 32);
        stage_plane(sB2, a.B2, 32, 32);
    } else if (T == 2 || T == 4 || T == 6) {                     // A -> tensor memory columns 256..287
        uint32_t r[32];
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(a.A[tid * 32 + j]);
        tmem_st32(lane_addr + 256, r);
        stage_plane(sB, a.B, 32, 32);
        stage_plane(sB2, a.B2, 32, 32);
    } else if (T == 5) {                                         // A = G [128 tok][64 ch], B = X [128 tok][32 ch]
        stage_plane(sA, a.A, 128, 64);
        stage_plane(sB, a.B, 128, 32);
    } else if (T == 7) {                                         // A = G [32 k][128 m] quad-slab over m, B [32 n][32 k]
        stage_plane(sA, a.A, 32, 128);
        stage_plane(sB, a.B, 32, 32);
    } else if (T == 8) {
        stage_plane(sA, a.A, 128, 32);
        stage_plane(sB, a.B, 64, 32);
    } else if (T == 9) {                                         // G [128 tok][64 ch] as two natural tiles, X [128 tok][32 ch]
        stage_nat(sA, a.A, 64, 0, 128);
        stage_nat(sA + 128 * 32, a.A, 64, 32, 128);
        stage_nat(sB, a.B, 32, 0, 128);
    } else if (T == 10) {                                        // dY -> tensor memory, W [32 out][32 in] natural tile
        uint32_t r[32];
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(a.A[tid * 32 + j]);
        tmem_st32(lane_addr + 256, r);
        stage_nat(sB, a.B, 32, 0, 32);
    } else if (T == 12) {                                        // A [128][32], B [32][32]: natural tiles, SWIZZLE_128B, K-major
        stage_nat128(sA, a.A, 32, 0, 128);
        stage_nat128(sB, a.B, 32, 0, 32);
    } else if (T == 13) {                                        // wgrad form on SWIZZLE_128B natural tiles, MN-major
        stage_nat128(sA, a.A, 64, 0, 128);
        stage_nat128(sA + 128 * 32, a.A, 64, 32, 128);
        stage_nat128(sB, a.B, 32, 0, 128);
    } else if (T == 14) {                                        // A [128][32], B [32][32]: the MN-major natural tiles, read K-major
        stage_nat(sA, a.A, 32, 0, 128);
        stage_nat(sB, a.B, 32, 0, 32);
    } else if (T == 15) {
        stage_plane_dirty(sA, a.A, 128, 32);
        stage_plane_dirty(sB, a.B, 32, 32);
    } else if (T == 16) {
        uint32_t r[32];
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(dirty(a.A[tid * 32 + j], tid * 32 + j + 3));
        tmem_st32(lane_addr + 256, r);
        stage_plane_dirty(sB, a.B, 32, 32);
    } else if (T == 11) {                                        // G [32 k][128 m] as four natural tiles, B [32 n][32 k] K-major
        for (int q = 0; q < 4; ++q) stage_nat(sA + q * 32 * 32, a.A, 128, 32 * q, 32);
        stage_plane(sB, a.B, 32, 32);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aA = smem_u32(sA), aB = smem_u32(sB), aB2 = smem_u32(sB2);
        if (T == 1) {
            const uint32_t id = make_idesc(128, 32, 0, 0);
            for (int ks = 0; ks < 4; ++ks)
                mma_ss(tm, make_desc(aA + ks * 2 * 2048, 2048, 128), make_desc(aB + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u);
        } else if (T == 2) {
            const uint32_t id = make_idesc(128, 32, 0, 0);
            for (int ks = 0; ks < 4; ++ks) mma_ts(tm, tm + 256 + ks * 8, make_desc(aB + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u);
        } else if (T == 3) {
            const uint32_t id = make_idesc(128, 32, 0, 0);
            for (int ks = 0; ks < 4; ++ks)
                mma_ss_mask(tm, make_desc(aA + ks * 2 * 2048, 2048, 128), make_desc(aB + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u, 0u, 0u,
                            0xffffffffu, 0xffffffffu);
            for (int ks = 0; ks < 4; ++ks)
                mma_ss_mask(tm, make_desc(aA + ks * 2 * 2048, 2048, 128), make_desc(aB2 + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u,
                            0xffffffffu, 0xffffffffu, 0u, 0u);
        } else if (T == 4) {
            const uint32_t id = make_idesc(128, 32, 0, 0);
            for (int ks = 0; ks < 4; ++ks)
                mma_ts_mask(tm, tm + 256 + ks * 8, make_desc(aB + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u, 0u, 0u, 0xffffffffu, 0xffffffffu);
            for (int ks = 0; ks < 4; ++ks)
                mma_ts_mask(tm, tm + 256 + ks * 8, make_desc(aB2 + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u, 0xffffffffu, 0xffffffffu, 0u, 0u);
        } else if (T == 5) {
            // D[m = G channel (64)][n = X channel (32)] = sum over the 128 tokens; plane [quad][128 tok][4]: quad stride 2048 B,
            // 8-token group stride 128 B
            const uint32_t id = make_idesc(64, 32, 1, 1);
            for (int ks = 0; ks < 16; ++ks)
                mma_ss(tm, make_desc(aA + ks * 128, 128, 2048), make_desc(aB + ks * 128, 128, 2048), id, ks ? 1u : 0u);
        } else if (T == 6) {
            // dX[tok][in] = sum_out dY[tok][out] W[out][in]: W plane [in quad][32 out][4] read MN-major (n = in, k = out)
            const uint32_t id = make_idesc(128, 32, 0, 1);
            for (int ks = 0; ks < 4; ++ks) mma_ts(tm, tm + 256 + ks * 8, make_desc(aB + ks * 128, 128, 512), id, ks ? 1u : 0u);
        } else if (T == 7) {
            // D[m (128)][n (32)] = sum_k G[k][m] B[n][k]: A plane [m quad (32)][32 k][4] MN-major, B K-major
            const uint32_t id = make_idesc(128, 32, 1, 0);
            for (int ks = 0; ks < 4; ++ks)
                mma_ss(tm, make_desc(aA + ks * 128, 128, 512), make_desc(aB + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u);
        } else if (T == 9) {
            // D[m = G channel (64)][n = X channel (32)] over 128 tokens: 32-channel groups LBO = 16 KB apart, 4-token groups SBO = 512 B
            const uint32_t id = make_idesc(64, 32, 1, 1);
            for (int ks = 0; ks < 16; ++ks)
                mma_ss(tm, make_desc_sw(aA + ks * 1024, 16384, 512, 1), make_desc_sw(aB + ks * 1024, 16384, 512, 1), id, ks ? 1u : 0u);
        } else if (T == 14) {
            const uint32_t id = make_idesc(128, 32, 0, 0);
            for (int ks = 0; ks < 4; ++ks)
                mma_ss(tm, make_desc_sw(aA + ks * 32, 16, 1024, 1), make_desc_sw(aB + ks * 32, 16, 1024, 1), id, ks ? 1u : 0u);
        } else if (T == 15) {
            const uint32_t id = make_idesc(128, 32, 0, 0);
            for (int ks = 0; ks < 4; ++ks)
                mma_ss(tm, make_desc(aA + ks * 2 * 2048, 2048, 128), make_desc(aB + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u);
        } else if (T == 16) {
            const uint32_t id = make_idesc(128, 32, 0, 0);
            for (int ks = 0; ks < 4; ++ks) mma_ts(tm, tm + 256 + ks * 8, make_desc(aB + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u);
        } else if (T == 12) {
            const uint32_t id = make_idesc(128, 32, 0, 0);
            for (int ks = 0; ks < 4; ++ks)
                mma_ss(tm, make_desc_sw(aA + ks * 32, 16, 1024, 2), make_desc_sw(aB + ks * 32, 16, 1024, 2), id, ks ? 1u : 0u);
        } else if (T == 13) {
            const uint32_t id = make_idesc(64, 32, 1, 1);
            for (int ks = 0; ks < 16; ++ks)
                mma_ss(tm, make_desc_sw(aA + ks * 1024, 16384, 1024, 2), make_desc_sw(aB + ks * 1024, 16384, 1024, 2), id, ks ? 1u : 0u);
        } else if (T == 10) {
            const uint32_t id = make_idesc(128, 32, 0, 1);
            for (int ks = 0; ks < 4; ++ks) mma_ts(tm, tm + 256 + ks * 8, make_desc_sw(aB + ks * 1024, 4096, 512, 1), id, ks ? 1u : 0u);
        } else if (T == 11) {
            const uint32_t id = make_idesc(128, 32, 1, 0);
            for (int ks = 0; ks < 4; ++ks)
                mma_ss(tm, make_desc_sw(aA + ks * 1024, 4096, 512, 1), make_desc(aB + ks * 2 * 512, 512, 128), id, ks ? 1u : 0u);
        } else if (T == 8) {
            const uint32_t id = make_idesc(128, 64, 0, 0);
            t0 = clock64();
            for (int rep = 0; rep < 3; ++rep)
                for (int ks = 0; ks < 4; ++ks)
                    mma_ss(tm, make_desc(aA + ks * 2 * 2048, 2048, 128), make_desc(aB + ks * 2 * 1024, 1024, 128), id, (rep | ks) ? 1u : 0u);
        }
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0 && T == 8) {
        t1 = clock64();
        a.cycles[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        uint32_t r[32];
        tmem_ld32(lane_addr + 0, r);
        for (int j = 0; j < 32; ++j) a.D[tid * 64 + j] = __uint_as_float(r[j]);
        tmem_ld32(lane_addr + 32, r);
        for (int j = 0; j < 32; ++j) a.D[tid * 64 + 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// test 17: D[16][8] = A[16][8] B[8][8]^T on one warp with mma.sync tf32, operands carrying garbage low bits
__global__ void probe_mma_sync(const float* A, const float* B, float* D) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    uint32_t a[4], b[2];
    a[0] = __float_as_uint(dirty(A[g * 32 + t], lane));
    a[1] = __float_as_uint(dirty(A[(g + 8) * 32 + t], lane + 40));
    a[2] = __float_as_uint(dirty(A[g * 32 + t + 4], lane + 80));
    a[3] = __float_as_uint(dirty(A[(g + 8) * 32 + t + 4], lane + 120));
    b[0] = __float_as_uint(dirty(B[g * 32 + t], lane + 160));
    b[1] = __float_as_uint(dirty(B[g * 32 + t + 4], lane + 200));
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    D[g * 8 + 2 * t] = d[0]; D[g * 8 + 2 * t + 1] = d[1]; D[(g + 8) * 8 + 2 * t] = d[2]; D[(g + 8) * 8 + 2 * t + 1] = d[3];
}

static std::vector<float> ints(int n, unsigned seed) {
    std::vector<float> v(n);
    unsigned s = seed * 2654435761u + 12345u;
    for (int i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        v[i] = (float)((int)((s >> 16) % 9) - 4);
    }
    return v;
}

int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 1;
    std::vector<float> A = ints(128 * 64, 1), B = ints(128 * 32, 2), B2 = ints(32 * 32, 3), D(128 * 64, 0.f);
    float *dA, *dB, *dB2, *dD;
    long long* dC;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dB2, B2.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB2, B2.data(), B2.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    const int smem = 128 * 64 * 4 + 128 * 32 * 4 + 32 * 32 * 4;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (T == 17) {
        probe_mma_sync<<<1, 32>>>(dA, dB, dD);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, 16 * 8 * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 16; ++m)
            for (int n = 0; n < 8; ++n) {
                float e = 0;
                for (int k = 0; k < 8; ++k) e += A[m * 32 + k] * B[n * 32 + k];
                if (D[m * 8 + n] != e) { if (bad < 8) printf("  mismatch %d %d: got %.9g want %g\n", m, n, D[m * 8 + n], e); ++bad; }
            }
        printf("test 17: %s (%d mismatches)\n", bad ? "FAIL" : "PASS", bad);
        return bad ? 1 : 0;
    }
    Args a{T, dA, dB, dB2, dD, dC};
    probe<<<1, 128, smem>>>(a);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
    // expected accumulator E[lane][col]
    std::vector<float> E(128 * 64, -777.0f);
    auto dot = [&](const float* x, int sx, const float* y, int sy, int n) { float s = 0; for (int k = 0; k < n; ++k) s += x[k * sx] * y[k * sy]; return s; };
    if (T == 1 || T == 2 || T == 12 || T == 14 || T == 15 || T == 16) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) E[m * 64 + n] = dot(&A[m * 32], 1, &B[n * 32], 1, 32);
    } else if (T == 3 || T == 4) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) E[m * 64 + n] = dot(&A[m * 32], 1, m < 64 ? &B[n * 32] : &B2[n * 32], 1, 32);
    } else if (T == 10) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) E[m * 64 + n] = dot(&A[m * 32], 1, &B[n], 32, 32);
    } else if (T == 11) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) E[m * 64 + n] = dot(&A[m], 128, &B[n * 32], 1, 32);
    } else if (T == 5 || T == 9 || T == 13) {
        for (int m = 0; m < 64; ++m) {
            const int lane = (m % 16) + 32 * (m / 16);
            for (int n = 0; n < 32; ++n) E[lane * 64 + n] = dot(&A[m], 64, &B[n], 32, 128);
        }
    } else if (T == 6) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) E[m * 64 + n] = dot(&A[m * 32], 1, &B[n], 32, 32);
    } else if (T == 7) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) E[m * 64 + n] = dot(&A[m], 128, &B[n * 32], 1, 32);
    }
    if (T == 8) { printf("test 8: round trip of 12 MMAs (128x64x8) = %lld cycles\n", cyc); return 0; }
    int bad = 0;
    for (int i = 0; i < 128 * 64; ++i) {
        if ((i % 64) >= 32 && T != 99) { if (D[i] != -777.0f) { if (bad < 5) printf("  col>=32 touched: lane %d col %d = %g\n", i / 64, i % 64, D[i]); ++bad; } continue; }
        if (D[i] != E[i]) { if (bad < 8) printf("  mismatch lane %d col %d: got %g want %g\n", i / 64, i % 64, D[i], E[i]); ++bad; }
    }
    printf("test %d: %s (%d mismatches)\n", T, bad ? "FAIL" : "PASS", bad);
    if (bad && (T == 5 || T == 9 || T == 13)) {    // where did the rows land?
        for (int lane = 0; lane < 128; lane += 1) {
            int hit = -1;
            for (int m = 0; m < 64 && hit < 0; ++m) {
                bool ok = true;
                for (int n = 0; n < 32 && ok; ++n) ok = D[lane * 64 + n] == dot(&A[m], 64, &B[n], 32, 128);
                if (ok) hit = m;
            }
            if (hit >= 0) printf("  lane %d holds row %d\n", lane, hit);
        }
    }
    return bad ? 1 : 0;
}
