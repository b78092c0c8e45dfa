// Warp-level tensor-core primitive for the fp32-parity contractions: m16n8k8 TF32 MMA with fp32
// accumulation, used three times per product ("3xTF32": a = a_hi + a_lo, b = b_hi + b_lo,
// d += a_lo*b_hi + a_hi*b_lo + a_hi*b_hi) so the result carries ~21 mantissa bits - inside the
// 1e-5 parity budget that plain TF32/BF16 tensor-core math would break.
//
// Fragment layout (PTX ISA, mma.m16n8k8 .tf32): g = lane >> 2, t = lane & 3
//   A (16x8, row):  a0=(g, t)  a1=(g+8, t)  a2=(g, t+4)  a3=(g+8, t+4)
//   B (8x8,  col):  b0=(k=t, n=g)  b1=(k=t+4, n=g)
//   C/D (16x8):     c0=(g, 2t)  c1=(g, 2t+1)  c2=(g+8, 2t)  c3=(g+8, 2t+1)
#pragma once
#include "common.cuh"

namespace intel {

#ifndef INTEL_EMU
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#else
// emulator: round-to-nearest (ties away) to a 10-bit mantissa, and a cooperative MMA through a per-warp
// scratch so the fragment index algebra of the kernels is exercised on the CPU
static inline uint32_t to_tf32(float x) {
    uint32_t u = emu::from_bits<uint32_t>(emu::to_bits(x));
    if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;
    return u & 0xffffe000u;
}
static inline void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    static float sa[64][32][4], sb[64][32][2];
    const int lin = emu::S().cur->lin, w = lin >> 5, lane = lin & 31;
    // the tensor core reads the upper 19 bits of an operand register (tests/hw/umma_probe.cu, tests 15-17)
    for (int i = 0; i < 4; ++i) sa[w][lane][i] = emu::from_bits<float>((uint64_t)(a[i] & 0xffffe000u));
    for (int i = 0; i < 2; ++i) sb[w][lane][i] = emu::from_bits<float>((uint64_t)(b[i] & 0xffffe000u));
    emu::warp_barrier();
    const int g = lane >> 2, t = lane & 3;
    auto A = [&](int r, int k) { return sa[w][(r & 7) * 4 + (k & 3)][(r >> 3) + 2 * (k >> 2)]; };
    auto B = [&](int k, int n) { return sb[w][n * 4 + (k & 3)][k >> 2]; };
    const int rows[4] = {g, g, g + 8, g + 8}, cols[4] = {2 * t, 2 * t + 1, 2 * t, 2 * t + 1};
    for (int i = 0; i < 4; ++i) {
        float acc = d[i];
        for (int k = 0; k < 8; ++k) acc += A(rows[i], k) * B(k, cols[i]);
        d[i] = acc;
    }
    emu::warp_barrier();
}
#endif

// Operand split for 3xTF32.  cvt.rna.tf32.f32 expands to a ~5-instruction sequence on sm_100a and made the
// MMA kernels issue-bound (profiles/r01_summary.md), so the split is done with integer ops on the bit pattern:
//   hi = round-to-nearest of x to 10 mantissa bits  ((bits + 0x1000) & ~0x1fff; operands are finite)
//   lo = x - hi (exact in fp32); its low 13 bits are left in place: the tensor core reads only the upper 19 bits of an
//        operand (mma.sync and tcgen05, shared-memory and tensor-memory operands alike: tests/hw/umma_probe.cu, tests 15-17),
//        so clearing them would be a wasted instruction in the hottest sequence of every 3xTF32 kernel.
// |x - hi - trunc(lo)| <= 2^-21 |x|.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// d += a * b at ~fp32 accuracy from three TF32 MMAs; af / bf are fp32 fragment values
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const float (&af)[4], const float (&bf)[2]) {
    uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(af[i], ah[i], al[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) split_tf32(bf[i], bh[i], bl[i]);
    mma_tf32(d, al, bh);
    mma_tf32(d, ah, bl);
    mma_tf32(d, ah, bh);
}

}  // namespace intel
