// Multi-head self-attention without output projection (modules/layers.py:31-60), forward and backward,
// one CTA per (session, head).  A whole head of a session - Q, K, V (and dO) tiles plus one 64-row block of
// the score matrix - lives in shared memory, so the reference's [B,h,L,L] tensors never reach HBM.  Every
// product (Q K^T, P V, dO V^T, dS K, dS^T Q, P^T dO) is issued as warp-level 3xTF32 tensor-core MMAs over
// the shared-memory tiles (mma.cuh); softmax and the dS algebra run one warp per row.
//
// The reference shifts the softmax by the *global* max of the score tensor (layers.py:57), a pure numerics
// choice; the row max is used here (softmax is shift invariant).  Keys j >= lens[b] are masked when `lens`
// is given (BERT4Rec, GeneralSeq.py:100-101); the IntEL stacks pass none, so pad slots are live keys and
// queries exactly as in the reference.
#include "kernels.h"
#include "mma.cuh"

namespace intel {

static const int MHA_WARPS = 4;
static const int QB = 16 * MHA_WARPS;     // query rows per block: one m-tile per warp

// One warp: C[16 x 8*n_tiles] = A[16 x 8*k_steps] * B, operands addressed through strides, results handed
// to epi(row, col, v_col, v_col+1) per accumulator pair.  Four n-tiles are held in registers per pass.
template <class Epi>
__device__ __forceinline__ void warp_mma(const float* __restrict__ A, int a_rs, int a_cs, int a_row0,
                                         const float* __restrict__ B, int b_ks, int b_ns, int n_tiles, int k_steps,
                                         int lane, Epi epi) {
    const int gq = lane >> 2, tq = lane & 3;
    for (int nt0 = 0; nt0 < n_tiles; nt0 += 4) {
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
        for (int ks = 0; ks < k_steps; ++ks) {
            const int k0 = ks * 8 + tq;
            float af[4];
            af[0] = A[(a_row0 + gq) * a_rs + k0 * a_cs];
            af[1] = A[(a_row0 + gq + 8) * a_rs + k0 * a_cs];
            af[2] = A[(a_row0 + gq) * a_rs + (k0 + 4) * a_cs];
            af[3] = A[(a_row0 + gq + 8) * a_rs + (k0 + 4) * a_cs];
            uint32_t ah[4], al[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                split_tf32(af[c], ah[c], al[c]);
            }
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // tiles past n_tiles read column 0 of B (always mapped) and are never stored
                const int n = (nt0 + j < n_tiles) ? (nt0 + j) * 8 + gq : gq;
                const float b0 = B[k0 * b_ks + n * b_ns], b1 = B[(k0 + 4) * b_ks + n * b_ns];
                split_tf32(b0, bh[j][0], bl[j][0]);
                split_tf32(b1, bh[j][1], bl[j][1]);
            }
            // three independent passes over the four accumulators: no back-to-back dependent MMAs
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_tf32(acc[j], al, bh[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_tf32(acc[j], ah, bl[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_tf32(acc[j], ah, bh[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (nt0 + j < n_tiles) {
                const int col = (nt0 + j) * 8 + 2 * tq;
                epi(a_row0 + gq, col, acc[j][0], acc[j][1]);
                epi(a_row0 + gq + 8, col, acc[j][2], acc[j][3]);
            }
        }
    }
}

template <int DK>
struct MhaDims {
    static constexpr int dk = DK, st = DK + 4;   // head width and tile stride (conflict-free fragment reads)
    int T, Tp, d, heads, ss;                     // Tp = T rounded up to 16; ss = score-block stride
};

template <int DK>
__device__ __forceinline__ MhaDims<DK> mha_dims(int64_t T, int d, int heads) {
    MhaDims<DK> m;
    m.T = (int)T; m.Tp = ((int)T + 15) / 16 * 16; m.d = d; m.heads = heads;
    m.ss = m.Tp + 4;
    return m;
}

// stage rows [0, Tp) of one head's slice into a [Tp][st] tile, zero beyond `rows_valid`
template <int DK>
__device__ __forceinline__ void stage_tile(float* dst, const float* __restrict__ src, int64_t row_stride, int rows_valid,
                                           const MhaDims<DK>& m) {
    // 16-byte chunks: head slices start at multiples of DK floats and DK % 4 == 0, row strides are multiples of 4
    for (int e = threadIdx.x; e < m.Tp * (DK / 4); e += blockDim.x) {
        const int j = e / (DK / 4), c = (e % (DK / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < rows_valid) v = *reinterpret_cast<const float4*>(src + (int64_t)j * row_stride + c);
        *reinterpret_cast<float4*>(dst + j * m.st + c) = v;
    }
}

// row-wise softmax of one score block: rows [0,QB), valid queries i0+r < T, keys j < nk; masked entries -> 0
template <int DK>
__device__ __forceinline__ void softmax_block(float* S, const MhaDims<DK>& m, int i0, int nk, int lane, int warp,
                                              float* row_inv /* nullable: 1/sum per row */) {
    for (int r = warp; r < QB; r += MHA_WARPS) {
        float* s = S + r * m.ss;
        if (i0 + r >= m.T || nk <= 0) {
            for (int j = lane; j < m.Tp; j += 32) s[j] = 0.f;
            continue;
        }
        float mx = -INFINITY;
        for (int j = lane; j < nk; j += 32) mx = fmaxf(mx, s[j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < nk; j += 32) {
            const float e = expf(s[j] - mx);
            s[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < m.Tp; j += 32) s[j] = (j < nk) ? s[j] * inv : 0.f;
        if (row_inv && lane == 0) row_inv[r] = inv;
    }
}

template <int DK>
__global__ void __launch_bounds__(MHA_WARPS * 32) mha_fwd_kernel(int64_t T, int d, int heads,
                                                                 const float* __restrict__ QKV,
                                                                 const int64_t* __restrict__ lens,
                                                                 float* __restrict__ O) {
    DYN_SMEM(float, sm);
    const MhaDims<DK> m = mha_dims<DK>(T, d, heads);
    const int64_t b = blockIdx.x / heads;
    const int h = blockIdx.x % heads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int nk = lens ? (int)(lens[b] < T ? lens[b] : T) : m.T;
    float* Ks = sm;
    float* Vs = Ks + m.Tp * m.st;
    float* Qs = Vs + m.Tp * m.st;          // [QB][st]
    float* S = Qs + QB * m.st;             // [QB][ss]
    const float* base = QKV + b * T * 3 * d + h * m.dk;
    stage_tile(Ks, base + d, 3 * d, m.T, m);
    stage_tile(Vs, base + 2 * d, 3 * d, m.T, m);
    const float scale = 1.0f / sqrtf((float)m.dk);
    for (int i0 = 0; i0 < m.T; i0 += QB) {
        __syncthreads();
        for (int e = threadIdx.x; e < QB * (DK / 4); e += blockDim.x) {
            const int r = e / (DK / 4), c = (e % (DK / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i0 + r < m.T) v = *reinterpret_cast<const float4*>(base + (int64_t)(i0 + r) * 3 * d + c);
            *reinterpret_cast<float4*>(Qs + r * m.st + c) = v;
        }
        __syncthreads();
        // S = scale * Q K^T   (B(k=c, n=j) = K[j][c])
        warp_mma(Qs, m.st, 1, warp * 16, Ks, 1, m.st, m.Tp / 8, m.dk / 8, lane,
                 [&](int r, int c, float v0, float v1) { S[r * m.ss + c] = v0 * scale; S[r * m.ss + c + 1] = v1 * scale; });
        __syncthreads();
        softmax_block(S, m, i0, nk, lane, warp, nullptr);
        __syncthreads();
        // O = P V   (B(k=j, n=c) = V[j][c])
        float* Ob = O + (b * T + i0) * d + h * m.dk;
        warp_mma(S, m.ss, 1, warp * 16, Vs, m.st, 1, m.dk / 8, m.Tp / 8, lane,
                 [&](int r, int c, float v0, float v1) {
                     if (i0 + r < m.T) { Ob[(int64_t)r * d + c] = v0; Ob[(int64_t)r * d + c + 1] = v1; }
                 });
    }
}

static size_t mha_fwd_smem(int64_t T, int dk) {
    const int64_t Tp = (T + 15) / 16 * 16;
    return (size_t)(2 * Tp * (dk + 4) + QB * (dk + 4) + QB * (Tp + 4)) * 4;
}
static size_t mha_bwd_smem(int64_t T, int dk) {
    const int64_t Tp = (T + 15) / 16 * 16;
    return (size_t)(4 * Tp * (dk + 4) + 2 * QB * (Tp + 4) + 2 * QB) * 4;
}
static const size_t kMaxSmem = 227 * 1024;

int mha_fwd(int64_t B, int64_t T, int d, int heads, const float* QKV, const int64_t* lens, float* O, cudaStream_t s) {
    if (B <= 0 || T <= 0) return INTEL_OK;
    INTEL_REQUIRE(heads > 0 && d % heads == 0 && (d / heads) % 8 == 0, INTEL_ERR_UNSUPPORTED,
                  "mha: head width d/heads = %d/%d must be a multiple of 8", d, heads);
    const size_t smem = mha_fwd_smem(T, d / heads);
    INTEL_REQUIRE(smem <= kMaxSmem, INTEL_ERR_UNSUPPORTED, "mha_fwd: list length %lld too long for one SM", (long long)T);
    INTEL_REQUIRE(d % 4 == 0 && ((uintptr_t)QKV % 16 == 0), INTEL_ERR_ARG, "mha: qkv must be 16-byte aligned");
    const dim3 grid((unsigned)(B * heads)), block(MHA_WARPS * 32);
#define INTEL_MHA_FWD(DKV)                                                                                         \
    case DKV: {                                                                                                    \
        auto k = mha_fwd_kernel<DKV>;                                                                              \
        ensure_smem(k, smem);     \
        LAUNCH(k, grid, block, smem, s, T, d, heads, QKV, lens, O);                                                \
    } break;
    switch (d / heads) {
        INTEL_MHA_FWD(8) INTEL_MHA_FWD(16) INTEL_MHA_FWD(24) INTEL_MHA_FWD(32) INTEL_MHA_FWD(48) INTEL_MHA_FWD(64)
        default:
            set_error("mha: head width %d is not instantiated (8,16,24,32,48,64)", d / heads);
            return INTEL_ERR_UNSUPPORTED;
    }
#undef INTEL_MHA_FWD
    return check_launch("mha_fwd", 16.0 * B * T * d, 4.0 * B * T * T * d);
}

// Backward by recomputation, query block by query block: P and dP = dO V^T are rebuilt in shared memory,
// dS = P (dP - rowsum(P dP)) / sqrt(dk), then dQ = dS K, dK += dS^T Q, dV += P^T dO.
template <int DK>
__global__ void __launch_bounds__(MHA_WARPS * 32) mha_bwd_kernel(int64_t T, int d, int heads,
                                                                 const float* __restrict__ QKV,
                                                                 const int64_t* __restrict__ lens,
                                                                 const float* __restrict__ dO,
                                                                 float* __restrict__ dQKV) {
    DYN_SMEM(float, sm);
    const MhaDims<DK> m = mha_dims<DK>(T, d, heads);
    const int64_t b = blockIdx.x / heads;
    const int h = blockIdx.x % heads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int nk = lens ? (int)(lens[b] < T ? lens[b] : T) : m.T;
    float* Qs = sm;
    float* Ks = Qs + m.Tp * m.st;
    float* Vs = Ks + m.Tp * m.st;
    float* Gs = Vs + m.Tp * m.st;          // dO tile
    float* P = Gs + m.Tp * m.st;           // [QB][ss]
    float* D = P + QB * m.ss;              // [QB][ss]  dP, then dS
    const float* base = QKV + b * T * 3 * d + h * m.dk;
    float* dbase = dQKV + b * T * 3 * d + h * m.dk;
    stage_tile(Qs, base, 3 * d, m.T, m);
    stage_tile(Ks, base + d, 3 * d, m.T, m);
    stage_tile(Vs, base + 2 * d, 3 * d, m.T, m);
    stage_tile(Gs, dO + b * T * d + h * m.dk, d, m.T, m);
    const float scale = 1.0f / sqrtf((float)m.dk);
    for (int i0 = 0; i0 < m.T; i0 += QB) {
        __syncthreads();
        // scores and dP for this query block (rows beyond Tp of the tiles do not exist: clamp the m-tile)
        const bool live = (i0 + warp * 16) < m.Tp;
        if (live) {
            warp_mma(Qs + i0 * m.st, m.st, 1, warp * 16, Ks, 1, m.st, m.Tp / 8, m.dk / 8, lane,
                     [&](int r, int c, float v0, float v1) { P[r * m.ss + c] = v0 * scale; P[r * m.ss + c + 1] = v1 * scale; });
            warp_mma(Gs + i0 * m.st, m.st, 1, warp * 16, Vs, 1, m.st, m.Tp / 8, m.dk / 8, lane,
                     [&](int r, int c, float v0, float v1) { D[r * m.ss + c] = v0; D[r * m.ss + c + 1] = v1; });
        } else {
            for (int e = lane; e < 16 * m.ss; e += 32) { P[warp * 16 * m.ss + e] = 0.f; D[warp * 16 * m.ss + e] = 0.f; }
        }
        __syncthreads();
        softmax_block(P, m, i0, nk, lane, warp, nullptr);
        __syncthreads();
        for (int r = warp; r < QB; r += MHA_WARPS) {
            float* p = P + r * m.ss;
            float* g = D + r * m.ss;
            float delta = 0.f;
            for (int j = lane; j < m.Tp; j += 32) delta = fmaf(p[j], g[j], delta);
            delta = warp_sum(delta);
            for (int j = lane; j < m.Tp; j += 32) g[j] = p[j] * (g[j] - delta) * scale;
        }
        __syncthreads();
        // dQ = dS K   (B(k=j, n=c) = K[j][c])
        if (live)
            warp_mma(D, m.ss, 1, warp * 16, Ks, m.st, 1, m.dk / 8, m.Tp / 8, lane,
                     [&](int r, int c, float v0, float v1) {
                         if (i0 + r < m.T) { dbase[(int64_t)(i0 + r) * 3 * d + c] = v0; dbase[(int64_t)(i0 + r) * 3 * d + c + 1] = v1; }
                     });
        // dK (+)= dS^T Q_blk, dV (+)= P^T dO_blk: key m-tiles are distributed over the warps; A(m=j, k=r) = X[r][j]
        const int rows_blk = (m.Tp - i0) < QB ? (m.Tp - i0) : QB;     // tile rows that exist in this block
        for (int mt = warp; mt < m.Tp / 16; mt += MHA_WARPS) {
            warp_mma(D, 1, m.ss, mt * 16, Qs + i0 * m.st, m.st, 1, m.dk / 8, rows_blk / 8, lane,
                     [&](int j, int c, float v0, float v1) {
                         if (j < m.T) {
                             float* o = dbase + (int64_t)j * 3 * d + d + c;
                             if (i0 == 0) { o[0] = v0; o[1] = v1; } else { o[0] += v0; o[1] += v1; }
                         }
                     });
            warp_mma(P, 1, m.ss, mt * 16, Gs + i0 * m.st, m.st, 1, m.dk / 8, rows_blk / 8, lane,
                     [&](int j, int c, float v0, float v1) {
                         if (j < m.T) {
                             float* o = dbase + (int64_t)j * 3 * d + 2 * d + c;
                             if (i0 == 0) { o[0] = v0; o[1] = v1; } else { o[0] += v0; o[1] += v1; }
                         }
                     });
        }
    }
}

int mha_bwd(int64_t B, int64_t T, int d, int heads, const float* QKV, const int64_t* lens, const float* dO, float* dQKV,
            cudaStream_t s) {
    if (B <= 0 || T <= 0) return INTEL_OK;
    INTEL_REQUIRE(heads > 0 && d % heads == 0 && (d / heads) % 8 == 0, INTEL_ERR_UNSUPPORTED,
                  "mha: head width d/heads = %d/%d must be a multiple of 8", d, heads);
    const size_t smem = mha_bwd_smem(T, d / heads);
    INTEL_REQUIRE(smem <= kMaxSmem, INTEL_ERR_UNSUPPORTED,
                  "mha_bwd: list length %lld with head width %d needs %zu bytes of shared memory (max %zu)", (long long)T,
                  d / heads, smem, kMaxSmem);
    INTEL_REQUIRE(d % 4 == 0 && ((uintptr_t)QKV % 16 == 0) && ((uintptr_t)dO % 16 == 0), INTEL_ERR_ARG,
                  "mha: qkv / d_out must be 16-byte aligned");
    const dim3 grid((unsigned)(B * heads)), block(MHA_WARPS * 32);
#define INTEL_MHA_BWD(DKV)                                                                                         \
    case DKV: {                                                                                                    \
        auto k = mha_bwd_kernel<DKV>;                                                                              \
        ensure_smem(k, smem);     \
        LAUNCH(k, grid, block, smem, s, T, d, heads, QKV, lens, dO, dQKV);                                         \
    } break;
    switch (d / heads) {
        INTEL_MHA_BWD(8) INTEL_MHA_BWD(16) INTEL_MHA_BWD(24) INTEL_MHA_BWD(32) INTEL_MHA_BWD(48) INTEL_MHA_BWD(64)
        default:
            set_error("mha: head width %d is not instantiated (8,16,24,32,48,64)", d / heads);
            return INTEL_ERR_UNSUPPORTED;
    }
#undef INTEL_MHA_BWD
    return check_launch("mha_bwd", 28.0 * B * T * d, 16.0 * B * T * T * d);
}

}  // namespace intel
