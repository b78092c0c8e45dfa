// Fused self-attention stack of IntEL.predict_ensemble (IntEL.py:182-197, layers.py:31-60) for the shape
// every script of the reference uses: stream width d = 32, list length L <= 64.
//
// One CTA works on one session at a time (persistent over sessions).  The five 32x32 weights, the session's
// token tile and every intermediate (Q, K, V, scores, attention output, FFN hidden, pre-LN sum) live in
// shared memory for all N weight-shared layers; every product is a warp-level 3xTF32 tensor-core MMA
// (mma.cuh).  The forward pass writes nothing but each layer's output X[l+1] (the only thing the backward
// pass needs besides X[0]); the backward pass recomputes the layer on chip, back-propagates through it and
// accumulates the weight gradients in shared memory ACROSS sessions, flushing them with one atomicAdd per
// element and CTA at the end.  Compared with the staged kernels this removes ~30 launches and all
// [B*L,32..96] round trips through HBM per layer and stream.
#include "kernels.h"
#include "mma.cuh"

namespace intel {

static const int TD = 32;            // stream width
static const int TS = TD + 4;        // tile row stride (conflict-free fragment reads)
static const int TW = 8;             // warps per CTA

struct TrunkArgs {
    int64_t B;
    int L, heads, layers;
    const float *wq, *wk, *wv, *w1, *b1, *w2, *b2, *lnw, *lnb;
    float* X[9];                     // X[0] input, X[l+1] output of layer l, each [B*L, 32]
    // activations of layer l kept for the backward pass (written by the forward kernel when save != 0):
    float *QKV[8], *A[8], *U[8], *Z[8], *ST[8];   // [B*L,96] q|k|v, [B*L,32] x3, [B*L,2] LN mean / rstd
    int save;
    Dropout drop[8];                 // per layer (p = 0: off)
    // backward only
    float* dX;                       // [B*L,32]: d loss / d X[layers] on entry, d loss / d X[0] on return
    float *gwq, *gwk, *gwv, *gw1, *gb1, *gw2, *gb2, *glnw, *glnb;
};

// acc(16 x 8*NTN) += A * B for one warp: rows m0.., n-tiles nt0.., A(m,k) = A[m*ARS + k*ACS], B(k,n) = B[k*BKS + n*BNS]
template <int NTN, int KS, int ARS, int ACS, int BKS, int BNS>
__device__ __forceinline__ void tile_mma_acc(float (&acc)[NTN][4], const float* __restrict__ A, const float* __restrict__ B,
                                             int m0, int nt0, int lane) {
    const int gq = lane >> 2, tq = lane & 3;
    // per-lane base pointers; every further offset is a compile-time constant once the k loop is unrolled
    const float* pa = A + (m0 + gq) * ARS + tq * ACS;
    const float* pb = B + tq * BKS + (nt0 * 8 + gq) * BNS;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4];
        split_tf32(pa[(ks * 8) * ACS], ah[0], al[0]);
        split_tf32(pa[8 * ARS + (ks * 8) * ACS], ah[1], al[1]);
        split_tf32(pa[(ks * 8 + 4) * ACS], ah[2], al[2]);
        split_tf32(pa[8 * ARS + (ks * 8 + 4) * ACS], ah[3], al[3]);
        uint32_t bh[NTN][2], bl[NTN][2];
#pragma unroll
        for (int j = 0; j < NTN; ++j) {
            split_tf32(pb[(ks * 8) * BKS + (j * 8) * BNS], bh[j][0], bl[j][0]);
            split_tf32(pb[(ks * 8 + 4) * BKS + (j * 8) * BNS], bh[j][1], bl[j][1]);
        }
#pragma unroll
        for (int j = 0; j < NTN; ++j) mma_tf32(acc[j], al, bh[j]);
#pragma unroll
        for (int j = 0; j < NTN; ++j) mma_tf32(acc[j], ah, bl[j]);
#pragma unroll
        for (int j = 0; j < NTN; ++j) mma_tf32(acc[j], ah, bh[j]);
    }
}

template <int NTN, int KS, int ARS, int ACS, int BKS, int BNS, class Epi>
__device__ __forceinline__ void tile_mma(const float* __restrict__ A, const float* __restrict__ B, int m0, int nt0, int lane,
                                         Epi epi) {
    const int gq = lane >> 2, tq = lane & 3;
    float acc[NTN][4];
#pragma unroll
    for (int j = 0; j < NTN; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
    tile_mma_acc<NTN, KS, ARS, ACS, BKS, BNS>(acc, A, B, m0, nt0, lane);
#pragma unroll
    for (int j = 0; j < NTN; ++j) {
        const int col = (nt0 + j) * 8 + 2 * tq;
        epi(m0 + gq, col, acc[j][0], acc[j][1]);
        epi(m0 + gq + 8, col, acc[j][2], acc[j][3]);
    }
}

// shared-memory plan shared by the forward and the backward kernel
template <int TP>
struct TrunkSmem {
    static constexpr int SS = TP + 4;                 // score block stride
    static constexpr int TILE = TP * TS;
    static constexpr int WMAT = TD * TS;
    // weights (5 matrices + 4 vectors), then tiles
    static constexpr int off_w = 0;                                   // wq wk wv w1 w2
    static constexpr int off_v = off_w + 5 * WMAT;                    // b1 b2 lnw lnb
    static constexpr int off_x = off_v + 4 * TD;
    static constexpr int fwd_floats = off_x + 5 * TILE + TP * SS;                     // X Q K V A + S
    static constexpr int bwd_floats = off_x + 10 * TILE + 2 * TP * SS + 5 * WMAT + 4 * TD + 2 * TP;
};

__device__ __forceinline__ void stage_weights(float* sm, const TrunkArgs& a) {
    const float* src[5] = {a.wq, a.wk, a.wv, a.w1, a.w2};
    for (int m = 0; m < 5; ++m)
        for (int e = threadIdx.x; e < TD * (TD / 4); e += blockDim.x) {
            const int r = e / (TD / 4), c = (e % (TD / 4)) * 4;
            *reinterpret_cast<float4*>(sm + m * TD * TS + r * TS + c) = *reinterpret_cast<const float4*>(src[m] + r * TD + c);
        }
    float* v = sm + 5 * TD * TS;
    for (int e = threadIdx.x; e < TD; e += blockDim.x) {
        v[e] = a.b1[e]; v[TD + e] = a.b2[e]; v[2 * TD + e] = a.lnw[e]; v[3 * TD + e] = a.lnb[e];
    }
}

template <int TP>
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, int L) {
    for (int e = threadIdx.x; e < TP * (TD / 4); e += blockDim.x) {
        const int r = e / (TD / 4), c = (e % (TD / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < L) v = *reinterpret_cast<const float4*>(src + (int64_t)r * TD + c);
        *reinterpret_cast<float4*>(dst + r * TS + c) = v;
    }
}

// One layer forward on the tiles of a session.  Xs (input) is left intact; Zs receives the pre-LN sum,
// Us the FFN hidden (pre-relu), As the attention output, stats (nullable) the LN mean / rstd per row.
// Xout tile gets LN(Z).  Pad rows (>= L) of every tile stay exactly zero.
template <int TP, int DK>
__device__ __forceinline__ void layer_forward(const float* W, const float* V, const float* Xs, float* Qs, float* Ks, float* Vs,
                                              float* As, float* Us, float* Zs, float* S, float* Xout, float* stats, int L,
                                              int heads, int lane, int warp, float* gQKV, float* gA, float* gU, float* gZ,
                                              float* gST, const Dropout& dr, int64_t row0) {
    constexpr int SS = TrunkSmem<TP>::SS, WMAT = TrunkSmem<TP>::WMAT, MT = TP / 16;
    const int TW = blockDim.x >> 5;
    const float scale = 1.0f / sqrtf((float)DK);
    // Q K V = X W^T   (B(k, n) = W[n][k]); work item = (matrix, m-tile, n-half)
    for (int it = warp; it < 3 * MT * 2; it += TW) {
        const int mat = it / (MT * 2), mt = (it / 2) % MT, nh = it & 1;
        float* dst = mat == 0 ? Qs : (mat == 1 ? Ks : Vs);
        tile_mma<2, TD / 8, TS, 1, 1, TS>(Xs, W + mat * WMAT, mt * 16, nh * 2, lane,
                                          [&](int r, int c, float v0, float v1) { dst[r * TS + c] = v0; dst[r * TS + c + 1] = v1; });
    }
    __syncthreads();
    if (gQKV) {
        for (int e = threadIdx.x; e < L * 3 * (TD / 4); e += blockDim.x) {
            const int r = e / (3 * TD / 4), q = (e / (TD / 4)) % 3, c = (e % (TD / 4)) * 4;
            const float* src = q == 0 ? Qs : (q == 1 ? Ks : Vs);
            *reinterpret_cast<float4*>(gQKV + (int64_t)r * 3 * TD + q * TD + c) = *reinterpret_cast<const float4*>(src + r * TS + c);
        }
    }
    for (int hd = 0; hd < heads; ++hd) {
        const int ho = hd * DK;
        // S = scale * Q_h K_h^T over all TP slots (pad slots are live keys: their K rows are W_k * 0 = 0 -> score 0)
        for (int it = warp; it < MT * MT; it += TW) {
            const int mt = it % MT, nq = it / MT;
            tile_mma<2, DK / 8, TS, 1, 1, TS>(Qs + ho, Ks + ho, mt * 16, nq * 2, lane,
                                              [&](int r, int c, float v0, float v1) { S[r * SS + c] = v0 * scale; S[r * SS + c + 1] = v1 * scale; });
        }
        __syncthreads();
        // softmax over the L live slots of every row (rows >= L are zeroed)
        for (int r = warp; r < TP; r += TW) {
            float* s = S + r * SS;
            if (r >= L) { for (int j = lane; j < TP; j += 32) s[j] = 0.f; continue; }
            float mx = -INFINITY;
            for (int j = lane; j < L; j += 32) mx = fmaxf(mx, s[j]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int j = lane; j < L; j += 32) { const float e = expf(s[j] - mx); s[j] = e; sum += e; }
            sum = warp_sum(sum);
            const float inv = 1.0f / sum;
            for (int j = lane; j < TP; j += 32) s[j] = (j < L) ? s[j] * inv : 0.f;
        }
        __syncthreads();
        // A_h = P V_h   (B(k = j, n = c) = V[j][ho + c])
        for (int it = warp; it < MT * (DK / 16); it += TW) {
            const int mt = it % MT, nq = it / MT;
            tile_mma<2, TP / 8, SS, 1, TS, 1>(S, Vs + ho, mt * 16, nq * 2, lane,
                                              [&](int r, int c, float v0, float v1) { As[r * TS + ho + c] = v0; As[r * TS + ho + c + 1] = v1; });
        }
        __syncthreads();
    }
    if (gA) {
        for (int e = threadIdx.x; e < L * (TD / 4); e += blockDim.x) {
            const int r = e / (TD / 4), c = (e % (TD / 4)) * 4;
            *reinterpret_cast<float4*>(gA + (int64_t)r * TD + c) = *reinterpret_cast<const float4*>(As + r * TS + c);
        }
    }
    // U = A W1^T + b1 (kept pre-relu)
    for (int it = warp; it < MT * 2; it += TW) {
        const int mt = it / 2, nh = it & 1;
        tile_mma<2, TD / 8, TS, 1, 1, TS>(As, W + 3 * WMAT, mt * 16, nh * 2, lane, [&](int r, int c, float v0, float v1) {
            Us[r * TS + c] = (r < L) ? v0 + V[c] : 0.f;
            Us[r * TS + c + 1] = (r < L) ? v1 + V[c + 1] : 0.f;
        });
    }
    __syncthreads();
    if (gU) {
        for (int e = threadIdx.x; e < L * (TD / 4); e += blockDim.x) {
            const int r = e / (TD / 4), c = (e % (TD / 4)) * 4;
            *reinterpret_cast<float4*>(gU + (int64_t)r * TD + c) = *reinterpret_cast<const float4*>(Us + r * TS + c);
        }
    }
    // Z = relu(U) W2^T + b2 + X
    for (int it = warp; it < MT * 2; it += TW) {
        const int mt = it / 2, nh = it & 1;
        const int gq = lane >> 2, tq = lane & 3;
        // relu on the A operand: tile_mma reads A through a pointer, so materialise relu lazily via a lambda-free trick:
        // the hidden tile is read with fmaxf inside a dedicated loop below.
        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
        const float* Bw = W + 4 * WMAT;
#pragma unroll
        for (int ks = 0; ks < TD / 8; ++ks) {
            const int k0 = ks * 8 + tq, m0 = mt * 16;
            const float a0 = fmaxf(Us[(m0 + gq) * TS + k0], 0.f), a1 = fmaxf(Us[(m0 + gq + 8) * TS + k0], 0.f);
            const float a2 = fmaxf(Us[(m0 + gq) * TS + k0 + 4], 0.f), a3 = fmaxf(Us[(m0 + gq + 8) * TS + k0 + 4], 0.f);
            uint32_t ah[4], al[4];
            split_tf32(a0, ah[0], al[0]);
            split_tf32(a1, ah[1], al[1]);
            split_tf32(a2, ah[2], al[2]);
            split_tf32(a3, ah[3], al[3]);
            uint32_t bh[2][2], bl[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = (nh * 2 + j) * 8 + gq;
                const float b0 = Bw[n * TS + k0], b1 = Bw[n * TS + k0 + 4];
                split_tf32(b0, bh[j][0], bl[j][0]);
                split_tf32(b1, bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) mma_tf32(acc[j], al, bh[j]);
#pragma unroll
            for (int j = 0; j < 2; ++j) mma_tf32(acc[j], ah, bl[j]);
#pragma unroll
            for (int j = 0; j < 2; ++j) mma_tf32(acc[j], ah, bh[j]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = (nh * 2 + j) * 8 + 2 * tq;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int r = mt * 16 + gq + 8 * hh;
                // dropout acts on the FFN output before the residual (IntEL.py:187-188)
                Zs[r * TS + c] = (r < L) ? (acc[j][2 * hh] + V[TD + c]) * dropout_scale(dr, row0 + r, c, TD) + Xs[r * TS + c] : 0.f;
                Zs[r * TS + c + 1] =
                    (r < L) ? (acc[j][2 * hh + 1] + V[TD + c + 1]) * dropout_scale(dr, row0 + r, c + 1, TD) + Xs[r * TS + c + 1] : 0.f;
            }
        }
    }
    __syncthreads();
    // LayerNorm (eps 1e-5): one lane per channel, one warp per row
    for (int r = warp; r < TP; r += TW) {
        if (r >= L) { Xout[r * TS + lane] = 0.f; continue; }
        const float z = Zs[r * TS + lane];
        const float mean = warp_sum(z) * (1.0f / TD);
        const float dlt = z - mean;
        const float rstd = rsqrtf(warp_sum(dlt * dlt) * (1.0f / TD) + 1e-5f);
        Xout[r * TS + lane] = dlt * rstd * V[2 * TD + lane] + V[3 * TD + lane];
        if (stats && lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
        if (gZ) {
            gZ[(int64_t)r * TD + lane] = z;
            if (lane == 0) { gST[2 * r] = mean; gST[2 * r + 1] = rstd; }
        }
    }
    __syncthreads();
}

template <int TP, int DK>
__global__ void __launch_bounds__(TW * 32) trunk_fwd_kernel(TrunkArgs a) {
    DYN_SMEM(float, sm);
    using SMP = TrunkSmem<TP>;
    float* W = sm + SMP::off_w;
    float* V = sm + SMP::off_v;
    float* Xs = sm + SMP::off_x;
    float* Qs = Xs + SMP::TILE;
    float* Ks = Qs + SMP::TILE;
    float* Vs = Ks + SMP::TILE;
    float* As = Vs + SMP::TILE;
    float* S = As + SMP::TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    stage_weights(sm, a);
    for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();
        load_tile<TP>(Xs, a.X[0] + b * a.L * TD, a.L);
        __syncthreads();
        for (int l = 0; l < a.layers; ++l) {
            // U reuses the Q tile, Z the K tile (both are dead once the attention output exists)
            const int64_t ro = b * a.L;
            const bool sv = a.save != 0;
            layer_forward<TP, DK>(W, V, Xs, Qs, Ks, Vs, As, Qs, Ks, S, Xs, nullptr, a.L, a.heads, lane, warp,
                                  sv ? a.QKV[l] + ro * 3 * TD : nullptr, sv ? a.A[l] + ro * TD : nullptr,
                                  sv ? a.U[l] + ro * TD : nullptr, sv ? a.Z[l] + ro * TD : nullptr,
                                  sv ? a.ST[l] + ro * 2 : nullptr, a.drop[l], ro);
            float* out = a.X[l + 1] + b * a.L * TD;
            for (int e = threadIdx.x; e < a.L * (TD / 4); e += blockDim.x) {
                const int r = e / (TD / 4), c = (e % (TD / 4)) * 4;
                *reinterpret_cast<float4*>(out + (int64_t)r * TD + c) = *reinterpret_cast<const float4*>(Xs + r * TS + c);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward: for each session, layers in reverse; everything is recomputed from X[l]
template <int TP, int DK>
__global__ void __launch_bounds__(512) trunk_bwd_kernel(TrunkArgs a) {
    DYN_SMEM(float, sm);
    using SMP = TrunkSmem<TP>;
    constexpr int SS = SMP::SS, TILE = SMP::TILE, WMAT = SMP::WMAT, MT = TP / 16;
    float* W = sm + SMP::off_w;
    float* V = sm + SMP::off_v;
    float* Xs = sm + SMP::off_x;          // layer input
    float* Qs = Xs + TILE;
    float* Ks = Qs + TILE;
    float* Vs = Ks + TILE;
    float* As = Vs + TILE;                // attention output, then dA
    float* Us = As + TILE;                // FFN hidden (pre-relu), then dU
    float* Zs = Us + TILE;                // pre-LN sum, then dZ
    float* Gs = Zs + TILE;                // incoming gradient d X[l+1]; then dX[l]
    float* dQs = Gs + TILE;
    float* dKs = dQs + TILE;              // (dV is written over the V tile's sibling below)
    float* P = dKs + TILE;                // [TP][SS] probabilities of the current head
    float* D = P + TP * SS;               // [TP][SS] dP, then dS
    float* GW = D + TP * SS;              // weight-gradient accumulators: wq wk wv w1 w2
    float* GV = GW + 5 * WMAT;            // b1 b2 lnw lnb
    float* stats = GV + 4 * TD;           // [TP][2]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const float scale = 1.0f / sqrtf((float)DK);
    const int TW = blockDim.x >> 5;                 // 16 warps: the weight-gradient blocks split the tokens in halves
    const int w8 = warp & 7, kh = warp >> 3, KH = TW >> 3;
    constexpr int KSTEPS = TP / 8;
    const int ks_beg = kh * (KSTEPS / KH), ks_end = (kh == KH - 1) ? KSTEPS : (kh + 1) * (KSTEPS / KH);
    stage_weights(sm, a);
    for (int e = threadIdx.x; e < 5 * WMAT + 4 * TD; e += blockDim.x) GW[e] = 0.f;
    const int L = a.L;

    for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();
        load_tile<TP>(Gs, a.dX + b * L * TD, L);
        for (int l = a.layers - 1; l >= 0; --l) {
            __syncthreads();
            // ---- reload the layer's activations saved by the forward kernel ----
            {
                const int64_t ro = b * L;
                load_tile<TP>(Xs, a.X[l] + ro * TD, L);
                load_tile<TP>(As, a.A[l] + ro * TD, L);
                load_tile<TP>(Us, a.U[l] + ro * TD, L);
                load_tile<TP>(Zs, a.Z[l] + ro * TD, L);
                const float* gq = a.QKV[l] + ro * 3 * TD;
                for (int e = threadIdx.x; e < TP * 3 * (TD / 4); e += blockDim.x) {
                    const int r = e / (3 * TD / 4), q = (e / (TD / 4)) % 3, c = (e % (TD / 4)) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < L) v = *reinterpret_cast<const float4*>(gq + (int64_t)r * 3 * TD + q * TD + c);
                    float* dst = q == 0 ? Qs : (q == 1 ? Ks : Vs);
                    *reinterpret_cast<float4*>(dst + r * TS + c) = v;
                }
                for (int e = threadIdx.x; e < 2 * L; e += blockDim.x) stats[e] = a.ST[l][ro * 2 + e];
            }
            __syncthreads();
            // ---- LayerNorm backward: dZ (into Zs), d gamma / d beta ----
            {
                float pg = 0.f, pb = 0.f;
                for (int r = warp; r < L; r += TW) {
                    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
                    const float xh = (Zs[r * TS + lane] - mean) * rstd;
                    const float dy = Gs[r * TS + lane];
                    const float g = dy * V[2 * TD + lane];
                    pg = fmaf(dy, xh, pg);
                    pb += dy;
                    const float s1 = warp_sum(g) * (1.0f / TD), s2 = warp_sum(g * xh) * (1.0f / TD);
                    Zs[r * TS + lane] = rstd * (g - s1 - xh * s2);
                }
                atomicAdd(GV + 2 * TD + lane, pg);      // shared-memory atomics: 8 warps per channel
                atomicAdd(GV + 3 * TD + lane, pb);
            }
            __syncthreads();
            // dropout sits between the FFN output and the residual: the FFN branch sees dZ * mask / (1-p),
            // the residual branch the plain dZ.  The masked copy lives in the (still unused) dQ tile.
            const float* Fs = Zs;
            if (a.drop[l].p > 0.f) {
                for (int e = threadIdx.x; e < TP * TD; e += blockDim.x) {
                    const int r = e / TD, c = e % TD;
                    dQs[r * TS + c] = (r < L) ? Zs[r * TS + c] * dropout_scale(a.drop[l], b * L + r, c, TD) : 0.f;
                }
                Fs = dQs;
                __syncthreads();
            }
            // ---- FFN backward ----
            // dW2 += dZ^T relu(U), db2 += colsum(dZ); dU = (dZ W2) * (U > 0); 8 (m,n) tile pairs -> one per warp
            {
                const int mt = w8 & 1, np = w8 >> 1;                // output rows 16*mt.., n-tile np
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
                for (int ks = ks_beg; ks < ks_end; ++ks) {
                    const int k0 = ks * 8 + tq;
                    float af[4] = {Fs[k0 * TS + mt * 16 + gq], Fs[k0 * TS + mt * 16 + gq + 8], Fs[(k0 + 4) * TS + mt * 16 + gq],
                                   Fs[(k0 + 4) * TS + mt * 16 + gq + 8]};
                    float bf[2] = {fmaxf(Us[k0 * TS + np * 8 + gq], 0.f), fmaxf(Us[(k0 + 4) * TS + np * 8 + gq], 0.f)};
                    mma_3xtf32(acc, af, bf);
                }
                float* g2 = GW + 4 * WMAT;
                const int r = mt * 16 + gq, c = np * 8 + 2 * tq;
                atomicAdd(g2 + r * TS + c, acc[0]); atomicAdd(g2 + r * TS + c + 1, acc[1]);
                atomicAdd(g2 + (r + 8) * TS + c, acc[2]); atomicAdd(g2 + (r + 8) * TS + c + 1, acc[3]);
                if (warp == 0) {
                    float sb = 0.f;
                    for (int r2 = 0; r2 < L; ++r2) sb += Fs[r2 * TS + lane];
                    GV[TD + lane] += sb;
                }
            }
            __syncthreads();
            for (int it = warp; it < MT * 2; it += TW) {            // dU = (dZ W2) * (U > 0), in place over U
                const int mt = it / 2, nh = it & 1;
                tile_mma<2, TD / 8, TS, 1, TS, 1>(Fs, W + 4 * WMAT, mt * 16, nh * 2, lane, [&](int r, int c, float v0, float v1) {
                    Us[r * TS + c] = Us[r * TS + c] > 0.f ? v0 : 0.f;
                    Us[r * TS + c + 1] = Us[r * TS + c + 1] > 0.f ? v1 : 0.f;
                });
            }
            __syncthreads();
            {   // dW1 += dU^T A, db1 += colsum(dU)
                const int mt = w8 & 1, np = w8 >> 1;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
                for (int ks = ks_beg; ks < ks_end; ++ks) {
                    const int k0 = ks * 8 + tq;
                    float af[4] = {Us[k0 * TS + mt * 16 + gq], Us[k0 * TS + mt * 16 + gq + 8], Us[(k0 + 4) * TS + mt * 16 + gq],
                                   Us[(k0 + 4) * TS + mt * 16 + gq + 8]};
                    float bf[2] = {As[k0 * TS + np * 8 + gq], As[(k0 + 4) * TS + np * 8 + gq]};
                    mma_3xtf32(acc, af, bf);
                }
                float* g1 = GW + 3 * WMAT;
                const int r = mt * 16 + gq, c = np * 8 + 2 * tq;
                atomicAdd(g1 + r * TS + c, acc[0]); atomicAdd(g1 + r * TS + c + 1, acc[1]);
                atomicAdd(g1 + (r + 8) * TS + c, acc[2]); atomicAdd(g1 + (r + 8) * TS + c + 1, acc[3]);
                if (warp == 0) {
                    float sb = 0.f;
                    for (int r2 = 0; r2 < L; ++r2) sb += Us[r2 * TS + lane];
                    GV[lane] += sb;
                }
            }
            __syncthreads();
            for (int it = warp; it < MT * 2; it += TW) {            // dA = dU W1, over the A tile
                const int mt = it / 2, nh = it & 1;
                tile_mma<2, TD / 8, TS, 1, TS, 1>(Us, W + 3 * WMAT, mt * 16, nh * 2, lane,
                                                  [&](int r, int c, float v0, float v1) { As[r * TS + c] = v0; As[r * TS + c + 1] = v1; });
            }
            __syncthreads();
            // ---- attention backward, head by head: dQ -> dQs, dK -> dKs, dV -> Us (dU is dead) ----
            float* dVs = Us;
            for (int hd = 0; hd < a.heads; ++hd) {
                const int ho = hd * DK;
                for (int it = warp; it < MT * MT; it += TW) {               // P (scores) and dP = dA_h V_h^T
                    const int mt = it % MT, nq = it / MT;
                    tile_mma<2, DK / 8, TS, 1, 1, TS>(Qs + ho, Ks + ho, mt * 16, nq * 2, lane,
                                                      [&](int r, int c, float v0, float v1) { P[r * SS + c] = v0 * scale; P[r * SS + c + 1] = v1 * scale; });
                    tile_mma<2, DK / 8, TS, 1, 1, TS>(As + ho, Vs + ho, mt * 16, nq * 2, lane,
                                                      [&](int r, int c, float v0, float v1) { D[r * SS + c] = v0; D[r * SS + c + 1] = v1; });
                }
                __syncthreads();
                for (int r = warp; r < TP; r += TW) {                       // softmax, then dS = P (dP - sum P dP) * scale
                    float* p = P + r * SS;
                    float* g = D + r * SS;
                    if (r >= L) { for (int j = lane; j < TP; j += 32) { p[j] = 0.f; g[j] = 0.f; } continue; }
                    float mx = -INFINITY;
                    for (int j = lane; j < L; j += 32) mx = fmaxf(mx, p[j]);
                    mx = warp_max(mx);
                    float sum = 0.f;
                    for (int j = lane; j < L; j += 32) { const float e = expf(p[j] - mx); p[j] = e; sum += e; }
                    sum = warp_sum(sum);
                    const float inv = 1.0f / sum;
                    float delta = 0.f;
                    for (int j = lane; j < TP; j += 32) {
                        const float pj = (j < L) ? p[j] * inv : 0.f;
                        p[j] = pj;
                        delta = fmaf(pj, (j < L) ? g[j] : 0.f, delta);
                    }
                    delta = warp_sum(delta);
                    for (int j = lane; j < TP; j += 32) g[j] = (j < L) ? p[j] * (g[j] - delta) * scale : 0.f;
                }
                __syncthreads();
                for (int it = warp; it < MT * (DK / 16); it += TW) {        // dQ_h = dS K_h
                    const int mt = it % MT, nq = it / MT;
                    tile_mma<2, TP / 8, SS, 1, TS, 1>(D, Ks + ho, mt * 16, nq * 2, lane,
                                                      [&](int r, int c, float v0, float v1) { dQs[r * TS + ho + c] = v0; dQs[r * TS + ho + c + 1] = v1; });
                }
                for (int it = warp; it < MT * (DK / 16); it += TW) {        // dK_h = dS^T Q_h ; dV_h = P^T dA_h
                    const int mt = it % MT, nq = it / MT;
                    tile_mma<2, TP / 8, 1, SS, TS, 1>(D, Qs + ho, mt * 16, nq * 2, lane,
                                                      [&](int r, int c, float v0, float v1) { dKs[r * TS + ho + c] = v0; dKs[r * TS + ho + c + 1] = v1; });
                    tile_mma<2, TP / 8, 1, SS, TS, 1>(P, As + ho, mt * 16, nq * 2, lane,
                                                      [&](int r, int c, float v0, float v1) { dVs[r * TS + ho + c] = v0; dVs[r * TS + ho + c + 1] = v1; });
                }
                __syncthreads();
            }
            // ---- projections backward: dWq += dQ^T X (same for k, v);  dX = dQ Wq + dK Wk + dV Wv + dZ ----
            {
                const int mt = w8 & 1, np = w8 >> 1;
                const int r = mt * 16 + gq, c = np * 8 + 2 * tq;
                const float* src[3] = {dQs, dKs, dVs};
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
                    for (int ks = ks_beg; ks < ks_end; ++ks) {
                        const int k0 = ks * 8 + tq;
                        const float* G = src[m];
                        float af[4] = {G[k0 * TS + mt * 16 + gq], G[k0 * TS + mt * 16 + gq + 8], G[(k0 + 4) * TS + mt * 16 + gq],
                                       G[(k0 + 4) * TS + mt * 16 + gq + 8]};
                        float bf[2] = {Xs[k0 * TS + np * 8 + gq], Xs[(k0 + 4) * TS + np * 8 + gq]};
                        mma_3xtf32(acc, af, bf);
                    }
                    float* gm = GW + m * WMAT;
                    atomicAdd(gm + r * TS + c, acc[0]); atomicAdd(gm + r * TS + c + 1, acc[1]);
                    atomicAdd(gm + (r + 8) * TS + c, acc[2]); atomicAdd(gm + (r + 8) * TS + c + 1, acc[3]);
                }
            }
            for (int it = warp; it < MT * 2; it += TW) {
                const int mt = it / 2, nh = it & 1;
                float part[1][2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int c = 0; c < 4; ++c) part[0][j][c] = 0.f;
                tile_mma_acc<2, TD / 8, TS, 1, TS, 1>(part[0], dQs, W + 0 * WMAT, mt * 16, nh * 2, lane);
                tile_mma_acc<2, TD / 8, TS, 1, TS, 1>(part[0], dKs, W + 1 * WMAT, mt * 16, nh * 2, lane);
                tile_mma_acc<2, TD / 8, TS, 1, TS, 1>(part[0], dVs, W + 2 * WMAT, mt * 16, nh * 2, lane);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = (nh * 2 + j) * 8 + 2 * tq;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int r = mt * 16 + gq + 8 * hh;
                        Gs[r * TS + c] = (r < L) ? part[0][j][2 * hh] + Zs[r * TS + c] : 0.f;
                        Gs[r * TS + c + 1] = (r < L) ? part[0][j][2 * hh + 1] + Zs[r * TS + c + 1] : 0.f;
                    }
                }
            }
        }
        __syncthreads();
        float* out = a.dX + b * L * TD;
        for (int e = threadIdx.x; e < L * (TD / 4); e += blockDim.x) {
            const int r = e / (TD / 4), c = (e % (TD / 4)) * 4;
            *reinterpret_cast<float4*>(out + (int64_t)r * TD + c) = *reinterpret_cast<const float4*>(Gs + r * TS + c);
        }
    }
    __syncthreads();
    float* dst[5] = {a.gwq, a.gwk, a.gwv, a.gw1, a.gw2};
    for (int m = 0; m < 5; ++m)
        for (int e = threadIdx.x; e < TD * TD; e += blockDim.x) atomicAdd(dst[m] + e, GW[m * WMAT + (e / TD) * TS + (e % TD)]);
    for (int e = threadIdx.x; e < TD; e += blockDim.x) {
        atomicAdd(a.gb1 + e, GV[e]);
        atomicAdd(a.gb2 + e, GV[TD + e]);
        atomicAdd(a.glnw + e, GV[2 * TD + e]);
        atomicAdd(a.glnb + e, GV[3 * TD + e]);
    }
}

bool trunk_supported(int64_t L, int d, int heads, int layers) {
    return d == TD && L >= 1 && L <= 64 && layers >= 1 && layers <= 8 && (heads == 1 || heads == 2);
}

template <int TP, int DK>
static int trunk_launch(const TrunkArgs& a, bool bwd, cudaStream_t s) {
    const size_t smem = (size_t)(bwd ? TrunkSmem<TP>::bwd_floats : TrunkSmem<TP>::fwd_floats) * 4;
    const int per_sm = bwd ? 1 : 2;
    const unsigned grid = stream_grid(a.B, per_sm);
    if (bwd) {
        auto k = trunk_bwd_kernel<TP, DK>;
        ensure_smem(k, smem);
        LAUNCH(k, dim3(grid), dim3(512), smem, s, a);
    } else {
        auto k = trunk_fwd_kernel<TP, DK>;
        ensure_smem(k, smem);
        LAUNCH(k, dim3(grid), dim3(TW * 32), smem, s, a);
    }
    const double tok = (double)a.B * a.L * a.layers;
    const double flops = tok * (2.0 * 5 * TD * TD + 4.0 * a.L * TD) * (bwd ? 3.0 : 1.0);
    return check_launch(bwd ? "trunk_bwd" : "trunk_fwd", (double)a.B * a.L * TD * 4.0 * (a.layers + 1) * (bwd ? 1.0 : 1.0) +
                                                             (bwd ? 2.0 * a.B * a.L * TD * 4.0 : 0.0), flops);
}

int trunk_run(const TrunkArgs& a, bool bwd, cudaStream_t s) {
    if (a.B <= 0) return INTEL_OK;
    INTEL_REQUIRE(trunk_supported(a.L, TD, a.heads, a.layers), INTEL_ERR_UNSUPPORTED, "fused stack: unsupported shape");
    const int tp = (a.L + 15) / 16 * 16;
#define INTEL_TRUNK(TPV)                                                            \
    case TPV:                                                                       \
        return a.heads == 1 ? trunk_launch<TPV, 32>(a, bwd, s) : trunk_launch<TPV, 16>(a, bwd, s);
    switch (tp) {
        INTEL_TRUNK(16) INTEL_TRUNK(32) INTEL_TRUNK(48) INTEL_TRUNK(64)
    }
#undef INTEL_TRUNK
    set_error("fused stack: unsupported padded length %d", tp);
    return INTEL_ERR_UNSUPPORTED;
}

static TrunkArgs trunk_args(int64_t B, int64_t L, int heads, int layers, const StackParams& p, float* const* X,
                            const StackSaved& sv, float drop_p, uint64_t drop_seed, int stream_id) {
    TrunkArgs a;
    memset(&a, 0, sizeof(a));
    for (int l = 0; l < 8; ++l) a.drop[l] = make_dropout(drop_p, drop_seed, stream_id, l);
    for (int l = 0; l < layers && l < 8; ++l) {
        a.QKV[l] = sv.QKV[l]; a.A[l] = sv.A[l]; a.U[l] = sv.U[l]; a.Z[l] = sv.Z[l]; a.ST[l] = sv.ST[l];
    }
    a.save = 1;
    a.B = B; a.L = (int)L; a.heads = heads; a.layers = layers;
    a.wq = p.wq; a.wk = p.wk; a.wv = p.wv; a.w1 = p.w1; a.b1 = p.b1; a.w2 = p.w2; a.b2 = p.b2; a.lnw = p.lnw; a.lnb = p.lnb;
    for (int l = 0; l <= layers && l < 9; ++l) a.X[l] = X[l];
    return a;
}

int trunk_fwd(int64_t B, int64_t L, int heads, int layers, const StackParams& p, float* const* X, const StackSaved& sv,
              float drop_p, uint64_t drop_seed, int stream_id, cudaStream_t s) {
    TrunkArgs a = trunk_args(B, L, heads, layers, p, X, sv, drop_p, drop_seed, stream_id);
    return trunk_run(a, false, s);
}

int trunk_bwd(int64_t B, int64_t L, int heads, int layers, const StackParams& p, const StackGrads& g, float* const* X,
              const StackSaved& sv, float* dX, float drop_p, uint64_t drop_seed, int stream_id, cudaStream_t s) {
    TrunkArgs a = trunk_args(B, L, heads, layers, p, X, sv, drop_p, drop_seed, stream_id);
    a.dX = dX;
    a.gwq = g.wq; a.gwk = g.wk; a.gwv = g.wv; a.gw1 = g.w1; a.gb1 = g.b1; a.gw2 = g.w2; a.gb2 = g.b2; a.glnw = g.lnw; a.glnb = g.lnb;
    return trunk_run(a, true, s);
}

}  // namespace intel
