// Weight head + weighted fusion of the K basic lists (IntEL.py:212-215) and the small element-wise
// kernels around it.  HBM-bound streaming: the float64 score tensor is read once per pass.
#include "kernels.h"

namespace intel {

static const int FUSE_MAX_K = 16;

// One warp per session: the K-vector of weights is per session for valid rows (all valid rows share the
// pooled cross-attention vector) and a second vector for pad rows.
__global__ void __launch_bounds__(256) head_fuse_fwd_kernel(int64_t B, int64_t L, int K,
                                                            const float* __restrict__ w_valid,
                                                            const float* __restrict__ w_pad,
                                                            const double* __restrict__ scores,
                                                            const int64_t* __restrict__ lens,
                                                            float* __restrict__ weights, float* __restrict__ ens) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < B; b += nwarps) {
        const int64_t n = lens[b];
        float wv[FUSE_MAX_K], wp[FUSE_MAX_K];
#pragma unroll
        for (int k = 0; k < FUSE_MAX_K; ++k) {
            wv[k] = (k < K) ? w_valid[b * K + k] : 0.f;
            wp[k] = (k < K) ? w_pad[b * K + k] : 0.f;
        }
        for (int64_t l = lane; l < L; l += 32) {
            const bool valid = l < n;
            const double* x = scores + (b * L + l) * K;
            float* wo = weights + (b * L + l) * K;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < FUSE_MAX_K; ++k) {
                if (k < K) {
                    const float wk = valid ? wv[k] : wp[k];
                    wo[k] = wk;
                    acc = fmaf(wk, (float)x[k], acc);
                }
            }
            ens[b * L + l] = acc;
        }
    }
}

int head_fuse_fwd(int64_t B, int64_t L, int K, const float* w_valid, const float* w_pad, const double* scores,
                  const int64_t* lens, float* weights, float* ens, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    INTEL_REQUIRE(K <= FUSE_MAX_K, INTEL_ERR_UNSUPPORTED, "model_num %d > %d", K, FUSE_MAX_K);
    unsigned grid = stream_grid(ceil_div(B, 8), 8);
    LAUNCH(head_fuse_fwd_kernel, dim3(grid), dim3(256), 0, s, B, L, K, w_valid, w_pad, scores, lens, weights, ens);
    return check_launch("head_fuse_fwd", (double)B * L * (8.0 * K + 4.0 * K + 4.0), 2.0 * B * L * K);
}

__global__ void __launch_bounds__(256) head_fuse_bwd_kernel(int64_t B, int64_t L, int K,
                                                            const float* __restrict__ d_weights,
                                                            const float* __restrict__ d_ens,
                                                            const double* __restrict__ scores,
                                                            const int64_t* __restrict__ lens,
                                                            float* __restrict__ dw_valid, float* __restrict__ dw_pad) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < B; b += nwarps) {
        const int64_t n = lens[b];
        float av[FUSE_MAX_K], ap[FUSE_MAX_K];
#pragma unroll
        for (int k = 0; k < FUSE_MAX_K; ++k) { av[k] = 0.f; ap[k] = 0.f; }
        for (int64_t l = lane; l < L; l += 32) {
            const bool valid = l < n;
            const float ge = d_ens ? d_ens[b * L + l] : 0.f;
#pragma unroll
            for (int k = 0; k < FUSE_MAX_K; ++k) {
                if (k < K) {
                    float g = ge * (float)scores[(b * L + l) * K + k];
                    if (d_weights) g += d_weights[(b * L + l) * K + k];
                    if (valid) av[k] += g; else ap[k] += g;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < FUSE_MAX_K; ++k) {
            if (k < K) {
                const float sv = warp_sum(av[k]), sp = warp_sum(ap[k]);
                if (lane == 0) { dw_valid[b * K + k] = sv; dw_pad[b * K + k] = sp; }
            }
        }
    }
}

int head_fuse_bwd(int64_t B, int64_t L, int K, const float* d_weights, const float* d_ens, const double* scores,
                  const int64_t* lens, float* dw_valid, float* dw_pad, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    INTEL_REQUIRE(K <= FUSE_MAX_K, INTEL_ERR_UNSUPPORTED, "model_num %d > %d", K, FUSE_MAX_K);
    unsigned grid = stream_grid(ceil_div(B, 8), 8);
    LAUNCH(head_fuse_bwd_kernel, dim3(grid), dim3(256), 0, s, B, L, K, d_weights, d_ens, scores, lens, dw_valid, dw_pad);
    return check_launch("head_fuse_bwd", (double)B * L * (8.0 * K + 4.0 * K + 4.0), 2.0 * B * L * K);
}

// ---- weight head + fusion in one pass (IntEL.py:212-215): the head is nn.Linear(D, K) with K = model_num (2..4), far too
// small for a GEMM launch of its own (six launches of 13..27 us each in round 2).  One warp per session: lanes own the head
// inputs d = lane + 32 j, the K dot products of the valid rows (all D inputs) and of the pad rows (inputs >= off_u only: the
// pooled cross-attention blocks are zero there) are warp sums; the backward kernel forms d(all) on the same lanes and keeps
// the head's weight / bias gradient in registers across the sessions of a CTA (one flush per CTA).

template <int KK, int DV>
__global__ void __launch_bounds__(256) head_full_fwd_kernel(int64_t B, int64_t L, int D, int off_u, const float* __restrict__ all,
                                                            const float* __restrict__ Wh, const float* __restrict__ bh,
                                                            const double* __restrict__ scores, const int64_t* __restrict__ lens,
                                                            float* __restrict__ weights, float* __restrict__ ens) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float w[KK][DV];
#pragma unroll
    for (int k = 0; k < KK; ++k)
#pragma unroll
        for (int j = 0; j < DV; ++j) { const int d = lane + 32 * j; w[k][j] = d < D ? Wh[k * D + d] : 0.f; }
    for (int64_t b = warp; b < B; b += nwarps) {
        const int64_t n = lens[b];
        float x[DV];
#pragma unroll
        for (int j = 0; j < DV; ++j) { const int d = lane + 32 * j; x[j] = d < D ? all[b * D + d] : 0.f; }
        float wv[KK], wp[KK];
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            float av = 0.f, ap = 0.f;
#pragma unroll
            for (int j = 0; j < DV; ++j) {
                const float t = x[j] * w[k][j];
                av += t;
                if (lane + 32 * j >= off_u) ap += t;
            }
            wv[k] = warp_sum(av) + bh[k];
            wp[k] = warp_sum(ap) + bh[k];
        }
        for (int64_t l = lane; l < L; l += 32) {
            const bool valid = l < n;
            const double* xs = scores + (b * L + l) * KK;
            float* wo = weights + (b * L + l) * KK;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < KK; ++k) {
                const float wk = valid ? wv[k] : wp[k];
                wo[k] = wk;
                acc = fmaf(wk, (float)xs[k], acc);
            }
            ens[b * L + l] = acc;
        }
    }
}

template <int KK, int DV>
__global__ void __launch_bounds__(256) head_full_bwd_kernel(int64_t B, int64_t L, int D, int off_u, const float* __restrict__ all,
                                                            const float* __restrict__ Wh, const float* __restrict__ d_weights,
                                                            const float* __restrict__ d_ens, const double* __restrict__ scores,
                                                            const int64_t* __restrict__ lens, float* __restrict__ dall,
                                                            float* __restrict__ gWh, float* __restrict__ gbh) {
    __shared__ float red[8][KK * DV * 32 + KK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float w[KK][DV], gw[KK][DV], gb[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) {
        gb[k] = 0.f;
#pragma unroll
        for (int j = 0; j < DV; ++j) { const int d = lane + 32 * j; w[k][j] = d < D ? Wh[k * D + d] : 0.f; gw[k][j] = 0.f; }
    }
    for (int64_t b = warp; b < B; b += nwarps) {
        const int64_t n = lens[b];
        float av[KK], ap[KK];
#pragma unroll
        for (int k = 0; k < KK; ++k) { av[k] = 0.f; ap[k] = 0.f; }
        for (int64_t l = lane; l < L; l += 32) {
            const bool valid = l < n;
            const float ge = d_ens ? d_ens[b * L + l] : 0.f;
#pragma unroll
            for (int k = 0; k < KK; ++k) {
                float g = ge * (float)scores[(b * L + l) * KK + k];
                if (d_weights) g += d_weights[(b * L + l) * KK + k];
                if (valid) av[k] += g; else ap[k] += g;
            }
        }
        float x[DV], dx[DV];
#pragma unroll
        for (int j = 0; j < DV; ++j) { const int d = lane + 32 * j; x[j] = d < D ? all[b * D + d] : 0.f; dx[j] = 0.f; }
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            const float sv = warp_sum(av[k]), sp = warp_sum(ap[k]);     // d(w_valid[b,k]), d(w_pad[b,k])
            gb[k] += sv + sp;
#pragma unroll
            for (int j = 0; j < DV; ++j) {
                const float g = (lane + 32 * j >= off_u) ? sv + sp : sv;   // pad rows see the inputs >= off_u only
                dx[j] = fmaf(g, w[k][j], dx[j]);
                gw[k][j] = fmaf(g, x[j], gw[k][j]);
            }
        }
#pragma unroll
        for (int j = 0; j < DV; ++j) { const int d = lane + 32 * j; if (d < D) dall[b * D + d] = dx[j]; }
    }
    // the CTA's share of the head gradients: warps -> shared memory -> one atomic per entry
#pragma unroll
    for (int k = 0; k < KK; ++k) {
#pragma unroll
        for (int j = 0; j < DV; ++j) red[wib][(k * DV + j) * 32 + lane] = gw[k][j];
        if (lane == 0) red[wib][KK * DV * 32 + k] = gb[k];            // every lane holds the same warp sums
    }
    __syncthreads();
    for (int e = threadIdx.x; e < KK * DV * 32 + KK; e += blockDim.x) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][e];
        if (e < KK * DV * 32) {
            const int k = e / (DV * 32), j = (e / 32) % DV, d = (e & 31) + 32 * j;
            if (d < D) atomicAdd(gWh + k * D + d, t);
        } else {
            atomicAdd(gbh + (e - KK * DV * 32), t);
        }
    }
}

#define HEAD_DISPATCH(CALL)                                         \
    do {                                                            \
        const int dv = (D + 31) / 32;                               \
        if (K == 2 && dv <= 4) { CALL(2, 4); }                      \
        else if (K == 2 && dv <= 8) { CALL(2, 8); }                 \
        else if (K == 3 && dv <= 4) { CALL(3, 4); }                 \
        else if (K == 3 && dv <= 8) { CALL(3, 8); }                 \
        else if (K == 4 && dv <= 4) { CALL(4, 4); }                 \
        else if (K == 4 && dv <= 8) { CALL(4, 8); }                 \
        else return INTEL_ERR_UNSUPPORTED;                          \
    } while (0)

bool head_full_ok(int K, int D) { return (K >= 2 && K <= 4) && D >= 1 && D <= 256; }

int head_full_fwd(int64_t B, int64_t L, int K, int D, int off_u, const float* all, const float* Wh, const float* bh,
                  const double* scores, const int64_t* lens, float* weights, float* ens, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(B, 8), 4);
#define HEAD_FWD(KK, DV) LAUNCH((head_full_fwd_kernel<KK, DV>), dim3(grid), dim3(256), 0, s, B, L, D, off_u, all, Wh, bh, scores, lens, weights, ens)
    HEAD_DISPATCH(HEAD_FWD);
#undef HEAD_FWD
    return check_launch("head_fuse_fwd", (double)B * (L * (8.0 * K + 4.0 * K + 4.0) + 4.0 * D), 2.0 * B * (L * K + 2.0 * D * K));
}

int head_full_bwd(int64_t B, int64_t L, int K, int D, int off_u, const float* all, const float* Wh, const float* d_weights,
                  const float* d_ens, const double* scores, const int64_t* lens, float* dall, float* gWh, float* gbh, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(B, 8), 2);
#define HEAD_BWD(KK, DV) LAUNCH((head_full_bwd_kernel<KK, DV>), dim3(grid), dim3(256), 0, s, B, L, D, off_u, all, Wh, d_weights, d_ens, scores, lens, dall, gWh, gbh)
    HEAD_DISPATCH(HEAD_BWD);
#undef HEAD_BWD
    return check_launch("head_fuse_bwd", (double)B * (L * (8.0 * K + 4.0 * K + 4.0) + 8.0 * D), 2.0 * B * (L * K + 4.0 * D * K));
}

// ---- per-item fusion: ens[r] = sum_k weights[r,k] * float(scores[r,k]) (fixed-weight baselines and
//      the cross_attention=0 branch) ----
__global__ void __launch_bounds__(256) item_fuse_fwd_kernel(int64_t R, int K, const float* __restrict__ weights,
                                                            const double* __restrict__ scores, float* __restrict__ ens) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(weights[r * K + k], (float)scores[r * K + k], acc);
        ens[r] = acc;
    }
}
int item_fuse_fwd(int64_t R, int K, const float* weights, const double* scores, float* ens, cudaStream_t s) {
    if (R <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(R, 256), 8);
    LAUNCH(item_fuse_fwd_kernel, dim3(grid), dim3(256), 0, s, R, K, weights, scores, ens);
    return check_launch("item_fuse_fwd");
}

__global__ void __launch_bounds__(256) item_fuse_bwd_kernel(int64_t R, int K, const float* __restrict__ d_weights,
                                                            const float* __restrict__ d_ens,
                                                            const double* __restrict__ scores, float* __restrict__ g) {
    const int64_t total = R * K;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        float v = d_ens ? d_ens[e / K] * (float)scores[e] : 0.f;
        if (d_weights) v += d_weights[e];
        g[e] = v;
    }
}
int item_fuse_bwd(int64_t R, int K, const float* d_weights, const float* d_ens, const double* scores, float* g,
                  cudaStream_t s) {
    if (R <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(R * K, 256), 8);
    LAUNCH(item_fuse_bwd_kernel, dim3(grid), dim3(256), 0, s, R, K, d_weights, d_ens, scores, g);
    return check_launch("item_fuse_bwd");
}

// ---- gate (cross_attention = 0, IntEL.py:205-209): Y[b,l,:] = X[b,l,:] * m[b,:] ----
__global__ void __launch_bounds__(256) gate_fwd_kernel(int64_t B, int64_t L, int d, const float* __restrict__ X,
                                                       const float* __restrict__ m, float* __restrict__ Y, int64_t ldy) {
    const int64_t total = B * L * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % d);
        const int64_t row = e / d;
        Y[row * ldy + c] = X[e] * m[(row / L) * d + c];
    }
}
int gate_fwd(int64_t B, int64_t L, int d, const float* X, const float* m, float* Y, int64_t ldy, cudaStream_t s) {
    if (B * L <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(B * L * d, 256), 8);
    LAUNCH(gate_fwd_kernel, dim3(grid), dim3(256), 0, s, B, L, d, X, m, Y, ldy);
    return check_launch("gate_fwd");
}

// dX[b,l,:] = dY * m[b];  dm[b,:] = sum_l dY[b,l,:] * X[b,l,:]   (one warp per session)
__global__ void __launch_bounds__(256) gate_bwd_kernel(int64_t B, int64_t L, int d, const float* __restrict__ X,
                                                       const float* __restrict__ m, const float* __restrict__ dY,
                                                       int64_t lddy, float* __restrict__ dX, float* __restrict__ dm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp; b < B; b += nwarps) {
        for (int c = lane; c < d; c += 32) {
            const float mc = m[b * d + c];
            float acc = 0.f;
            for (int64_t l = 0; l < L; ++l) {
                const float g = dY[(b * L + l) * lddy + c];
                acc = fmaf(g, X[(b * L + l) * d + c], acc);
                dX[(b * L + l) * d + c] = g * mc;
            }
            dm[b * d + c] = acc;
        }
    }
}
int gate_bwd(int64_t B, int64_t L, int d, const float* X, const float* m, const float* dY, int64_t lddy, float* dX,
             float* dm, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(B, 8), 8);
    LAUNCH(gate_bwd_kernel, dim3(grid), dim3(256), 0, s, B, L, d, X, m, dY, lddy, dX, dm);
    return check_launch("gate_bwd");
}

// ---- broadcast a per-session vector over the list slots (h_u / h_intent repeat, IntEL.py:178,212) ----
__global__ void __launch_bounds__(256) bcast_rows_kernel(int64_t B, int64_t L, int d, const float* __restrict__ v,
                                                         int64_t ldv, float* __restrict__ out, int64_t ldo) {
    const int64_t total = B * L * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % d);
        const int64_t row = e / d;
        out[row * ldo + c] = v[(row / L) * ldv + c];
    }
}
int bcast_rows(int64_t B, int64_t L, int d, const float* v, int64_t ldv, float* out, int64_t ldo, cudaStream_t s) {
    if (B * L <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(B * L * d, 256), 8);
    LAUNCH(bcast_rows_kernel, dim3(grid), dim3(256), 0, s, B, L, d, v, ldv, out, ldo);
    return check_launch("bcast_rows");
}

__global__ void __launch_bounds__(256) bcast_rows_bwd_kernel(int64_t B, int64_t L, int d, const float* __restrict__ dout,
                                                             int64_t ldo, float* dv, int64_t ldv, int accumulate) {
    const int64_t total = B * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % d);
        const int64_t b = e / d;
        float acc = 0.f;
        for (int64_t l = 0; l < L; ++l) acc += dout[(b * L + l) * ldo + c];
        if (accumulate) dv[b * ldv + c] += acc; else dv[b * ldv + c] = acc;
    }
}
int bcast_rows_bwd(int64_t B, int64_t L, int d, const float* dout, int64_t ldo, float* dv, int64_t ldv, int accumulate,
                   cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(B * d, 256), 8);
    LAUNCH(bcast_rows_bwd_kernel, dim3(grid), dim3(256), 0, s, B, L, d, dout, ldo, dv, ldv, accumulate);
    return check_launch("bcast_rows_bwd");
}

// dx[r, c] = dy[r, c] * (v[r, c] > 0) on strided [rows, cols] views (dx may alias dy)
__global__ void __launch_bounds__(256) relu_bwd_kernel(int64_t rows, int cols, const float* dy, int64_t lddy,
                                                       const float* __restrict__ v, int64_t ldv, float* dx, int64_t lddx) {
    const int64_t n = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cols;
        const int c = (int)(e % cols);
        dx[r * lddx + c] = (v[r * ldv + c] > 0.f) ? dy[r * lddy + c] : 0.f;
    }
}
int relu_bwd(int64_t rows, int cols, const float* dy, int64_t lddy, const float* v, int64_t ldv, float* dx, int64_t lddx,
             cudaStream_t s) {
    if (rows * cols <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(rows * cols, 256), 8);
    LAUNCH(relu_bwd_kernel, dim3(grid), dim3(256), 0, s, rows, cols, dy, lddy, v, ldv, dx, lddx);
    return check_launch("relu_bwd");
}

// y[r, c] = relu(x[r, c] + b[c]) on strided [rows, cols] views (the intent_embeddings column block of the merged
// intent projection, IntEL.py:212)
__global__ void __launch_bounds__(256) bias_relu_kernel(int64_t rows, int cols, const float* __restrict__ x, int64_t ldx,
                                                        const float* __restrict__ b, float* __restrict__ y, int64_t ldy) {
    const int64_t n = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cols;
        const int c = (int)(e % cols);
        y[r * ldy + c] = fmaxf(x[r * ldx + c] + (b ? b[c] : 0.f), 0.f);
    }
}
int bias_relu_rows(int64_t rows, int cols, const float* x, int64_t ldx, const float* b, float* y, int64_t ldy, cudaStream_t s) {
    if (rows * cols <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(rows * cols, 256), 8);
    LAUNCH(bias_relu_kernel, dim3(grid), dim3(256), 0, s, rows, cols, x, ldx, b, y, ldy);
    return check_launch("bias_relu");
}

// staged path of the stacks: y[r,c] = x[r,c] * dropout_scale(r,c) (+ add[r,c]); y may alias x
__global__ void __launch_bounds__(256) dropout_kernel(int64_t rows, int width, const float* x, const float* __restrict__ add,
                                                      float* y, Dropout dr) {
    const int64_t n = rows * width;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[e] * dropout_scale(dr, e / width, (int)(e % width), width);
        y[e] = add ? v + add[e] : v;
    }
}
int dropout_apply(int64_t rows, int width, const float* x, const float* add, float* y, const Dropout& dr, cudaStream_t s) {
    if (rows * width <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(rows * width, 256), 8);
    LAUNCH(dropout_kernel, dim3(grid), dim3(256), 0, s, rows, width, x, add, y, dr);
    return check_launch("dropout", 12.0 * rows * width, 0.0);
}

__global__ void __launch_bounds__(256) add_inplace_kernel(int64_t n, float* y, const float* __restrict__ x) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        y[e] += x[e];
}
int add_inplace(int64_t n, float* y, const float* x, cudaStream_t s) {
    if (n <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(n, 256), 8);
    LAUNCH(add_inplace_kernel, dim3(grid), dim3(256), 0, s, n, y, x);
    return check_launch("add_inplace");
}

}  // namespace intel

// ------------------------------------------------------------------------------------------------
// aWELv (models/supervise/aWELv.py:28-39): per-user fusion weights w = softmax_k <U[u], M[k]>, the same for every
// slot of the list; ens = sum_k w_k score_k.  One warp per session, forward and backward.
namespace intel {

static const int AW_MAX_K = 16;

__global__ void __launch_bounds__(256) awelv_fwd_kernel(int64_t B, int64_t L, int K, int h, const float* __restrict__ U,
                                                        const float* __restrict__ M, const int64_t* __restrict__ uid,
                                                        const double* __restrict__ scores, float* __restrict__ weights,
                                                        float* __restrict__ ens, float* __restrict__ wsmall) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const float* u = U + uid[b] * h;
    float w[AW_MAX_K];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < AW_MAX_K; ++k) {
        w[k] = 0.f;
        if (k < K) {
            float a = 0.f;
            for (int c = lane; c < h; c += 32) a = fmaf(u[c], M[k * h + c], a);
            w[k] = warp_sum(a);
            mx = fmaxf(mx, w[k]);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < AW_MAX_K; ++k)
        if (k < K) { w[k] = expf(w[k] - mx); sum += w[k]; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int k = 0; k < AW_MAX_K; ++k)
        if (k < K) {
            w[k] *= inv;
            if (lane == 0) wsmall[b * K + k] = w[k];
        }
    for (int64_t l = lane; l < L; l += 32) {
        float e = 0.f;
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) {
                weights[(b * L + l) * K + k] = w[k];
                e = __fadd_rn(e, __fmul_rn(w[k], (float)scores[(b * L + l) * K + k]));    // torch.mul(w, s).sum(dim=2)
            }
        ens[b * L + l] = e;
    }
}

__global__ void __launch_bounds__(256) awelv_bwd_kernel(int64_t B, int64_t L, int K, int h, const float* __restrict__ U,
                                                        const float* __restrict__ M, const int64_t* __restrict__ uid,
                                                        const double* __restrict__ scores, const float* __restrict__ wsmall,
                                                        const float* __restrict__ d_weights, const float* __restrict__ d_ens,
                                                        float* gU, float* gM) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    float dw[AW_MAX_K], w[AW_MAX_K];
#pragma unroll
    for (int k = 0; k < AW_MAX_K; ++k) { dw[k] = 0.f; w[k] = (k < K) ? wsmall[b * K + k] : 0.f; }
    for (int64_t l = lane; l < L; l += 32) {
        const float de = d_ens ? d_ens[b * L + l] : 0.f;
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) {
                float g = de * (float)scores[(b * L + l) * K + k];
                if (d_weights) g += d_weights[(b * L + l) * K + k];
                dw[k] += g;
            }
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < AW_MAX_K; ++k)
        if (k < K) { dw[k] = warp_sum(dw[k]); dot = fmaf(dw[k], w[k], dot); }
    const int64_t u = uid[b];
    for (int c = lane; c < h; c += 32) {
        const float uc = U[u * h + c];
        float gu = 0.f;
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) {
                const float dl = w[k] * (dw[k] - dot);             // softmax backward
                gu = fmaf(dl, M[k * h + c], gu);
                atomicAdd(gM + k * h + c, dl * uc);
            }
        atomicAdd(gU + u * h + c, gu);
    }
}

int awelv_fwd(int64_t B, int64_t L, int K, int h, const float* U, const float* M, const int64_t* uid, const double* scores,
              float* weights, float* ens, float* wsmall, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    INTEL_REQUIRE(K >= 1 && K <= AW_MAX_K && h >= 1, INTEL_ERR_UNSUPPORTED, "aWELv: model_num %d > %d", K, AW_MAX_K);
    LAUNCH(awelv_fwd_kernel, dim3((unsigned)ceil_div(B, 8)), dim3(256), 0, s, B, L, K, h, U, M, uid, scores, weights, ens, wsmall);
    return check_launch("awelv_fwd", (double)B * L * K * 12.0 + (double)B * L * 4.0, 2.0 * B * L * K);
}

int awelv_bwd(int64_t B, int64_t L, int K, int h, const float* U, const float* M, const int64_t* uid, const double* scores,
              const float* wsmall, const float* d_weights, const float* d_ens, float* gU, float* gM, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    INTEL_REQUIRE(K >= 1 && K <= AW_MAX_K && h >= 1, INTEL_ERR_UNSUPPORTED, "aWELv: model_num %d > %d", K, AW_MAX_K);
    LAUNCH(awelv_bwd_kernel, dim3((unsigned)ceil_div(B, 8)), dim3(256), 0, s, B, L, K, h, U, M, uid, scores, wsmall, d_weights,
           d_ens, gU, gM);
    return check_launch("awelv_bwd", (double)B * L * K * 12.0 + (double)B * L * 4.0, 4.0 * B * L * K);
}

// ------------------------------------------------------------------------------------------------
// aWELv_IntEL head (models/supervise/aWELv_IntEL.py:190-201): the weight head reads the MEAN over all L list slots (pads
// included) of the gated streams.  The head is affine, so its output equals the mean over the slots of the per-slot
// head output of the cross_attention = 0 path (Wl [B,L,K], intel_ensemble_fwd): logits = mean_l Wl[b,l,:],
// p = softmax(logits), w = softmax(p) (the reference applies softmax twice, :197-198), weights[b,l,:] = w,
// ens[b,l] = sum_k w_k score_k.  One warp per session; p and w [B,K] are kept for the backward pass, which returns
// dWl[b,l,k] = dlogits_k / L.
__global__ void __launch_bounds__(256) pool_head_fwd_kernel(int64_t B, int64_t L, int K, const float* __restrict__ Wl,
                                                            const double* __restrict__ scores, float* __restrict__ weights,
                                                            float* __restrict__ ens, float* __restrict__ p_out,
                                                            float* __restrict__ w_out) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    float w[AW_MAX_K];
#pragma unroll
    for (int k = 0; k < AW_MAX_K; ++k) w[k] = 0.f;
    for (int64_t l = lane; l < L; l += 32)
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) w[k] += Wl[(b * L + l) * K + k];
    const float inv_l = 1.0f / (float)L;
    for (int pass = 0; pass < 2; ++pass) {          // pass 0: p = softmax(mean), pass 1: w = softmax(p)
        float mx = -INFINITY, sum = 0.f;
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) {
                if (pass == 0) w[k] = warp_sum(w[k]) * inv_l;
                mx = fmaxf(mx, w[k]);
            }
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) { w[k] = expf(w[k] - mx); sum += w[k]; }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) {
                w[k] *= inv;
                if (lane == 0) (pass == 0 ? p_out : w_out)[b * K + k] = w[k];
            }
    }
    for (int64_t l = lane; l < L; l += 32) {
        float e = 0.f;
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) {
                weights[(b * L + l) * K + k] = w[k];
                e = __fadd_rn(e, __fmul_rn(w[k], (float)scores[(b * L + l) * K + k]));
            }
        ens[b * L + l] = e;
    }
}

__global__ void __launch_bounds__(256) pool_head_bwd_kernel(int64_t B, int64_t L, int K, const double* __restrict__ scores,
                                                            const float* __restrict__ p_in, const float* __restrict__ w_in,
                                                            const float* __restrict__ d_weights, const float* __restrict__ d_ens,
                                                            float* __restrict__ dWl) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    float g[AW_MAX_K];
#pragma unroll
    for (int k = 0; k < AW_MAX_K; ++k) g[k] = 0.f;
    for (int64_t l = lane; l < L; l += 32) {
        const float de = d_ens ? d_ens[b * L + l] : 0.f;
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) {
                float t = de * (float)scores[(b * L + l) * K + k];
                if (d_weights) t += d_weights[(b * L + l) * K + k];
                g[k] += t;
            }
    }
#pragma unroll
    for (int k = 0; k < AW_MAX_K; ++k)
        if (k < K) g[k] = warp_sum(g[k]);
    for (int pass = 0; pass < 2; ++pass) {          // back through w = softmax(p), then p = softmax(logits)
        const float* y = pass == 0 ? w_in : p_in;
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) dot = fmaf(g[k], y[b * K + k], dot);
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) g[k] = y[b * K + k] * (g[k] - dot);
    }
    const float inv_l = 1.0f / (float)L;
    for (int64_t l = lane; l < L; l += 32)
#pragma unroll
        for (int k = 0; k < AW_MAX_K; ++k)
            if (k < K) dWl[(b * L + l) * K + k] = g[k] * inv_l;
}

int pool_head_fwd(int64_t B, int64_t L, int K, const float* Wl, const double* scores, float* weights, float* ens, float* p,
                  float* w, cudaStream_t s) {
    if (B <= 0 || L <= 0) return INTEL_OK;
    INTEL_REQUIRE(K >= 1 && K <= AW_MAX_K, INTEL_ERR_UNSUPPORTED, "aWELv_IntEL: model_num %d > %d", K, AW_MAX_K);
    INTEL_REQUIRE(Wl && scores && weights && ens && p && w, INTEL_ERR_ARG, "pool_head_fwd: null pointer");
    LAUNCH(pool_head_fwd_kernel, dim3((unsigned)ceil_div(B, 8)), dim3(256), 0, s, B, L, K, Wl, scores, weights, ens, p, w);
    return check_launch("pool_head_fwd", (double)B * L * K * 16.0 + (double)B * L * 4.0, 2.0 * B * L * K);
}

int pool_head_bwd(int64_t B, int64_t L, int K, const double* scores, const float* p, const float* w, const float* d_weights,
                  const float* d_ens, float* dWl, cudaStream_t s) {
    if (B <= 0 || L <= 0) return INTEL_OK;
    INTEL_REQUIRE(K >= 1 && K <= AW_MAX_K, INTEL_ERR_UNSUPPORTED, "aWELv_IntEL: model_num %d > %d", K, AW_MAX_K);
    INTEL_REQUIRE(scores && p && w && dWl, INTEL_ERR_ARG, "pool_head_bwd: null pointer");
    LAUNCH(pool_head_bwd_kernel, dim3((unsigned)ceil_div(B, 8)), dim3(256), 0, s, B, L, K, scores, p, w, d_weights, d_ens, dWl);
    return check_launch("pool_head_bwd", (double)B * L * K * 16.0 + (double)B * L * 4.0, 4.0 * B * L * K);
}

}  // namespace intel
