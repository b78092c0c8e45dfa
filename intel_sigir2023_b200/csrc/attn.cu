// Per-session attention kernels: LayerNorm, multi-head self-attention (fwd + bwd, recompute-based),
// pooled cross-attention, row softmax.  A whole session (L <= a few hundred tokens of 32 floats)
// fits one SM's shared memory, so every (session, head) pair is one CTA and nothing O(L^2) ever
// reaches HBM (the reference materialises [B,h,L,L] and [B,L,L] tensors, layers.py:54, attention.py:57).
#include "kernels.h"

namespace intel {

static const int LN_MAX_PER_LANE = 8;   // d <= 256

// ------------------------------------------------------------------------------------------------
// LayerNorm forward / backward: one warp per row, P = ceil(d / 32) values per lane (compile time), RB rows in flight
// per warp so that the two dependent warp reductions of a row overlap with those of its neighbours.
static const int LN_RB = 4;

template <int P>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(int64_t R, int d, const float* __restrict__ Z,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ Y,
                                                            float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float gm[P], bt[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int c = lane + 32 * i;
        gm[i] = (c < d) ? gamma[c] : 0.f;
        bt[i] = (c < d) ? beta[c] : 0.f;
    }
    for (int64_t r0 = warp * LN_RB; r0 < R; r0 += nwarps * LN_RB) {
        float v[LN_RB][P], sum[LN_RB], sq[LN_RB];
#pragma unroll
        for (int q = 0; q < LN_RB; ++q) {
            sum[q] = 0.f;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const int c = lane + 32 * i;
                v[q][i] = (c < d && r0 + q < R) ? Z[(r0 + q) * d + c] : 0.f;
                sum[q] += v[q][i];
            }
        }
#pragma unroll
        for (int q = 0; q < LN_RB; ++q) sum[q] = warp_sum(sum[q]) / (float)d;
#pragma unroll
        for (int q = 0; q < LN_RB; ++q) {
            sq[q] = 0.f;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const float t = (lane + 32 * i < d) ? v[q][i] - sum[q] : 0.f;
                sq[q] += t * t;
            }
        }
#pragma unroll
        for (int q = 0; q < LN_RB; ++q) sq[q] = rsqrtf(warp_sum(sq[q]) / (float)d + 1e-5f);
#pragma unroll
        for (int q = 0; q < LN_RB; ++q) {
            if (r0 + q < R) {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int c = lane + 32 * i;
                    if (c < d) Y[(r0 + q) * d + c] = (v[q][i] - sum[q]) * sq[q] * gm[i] + bt[i];
                }
                if (lane == 0) { stats[2 * (r0 + q)] = sum[q]; stats[2 * (r0 + q) + 1] = sq[q]; }
            }
        }
    }
}

int layernorm_fwd(int64_t R, int d, const float* Z, const float* gamma, const float* beta, float* Y, float* stats,
                  cudaStream_t s) {
    if (R <= 0) return INTEL_OK;
    INTEL_REQUIRE(d <= 32 * LN_MAX_PER_LANE, INTEL_ERR_UNSUPPORTED, "layernorm width %d > 256", d);
    unsigned grid = stream_grid(ceil_div(R, 8 * LN_RB), 8);
    if (d <= 32) LAUNCH(layernorm_fwd_kernel<1>, dim3(grid), dim3(256), 0, s, R, d, Z, gamma, beta, Y, stats);
    else if (d <= 64) LAUNCH(layernorm_fwd_kernel<2>, dim3(grid), dim3(256), 0, s, R, d, Z, gamma, beta, Y, stats);
    else if (d <= 128) LAUNCH(layernorm_fwd_kernel<4>, dim3(grid), dim3(256), 0, s, R, d, Z, gamma, beta, Y, stats);
    else LAUNCH(layernorm_fwd_kernel<8>, dim3(grid), dim3(256), 0, s, R, d, Z, gamma, beta, Y, stats);
    return check_launch("layernorm_fwd", 8.0 * R * d, 8.0 * R * d);
}

template <int P>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(int64_t R, int d, const float* __restrict__ dY,
                                                            const float* __restrict__ Z, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, float* __restrict__ dZ,
                                                            float* dgamma, float* dbeta) {
    __shared__ float red_g[8][32 * P];
    __shared__ float red_b[8][32 * P];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float pg[P], pb[P], gm[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        pg[i] = 0.f;
        pb[i] = 0.f;
        gm[i] = (lane + 32 * i < d) ? gamma[lane + 32 * i] : 0.f;
    }
    for (int64_t r0 = warp * LN_RB; r0 < R; r0 += nwarps * LN_RB) {
        float xh[LN_RB][P], g[LN_RB][P], s1[LN_RB], s2[LN_RB], rstd[LN_RB];
#pragma unroll
        for (int q = 0; q < LN_RB; ++q) {
            const bool ok = r0 + q < R;
            const float mean = ok ? stats[2 * (r0 + q)] : 0.f;
            rstd[q] = ok ? stats[2 * (r0 + q) + 1] : 0.f;
            s1[q] = 0.f;
            s2[q] = 0.f;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const int c = lane + 32 * i;
                if (ok && c < d) {
                    const float dy = dY[(r0 + q) * d + c];
                    xh[q][i] = (Z[(r0 + q) * d + c] - mean) * rstd[q];
                    g[q][i] = dy * gm[i];
                    pg[i] += dy * xh[q][i];
                    pb[i] += dy;
                } else { xh[q][i] = 0.f; g[q][i] = 0.f; }
                s1[q] += g[q][i];
                s2[q] += g[q][i] * xh[q][i];
            }
        }
#pragma unroll
        for (int q = 0; q < LN_RB; ++q) { s1[q] = warp_sum(s1[q]) / (float)d; s2[q] = warp_sum(s2[q]) / (float)d; }
#pragma unroll
        for (int q = 0; q < LN_RB; ++q) {
            if (r0 + q < R) {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int c = lane + 32 * i;
                    if (c < d) dZ[(r0 + q) * d + c] = rstd[q] * (g[q][i] - s1[q] - xh[q][i] * s2[q]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < P; ++i) { red_g[wib][lane + 32 * i] = pg[i]; red_b[wib][lane + 32 * i] = pb[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { tg += red_g[w][c]; tb += red_b[w][c]; }
        atomicAdd(dgamma + c, tg);
        atomicAdd(dbeta + c, tb);
    }
}

int layernorm_bwd(int64_t R, int d, const float* dY, const float* Z, const float* stats, const float* gamma, float* dZ,
                  float* dgamma, float* dbeta, cudaStream_t s) {
    if (R <= 0) return INTEL_OK;
    INTEL_REQUIRE(d <= 32 * LN_MAX_PER_LANE, INTEL_ERR_UNSUPPORTED, "layernorm width %d > 256", d);
    unsigned grid = stream_grid(ceil_div(R, 8 * LN_RB), 4);
    if (d <= 32) LAUNCH(layernorm_bwd_kernel<1>, dim3(grid), dim3(256), 0, s, R, d, dY, Z, stats, gamma, dZ, dgamma, dbeta);
    else if (d <= 64) LAUNCH(layernorm_bwd_kernel<2>, dim3(grid), dim3(256), 0, s, R, d, dY, Z, stats, gamma, dZ, dgamma, dbeta);
    else if (d <= 128) LAUNCH(layernorm_bwd_kernel<4>, dim3(grid), dim3(256), 0, s, R, d, dY, Z, stats, gamma, dZ, dgamma, dbeta);
    else LAUNCH(layernorm_bwd_kernel<8>, dim3(grid), dim3(256), 0, s, R, d, dY, Z, stats, gamma, dZ, dgamma, dbeta);
    return check_launch("layernorm_bwd", 12.0 * R * d, 12.0 * R * d);
}

// ------------------------------------------------------------------------------------------------
// Pooled cross attention: one warp per session.  The reference broadcasts one [1,L] attention row
// against the [L,L] pair mask (attention.py:57-60 via IntEL.py:201-204), so all valid rows of a
// session share one pooled vector; pad rows get 0.  qk = W_k^T q and the value projection are applied
// outside by GEMM, this kernel only sees X [B,L,d] and qk [B,d].
static const int XP_WARPS = 4;

__global__ void __launch_bounds__(XP_WARPS * 32) cross_pool_fwd_kernel(int64_t B, int64_t L, int d,
                                                                       const float* __restrict__ X,
                                                                       const float* __restrict__ qk,
                                                                       const int64_t* __restrict__ lens, float scale,
                                                                       float* __restrict__ p, float* __restrict__ xbar) {
    DYN_SMEM(float, sm);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * XP_WARPS + w;
    float* pb = sm + w * L;
    if (b >= B) return;
    int64_t n = lens[b];
    if (n > L) n = L;
    const float* x = X + b * L * d;
    const float* q = qk + b * d;
    float mx = -INFINITY;
    if (d == 32 || d == 64) {
        // d / 4 lanes per row, one 16-byte load each: a warp reads 512 contiguous bytes per pass
        const int lpr = d >> 2, rpp = 32 / lpr;                 // lanes per row, rows per pass
        const int sub = lane % lpr, rr = lane / lpr;
        const float4 qv = *reinterpret_cast<const float4*>(q + 4 * sub);
        for (int64_t j0 = 0; j0 < L; j0 += rpp) {
            const int64_t j = j0 + rr;
            float a = 0.f;
            if (j < L) {
                const float4 xv = *reinterpret_cast<const float4*>(x + j * d + 4 * sub);
                a = fmaf(xv.x, qv.x, fmaf(xv.y, qv.y, fmaf(xv.z, qv.z, xv.w * qv.w)));
            }
            for (int o = lpr >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            a *= scale;
            if (j < L) {
                if (sub == 0) pb[j] = a;
                if (j < n) mx = fmaxf(mx, a);
            }
        }
        __syncwarp();
    } else {
        for (int64_t j = lane; j < L; j += 32) {
            float a = 0.f;
            for (int c = 0; c < d; ++c) a = fmaf(x[j * d + c], q[c], a);
            a *= scale;
            pb[j] = a;
            if (j < n) mx = fmaxf(mx, a);
        }
    }
    // attention.py:57-60 subtracts the max over ALL slots, masks the pads to -inf and calls softmax, which shifts by the
    // largest VALID logit once more: that second shift is the one that decides the result, so it is the one used here
    // (a pad logit far above every valid one must not underflow the whole row to zero)
    mx = warp_max(mx);
    float sum = 0.f;
    for (int64_t j = lane; j < L; j += 32) {
        const float e = (j < n) ? expf(pb[j] - mx) : 0.f;
        pb[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    const float inv = (sum > 0.f) ? 1.0f / sum : 0.f;
    for (int64_t j = lane; j < L; j += 32) { pb[j] *= inv; p[b * L + j] = pb[j]; }
    __syncwarp();
    for (int c = lane; c < d; c += 32) {
        float acc = 0.f;
        for (int64_t j = 0; j < n; ++j) acc = fmaf(pb[j], x[j * d + c], acc);
        xbar[b * d + c] = acc;
    }
}

int cross_pool_fwd(int64_t B, int64_t L, int d, const float* X, const float* qk, const int64_t* lens, float scale,
                   float* p, float* xbar, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    const size_t smem = (size_t)XP_WARPS * L * 4;
    INTEL_REQUIRE(smem <= 48 * 1024, INTEL_ERR_UNSUPPORTED, "cross_pool: list length %lld too long", (long long)L);
    LAUNCH(cross_pool_fwd_kernel, dim3((unsigned)ceil_div(B, XP_WARPS)), dim3(XP_WARPS * 32), smem, s, B, L, d, X, qk,
           lens, scale, p, xbar);
    return check_launch("cross_pool_fwd", 4.0 * B * L * (d + 1), 4.0 * B * L * d);
}

__global__ void __launch_bounds__(XP_WARPS * 32) cross_pool_bwd_kernel(int64_t B, int64_t L, int d,
                                                                       const float* __restrict__ X,
                                                                       const float* __restrict__ qk,
                                                                       const int64_t* __restrict__ lens, float scale,
                                                                       const float* __restrict__ p,
                                                                       const float* __restrict__ dxbar,
                                                                       float* __restrict__ dX, float* __restrict__ dqk) {
    DYN_SMEM(float, sm);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * XP_WARPS + w;
    float* ab = sm + w * L;       // datt
    if (b >= B) return;
    int64_t n = lens[b];
    if (n > L) n = L;
    const float* x = X + b * L * d;
    const float* g = dxbar + b * d;
    float delta = 0.f;
    if (d == 32 || d == 64) {
        const int lpr = d >> 2, rpp = 32 / lpr;
        const int sub = lane % lpr, rr = lane / lpr;
        const float4 gv = *reinterpret_cast<const float4*>(g + 4 * sub);
        for (int64_t j0 = 0; j0 < n; j0 += rpp) {
            const int64_t j = j0 + rr;
            float a = 0.f;
            if (j < n) {
                const float4 xv = *reinterpret_cast<const float4*>(x + j * d + 4 * sub);
                a = fmaf(xv.x, gv.x, fmaf(xv.y, gv.y, fmaf(xv.z, gv.z, xv.w * gv.w)));
            }
            for (int o = lpr >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (j < n && sub == 0) {
                ab[j] = a;                            // dp_j
                delta = fmaf(p[b * L + j], a, delta);
            }
        }
        __syncwarp();
    } else {
        for (int64_t j = lane; j < n; j += 32) {
            float a = 0.f;
            for (int c = 0; c < d; ++c) a = fmaf(x[j * d + c], g[c], a);
            ab[j] = a;                                // dp_j
            delta = fmaf(p[b * L + j], a, delta);
        }
    }
    delta = warp_sum(delta);
    for (int64_t j = lane; j < n; j += 32) ab[j] = p[b * L + j] * (ab[j] - delta) * scale;   // datt_j * scale
    __syncwarp();
    for (int c = lane; c < d; c += 32) {
        const float gc = g[c], qc = qk[b * d + c];
        float acc = 0.f;
        for (int64_t j = 0; j < L; ++j) {
            float v = 0.f;
            if (j < n) {
                v = fmaf(p[b * L + j], gc, ab[j] * qc);
                acc = fmaf(ab[j], x[j * d + c], acc);
            }
            dX[(b * L + j) * d + c] = v;
        }
        dqk[b * d + c] = acc;
    }
}

int cross_pool_bwd(int64_t B, int64_t L, int d, const float* X, const float* qk, const int64_t* lens, float scale,
                   const float* p, const float* dxbar, float* dX, float* dqk, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    const size_t smem = (size_t)XP_WARPS * L * 4;
    INTEL_REQUIRE(smem <= 48 * 1024, INTEL_ERR_UNSUPPORTED, "cross_pool: list length %lld too long", (long long)L);
    LAUNCH(cross_pool_bwd_kernel, dim3((unsigned)ceil_div(B, XP_WARPS)), dim3(XP_WARPS * 32), smem, s, B, L, d, X, qk,
           lens, scale, p, dxbar, dX, dqk);
    return check_launch("cross_pool_bwd", 4.0 * B * L * (2 * d + 1), 8.0 * B * L * d);
}

// ------------------------------------------------------------------------------------------------
// The pooled cross attention of a width-32 stream with its three 32 x 32 projections folded in (CrossAtt / MultiQueryAtt,
// attention.py:54-63, 149-161): qk = W_k^T q before the pooling, out = W_v xbar after it; the backward kernel returns
// d(q) and keeps d(W_k), d(W_v) in registers across the sessions of a warp (lane c owns column c of both).  In round 2 these
// were twelve [4096 x 32 x 32] GEMM launches per step around the two pooling kernels.  One warp per session, persistent CTAs;
// the 32-vector-by-matrix products broadcast the vector with shuffles, the matrices sit in shared memory ([32][33]).
static const int XF_WARPS = 4;

__global__ void __launch_bounds__(XF_WARPS * 32) cross_full_fwd_kernel(int64_t B, int64_t L, const float* __restrict__ X,
                                                                       const float* __restrict__ q, int64_t ldq,
                                                                       const float* __restrict__ Wk, const float* __restrict__ Wv,
                                                                       const int64_t* __restrict__ lens, float scale,
                                                                       float* __restrict__ p, float* __restrict__ qk_out,
                                                                       float* __restrict__ xbar_out, float* __restrict__ out, int64_t ldo) {
    DYN_SMEM(float, sm);
    constexpr int d = 32;
    float* Wk_s = sm;                        // [a][c] stride 33
    float* Wv_s = Wk_s + 32 * 33;            // [e][c]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* vec = Wv_s + 32 * 33 + w * 32;    // the warp's qk
    float* pb = Wv_s + 32 * 33 + XF_WARPS * 32 + w * L;
#pragma unroll
    for (int e = threadIdx.x; e < 32 * 32 / 4; e += XF_WARPS * 32) {      // 16-byte loads, all in flight at once
        const float4 k4 = *reinterpret_cast<const float4*>(Wk + 4 * e), v4 = *reinterpret_cast<const float4*>(Wv + 4 * e);
        float* dk = Wk_s + (e >> 3) * 33 + 4 * (e & 7);
        float* dv = Wv_s + (e >> 3) * 33 + 4 * (e & 7);
        dk[0] = k4.x; dk[1] = k4.y; dk[2] = k4.z; dk[3] = k4.w;
        dv[0] = v4.x; dv[1] = v4.y; dv[2] = v4.z; dv[3] = v4.w;
    }
    __syncthreads();
    const int64_t nwarps = (int64_t)gridDim.x * XF_WARPS;
    for (int64_t b = (int64_t)blockIdx.x * XF_WARPS + w; b < B; b += nwarps) {
        int64_t n = lens[b];
        if (n > L) n = L;
        const float* x = X + b * L * d;
        // qk[c] = sum_a q[a] W_k[a][c]
        const float qa = q[b * ldq + lane];
        float qkc = 0.f;
#pragma unroll
        for (int a = 0; a < 32; ++a) qkc = fmaf(__shfl_sync(0xffffffffu, qa, a), Wk_s[a * 33 + lane], qkc);
        qk_out[b * d + lane] = qkc;
        vec[lane] = qkc;
        __syncwarp();
        float mx = -INFINITY;
        {
            const int sub = lane & 7, rr = lane >> 3;           // 8 lanes per row, 4 rows per pass
            const float4 qv = *reinterpret_cast<const float4*>(vec + 4 * sub);
            for (int64_t j0 = 0; j0 < L; j0 += 4) {
                const int64_t j = j0 + rr;
                float a = 0.f;
                if (j < L) {
                    const float4 xv = *reinterpret_cast<const float4*>(x + j * d + 4 * sub);
                    a = fmaf(xv.x, qv.x, fmaf(xv.y, qv.y, fmaf(xv.z, qv.z, xv.w * qv.w)));
                }
                a += __shfl_xor_sync(0xffffffffu, a, 4);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                a += __shfl_xor_sync(0xffffffffu, a, 1);
                a *= scale;
                if (j < L) {
                    if (sub == 0) pb[j] = a;
                    if (j < n) mx = fmaxf(mx, a);
                }
            }
            __syncwarp();
        }
        mx = warp_max(mx);                   // shift by the largest VALID logit (see cross_pool_fwd_kernel)
        float sum = 0.f;
        for (int64_t j = lane; j < L; j += 32) {
            const float e = (j < n) ? expf(pb[j] - mx) : 0.f;
            pb[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = (sum > 0.f) ? 1.0f / sum : 0.f;
        for (int64_t j = lane; j < L; j += 32) { pb[j] *= inv; p[b * L + j] = pb[j]; }
        __syncwarp();
        float xb = 0.f;
        for (int64_t j = 0; j < n; ++j) xb = fmaf(pb[j], x[j * d + lane], xb);
        xbar_out[b * d + lane] = xb;
        // out[e] = sum_c W_v[e][c] xbar[c]
        float oe = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) oe = fmaf(__shfl_sync(0xffffffffu, xb, c), Wv_s[lane * 33 + c], oe);
        out[b * ldo + lane] = oe;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(XF_WARPS * 32) cross_full_bwd_kernel(int64_t B, int64_t L, const float* __restrict__ X,
                                                                       const float* __restrict__ q, int64_t ldq,
                                                                       const float* __restrict__ qk, const float* __restrict__ Wk,
                                                                       const float* __restrict__ Wv, const int64_t* __restrict__ lens,
                                                                       float scale, const float* __restrict__ p,
                                                                       const float* __restrict__ xbar, const float* __restrict__ dout,
                                                                       int64_t ldd, float* __restrict__ dX, float* __restrict__ dq_out,
                                                                       int64_t lddq, float* __restrict__ gWk, float* __restrict__ gWv) {
    DYN_SMEM(float, sm);
    constexpr int d = 32;
    float* Wk_s = sm;
    float* Wv_s = Wk_s + 32 * 33;
    float* gk_s = Wv_s + 32 * 33;            // the CTA's gradient sums [a][c], [e][c]
    float* gv_s = gk_s + 32 * 33;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* vec = gv_s + 32 * 33 + w * 32;    // the warp's d(xbar)
    float* ab = gv_s + 32 * 33 + XF_WARPS * 32 + w * L;
#pragma unroll
    for (int e = threadIdx.x; e < 32 * 32 / 4; e += XF_WARPS * 32) {
        const float4 k4 = *reinterpret_cast<const float4*>(Wk + 4 * e), v4 = *reinterpret_cast<const float4*>(Wv + 4 * e);
        const int o = (e >> 3) * 33 + 4 * (e & 7);
        Wk_s[o] = k4.x; Wk_s[o + 1] = k4.y; Wk_s[o + 2] = k4.z; Wk_s[o + 3] = k4.w;
        Wv_s[o] = v4.x; Wv_s[o + 1] = v4.y; Wv_s[o + 2] = v4.z; Wv_s[o + 3] = v4.w;
        gk_s[o] = 0.f; gk_s[o + 1] = 0.f; gk_s[o + 2] = 0.f; gk_s[o + 3] = 0.f;
        gv_s[o] = 0.f; gv_s[o + 1] = 0.f; gv_s[o + 2] = 0.f; gv_s[o + 3] = 0.f;
    }
    __syncthreads();
    float gk[32], gv[32];                    // lane c: d W_k[a][c], d W_v[e][c]
#pragma unroll
    for (int i = 0; i < 32; ++i) { gk[i] = 0.f; gv[i] = 0.f; }
    const int64_t nwarps = (int64_t)gridDim.x * XF_WARPS;
    for (int64_t b = (int64_t)blockIdx.x * XF_WARPS + w; b < B; b += nwarps) {
        int64_t n = lens[b];
        if (n > L) n = L;
        const float* x = X + b * L * d;
        const float ge = dout[b * ldd + lane];
        const float xbc = xbar[b * d + lane];
        // d(xbar)[c] = sum_e dout[e] W_v[e][c];  d W_v[e][c] += dout[e] xbar[c]
        float dxb = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const float g = __shfl_sync(0xffffffffu, ge, e);
            dxb = fmaf(g, Wv_s[e * 33 + lane], dxb);
            gv[e] = fmaf(g, xbc, gv[e]);
        }
        vec[lane] = dxb;
        __syncwarp();
        float delta = 0.f;
        {
            const int sub = lane & 7, rr = lane >> 3;
            const float4 g4 = *reinterpret_cast<const float4*>(vec + 4 * sub);
            for (int64_t j0 = 0; j0 < n; j0 += 4) {
                const int64_t j = j0 + rr;
                float a = 0.f;
                if (j < n) {
                    const float4 xv = *reinterpret_cast<const float4*>(x + j * d + 4 * sub);
                    a = fmaf(xv.x, g4.x, fmaf(xv.y, g4.y, fmaf(xv.z, g4.z, xv.w * g4.w)));
                }
                a += __shfl_xor_sync(0xffffffffu, a, 4);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                a += __shfl_xor_sync(0xffffffffu, a, 1);
                if (j < n && sub == 0) {
                    ab[j] = a;                            // dp_j
                    delta = fmaf(p[b * L + j], a, delta);
                }
            }
            __syncwarp();
        }
        delta = warp_sum(delta);
        for (int64_t j = lane; j < n; j += 32) ab[j] = p[b * L + j] * (ab[j] - delta) * scale;   // datt_j * scale
        __syncwarp();
        const float qc = qk[b * d + lane];
        float dqk = 0.f;
        for (int64_t j = 0; j < L; ++j) {
            float v = 0.f;
            if (j < n) {
                v = fmaf(p[b * L + j], dxb, ab[j] * qc);
                dqk = fmaf(ab[j], x[j * d + lane], dqk);
            }
            dX[(b * L + j) * d + lane] = v;
        }
        // d W_k[a][c] += q[a] dqk[c];  d(q)[a] = sum_c W_k[a][c] dqk[c]
        const float qa = q[b * ldq + lane];
        float dqa = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            gk[i] = fmaf(__shfl_sync(0xffffffffu, qa, i), dqk, gk[i]);
            dqa = fmaf(__shfl_sync(0xffffffffu, dqk, i), Wk_s[lane * 33 + i], dqa);
        }
        dq_out[b * lddq + lane] = dqa;
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        atomicAdd(&gk_s[i * 33 + lane], gk[i]);
        atomicAdd(&gv_s[i * 33 + lane], gv[i]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
        atomicAdd(gWk + e, gk_s[(e >> 5) * 33 + (e & 31)]);
        atomicAdd(gWv + e, gv_s[(e >> 5) * 33 + (e & 31)]);
    }
}

bool cross_full_ok(int d, int64_t L) { return d == 32 && L >= 1 && L <= 2048; }      // (the caller also checks the 16-byte alignment of W_k / W_v)

int cross_full_fwd(int64_t B, int64_t L, const float* X, const float* q, int64_t ldq, const float* Wk, const float* Wv,
                   const int64_t* lens, float scale, float* p, float* qk_out, float* xbar_out, float* out, int64_t ldo, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    const size_t smem = (size_t)(2 * 32 * 33 + XF_WARPS * 32 + XF_WARPS * L) * 4;
    ensure_smem(cross_full_fwd_kernel, smem);
    const unsigned grid = stream_grid(ceil_div(B, XF_WARPS), 8);
    LAUNCH(cross_full_fwd_kernel, dim3(grid), dim3(XF_WARPS * 32), smem, s, B, L, X, q, ldq, Wk, Wv, lens, scale, p, qk_out, xbar_out,
           out, ldo);
    return check_launch("cross_pool_fwd", 4.0 * B * L * (32 + 1), 4.0 * B * L * 32);
}

int cross_full_bwd(int64_t B, int64_t L, const float* X, const float* q, int64_t ldq, const float* qk, const float* Wk, const float* Wv,
                   const int64_t* lens, float scale, const float* p, const float* xbar, const float* dout, int64_t ldd, float* dX,
                   float* dq_out, int64_t lddq, float* gWk, float* gWv, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    const size_t smem = (size_t)(4 * 32 * 33 + XF_WARPS * 32 + XF_WARPS * L) * 4;
    ensure_smem(cross_full_bwd_kernel, smem);
    const unsigned grid = stream_grid(ceil_div(B, XF_WARPS), 4);      // measured per step (two launches): 2 / 4 / 8 CTAs per SM = 141 / 89 / 101 us
    LAUNCH(cross_full_bwd_kernel, dim3(grid), dim3(XF_WARPS * 32), smem, s, B, L, X, q, ldq, qk, Wk, Wv, lens, scale, p, xbar, dout, ldd,
           dX, dq_out, lddq, gWk, gWv);
    return check_launch("cross_pool_bwd", 4.0 * B * L * (2 * 32 + 1), 8.0 * B * L * 32);
}

// ------------------------------------------------------------------------------------------------
// Row softmax (pred_layer(...).softmax, IntEL.py:153) and its backward; one warp per row.
// ldz: row stride of Z (the logits workspace pads its rows to a multiple of four floats so that the GEMM that writes it and
// the two that read its gradient move 16-byte vectors although N = intent_num is odd); P is dense [R, N].
__global__ void __launch_bounds__(256) softmax_rows_kernel(int64_t R, int64_t N, const float* __restrict__ Z,
                                                           float* __restrict__ P, int64_t ldz) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nwarps) {
        const float* z = Z + r * ldz;
        float mx = -INFINITY;
        for (int64_t c = lane; c < N; c += 32) mx = fmaxf(mx, z[c]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int64_t c = lane; c < N; c += 32) sum += expf(z[c] - mx);
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int64_t c = lane; c < N; c += 32) P[r * N + c] = expf(z[c] - mx) * inv;
    }
}
int softmax_rows(int64_t R, int64_t N, const float* Z, float* P, cudaStream_t s, int64_t ldz) {
    if (R <= 0) return INTEL_OK;
    if (ldz <= 0) ldz = N;
    unsigned grid = stream_grid(ceil_div(R, 8), 8);
    LAUNCH(softmax_rows_kernel, dim3(grid), dim3(256), 0, s, R, N, Z, P, ldz);
    return check_launch("softmax_rows", 8.0 * R * N, 4.0 * R * N);
}

__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(int64_t R, int64_t N, const float* __restrict__ P,
                                                               const float* __restrict__ dP,
                                                               const float* __restrict__ dP2, float* __restrict__ dZ, int64_t ldz) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nwarps) {
        float dot = 0.f;
        for (int64_t c = lane; c < N; c += 32) {
            const float g = dP[r * N + c] + (dP2 ? dP2[r * N + c] : 0.f);
            dot = fmaf(P[r * N + c], g, dot);
        }
        dot = warp_sum(dot);
        for (int64_t c = lane; c < N; c += 32) {
            const float g = dP[r * N + c] + (dP2 ? dP2[r * N + c] : 0.f);
            dZ[r * ldz + c] = P[r * N + c] * (g - dot);
        }
    }
}
int softmax_rows_bwd(int64_t R, int64_t N, const float* P, const float* dP, const float* dP2, float* dZ,
                     cudaStream_t s, int64_t ldz) {
    if (R <= 0) return INTEL_OK;
    if (ldz <= 0) ldz = N;
    unsigned grid = stream_grid(ceil_div(R, 8), 8);
    LAUNCH(softmax_rows_bwd_kernel, dim3(grid), dim3(256), 0, s, R, N, P, dP, dP2, dZ, ldz);
    return check_launch("softmax_rows_bwd", 12.0 * R * N, 4.0 * R * N);
}

}  // namespace intel
