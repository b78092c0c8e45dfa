// Per-session attention kernels: LayerNorm, multi-head self-attention (fwd + bwd, recompute-based),
// pooled cross-attention, row softmax.  A whole session (L <= a few hundred tokens of 32 floats)
// fits one SM's shared memory, so every (session, head) pair is one CTA and nothing O(L^2) ever
// reaches HBM (the reference materialises [B,h,L,L] and [B,L,L] tensors, layers.py:54, attention.py:57).
#include "kernels.h"

namespace intel {

static const int LN_MAX_PER_LANE = 8;   // d <= 256

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(int64_t R, int d, const float* __restrict__ Z,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ Y,
                                                            float* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nwarps) {
        float v[LN_MAX_PER_LANE];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
            int c = lane + 32 * i;
            v[i] = (c < d) ? Z[r * d + c] : 0.f;
            sum += v[i];
        }
        const float mean = warp_sum(sum) / (float)d;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
            int c = lane + 32 * i;
            float t = (c < d) ? v[i] - mean : 0.f;
            sq += t * t;
        }
        const float rstd = rsqrtf(warp_sum(sq) / (float)d + 1e-5f);
#pragma unroll
        for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
            int c = lane + 32 * i;
            if (c < d) Y[r * d + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
        }
        if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
    }
}

int layernorm_fwd(int64_t R, int d, const float* Z, const float* gamma, const float* beta, float* Y, float* stats,
                  cudaStream_t s) {
    if (R <= 0) return INTEL_OK;
    INTEL_REQUIRE(d <= 32 * LN_MAX_PER_LANE, INTEL_ERR_UNSUPPORTED, "layernorm width %d > 256", d);
    unsigned grid = stream_grid(ceil_div(R, 8), 8);
    LAUNCH(layernorm_fwd_kernel, dim3(grid), dim3(256), 0, s, R, d, Z, gamma, beta, Y, stats);
    return check_launch("layernorm_fwd", 8.0 * R * d, 8.0 * R * d);
}

__global__ void __launch_bounds__(256) layernorm_bwd_kernel(int64_t R, int d, const float* __restrict__ dY,
                                                            const float* __restrict__ Z, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, float* __restrict__ dZ,
                                                            float* dgamma, float* dbeta) {
    __shared__ float red_g[8][32 * LN_MAX_PER_LANE];
    __shared__ float red_b[8][32 * LN_MAX_PER_LANE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float pg[LN_MAX_PER_LANE], pb[LN_MAX_PER_LANE];
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) { pg[i] = 0.f; pb[i] = 0.f; }
    for (int64_t r = warp; r < R; r += nwarps) {
        const float mean = stats[2 * r], rstd = stats[2 * r + 1];
        float xh[LN_MAX_PER_LANE], g[LN_MAX_PER_LANE];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
            int c = lane + 32 * i;
            if (c < d) {
                const float dy = dY[r * d + c];
                xh[i] = (Z[r * d + c] - mean) * rstd;
                g[i] = dy * gamma[c];
                pg[i] += dy * xh[i];
                pb[i] += dy;
            } else { xh[i] = 0.f; g[i] = 0.f; }
            s1 += g[i];
            s2 += g[i] * xh[i];
        }
        s1 = warp_sum(s1) / (float)d;
        s2 = warp_sum(s2) / (float)d;
#pragma unroll
        for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
            int c = lane + 32 * i;
            if (c < d) dZ[r * d + c] = rstd * (g[i] - s1 - xh[i] * s2);
        }
    }
#pragma unroll
    for (int i = 0; i < LN_MAX_PER_LANE; ++i) { red_g[wib][lane + 32 * i] = pg[i]; red_b[wib][lane + 32 * i] = pb[i]; }
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
    unsigned grid = stream_grid(ceil_div(R, 8), 4);
    LAUNCH(layernorm_bwd_kernel, dim3(grid), dim3(256), 0, s, R, d, dY, Z, stats, gamma, dZ, dgamma, dbeta);
    return check_launch("layernorm_bwd", 12.0 * R * d, 12.0 * R * d);
}

// ------------------------------------------------------------------------------------------------
// Multi-head attention, one CTA per (session, head).  K and V tiles of the head live in shared memory
// (row stride dk+1: conflict-free both for lane-per-key and lane-per-channel access); each warp owns a
// query at a time.  The reference's shift by the *global* max of the score tensor (layers.py:57) is a
// pure numerics choice; the row max is used instead (softmax is shift invariant).
static const int MHA_WARPS = 4;

__global__ void __launch_bounds__(MHA_WARPS * 32) mha_fwd_kernel(int64_t T, int d, int heads,
                                                                 const float* __restrict__ QKV,
                                                                 const int64_t* __restrict__ lens,
                                                                 float* __restrict__ O) {
    DYN_SMEM(float, sm);
    const int dk = d / heads, st = dk + 1;
    const int64_t b = blockIdx.x / heads;
    const int h = blockIdx.x % heads;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t nk = lens ? lens[b] : T;
    if (nk > T) nk = T;
    float* Ks = sm;
    float* Vs = Ks + T * st;
    float* qb = Vs + T * st + w * dk;              // [MHA_WARPS][dk]
    float* pb = Vs + T * st + MHA_WARPS * dk + w * T;   // [MHA_WARPS][T]
    const float* base = QKV + b * T * 3 * d + h * dk;
    for (int64_t e = threadIdx.x; e < nk * dk; e += blockDim.x) {
        const int64_t j = e / dk;
        const int c = (int)(e % dk);
        Ks[j * st + c] = base[j * 3 * d + d + c];
        Vs[j * st + c] = base[j * 3 * d + 2 * d + c];
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)dk);
    for (int64_t i = w; i < T; i += MHA_WARPS) {
        for (int c = lane; c < dk; c += 32) qb[c] = base[i * 3 * d + c];
        __syncwarp();
        float mx = -INFINITY;
        for (int64_t j = lane; j < nk; j += 32) {
            float sdot = 0.f;
            for (int c = 0; c < dk; ++c) sdot = fmaf(qb[c], Ks[j * st + c], sdot);
            sdot *= scale;
            pb[j] = sdot;
            mx = fmaxf(mx, sdot);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int64_t j = lane; j < nk; j += 32) {
            const float e = expf(pb[j] - mx);
            pb[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = (nk > 0) ? 1.0f / sum : 0.f;
        __syncwarp();
        for (int c = lane; c < dk; c += 32) {
            float acc = 0.f;
            for (int64_t j = 0; j < nk; ++j) acc = fmaf(pb[j], Vs[j * st + c], acc);
            O[(b * T + i) * d + h * dk + c] = acc * inv;
        }
        __syncwarp();
    }
}

static size_t mha_fwd_smem(int64_t T, int dk) { return (size_t)(2 * T * (dk + 1) + MHA_WARPS * dk + MHA_WARPS * T) * 4; }
static size_t mha_bwd_smem(int64_t T, int dk) { return (size_t)(4 * T * (dk + 1) + 3 * T + 2 * MHA_WARPS * T) * 4; }
static const size_t kMaxSmem = 200 * 1024;

int mha_fwd(int64_t B, int64_t T, int d, int heads, const float* QKV, const int64_t* lens, float* O, cudaStream_t s) {
    if (B <= 0 || T <= 0) return INTEL_OK;
    INTEL_REQUIRE(heads > 0 && d % heads == 0, INTEL_ERR_ARG, "mha: d=%d not divisible by heads=%d", d, heads);
    const size_t smem = mha_fwd_smem(T, d / heads);
    INTEL_REQUIRE(smem <= kMaxSmem, INTEL_ERR_UNSUPPORTED, "mha_fwd: list length %lld too long for one SM", (long long)T);
    if (smem > 48 * 1024) cudaFuncSetAttribute(mha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LAUNCH(mha_fwd_kernel, dim3((unsigned)(B * heads)), dim3(MHA_WARPS * 32), smem, s, T, d, heads, QKV, lens, O);
    return check_launch("mha_fwd", 16.0 * B * T * d, 4.0 * B * T * T * d);
}

// Backward by recomputation.  Phase A (warp per query): softmax statistics (m, l), delta = sum_j p dP,
// and dQ.  Phase B (warp per key): p and dS are recomputed from the statistics -> dK, dV.  No atomics.
__global__ void __launch_bounds__(MHA_WARPS * 32) mha_bwd_kernel(int64_t T, int d, int heads,
                                                                 const float* __restrict__ QKV,
                                                                 const int64_t* __restrict__ lens,
                                                                 const float* __restrict__ dO,
                                                                 float* __restrict__ dQKV) {
    DYN_SMEM(float, sm);
    const int dk = d / heads, st = dk + 1;
    const int64_t b = blockIdx.x / heads;
    const int h = blockIdx.x % heads;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t nk = lens ? lens[b] : T;
    if (nk > T) nk = T;
    float* Qs = sm;
    float* Ks = Qs + T * st;
    float* Vs = Ks + T * st;
    float* Gs = Vs + T * st;               // dO tile
    float* sm_m = Gs + T * st;             // [T] row max
    float* sm_l = sm_m + T;                // [T] row sum
    float* sm_d = sm_l + T;                // [T] delta
    float* pb = sm_d + T + w * T;          // [MHA_WARPS][T]
    float* db = sm_d + T + MHA_WARPS * T + w * T;   // [MHA_WARPS][T]
    const float* base = QKV + b * T * 3 * d + h * dk;
    float* dbase = dQKV + b * T * 3 * d + h * dk;
    for (int64_t e = threadIdx.x; e < T * dk; e += blockDim.x) {
        const int64_t j = e / dk;
        const int c = (int)(e % dk);
        Qs[j * st + c] = base[j * 3 * d + c];
        Ks[j * st + c] = base[j * 3 * d + d + c];
        Vs[j * st + c] = base[j * 3 * d + 2 * d + c];
        Gs[j * st + c] = dO[(b * T + j) * d + h * dk + c];
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)dk);
    // ---- phase A: per query ----
    for (int64_t i = w; i < T; i += MHA_WARPS) {
        float mx = -INFINITY;
        for (int64_t j = lane; j < nk; j += 32) {
            float sdot = 0.f, gdot = 0.f;
            for (int c = 0; c < dk; ++c) {
                sdot = fmaf(Qs[i * st + c], Ks[j * st + c], sdot);
                gdot = fmaf(Gs[i * st + c], Vs[j * st + c], gdot);
            }
            sdot *= scale;
            pb[j] = sdot;
            db[j] = gdot;
            mx = fmaxf(mx, sdot);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int64_t j = lane; j < nk; j += 32) {
            const float e = expf(pb[j] - mx);
            pb[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = (nk > 0) ? 1.0f / sum : 0.f;
        float delta = 0.f;
        for (int64_t j = lane; j < nk; j += 32) delta = fmaf(pb[j] * inv, db[j], delta);
        delta = warp_sum(delta);
        for (int64_t j = lane; j < nk; j += 32) db[j] = pb[j] * inv * (db[j] - delta) * scale;   // dS
        if (lane == 0) { sm_m[i] = mx; sm_l[i] = inv; sm_d[i] = delta; }
        __syncwarp();
        for (int c = lane; c < dk; c += 32) {
            float acc = 0.f;
            for (int64_t j = 0; j < nk; ++j) acc = fmaf(db[j], Ks[j * st + c], acc);
            dbase[i * 3 * d + c] = acc;
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- phase B: per key ----
    for (int64_t j = w; j < T; j += MHA_WARPS) {
        if (j >= nk) {   // masked key: no gradient
            for (int c = lane; c < dk; c += 32) { dbase[j * 3 * d + d + c] = 0.f; dbase[j * 3 * d + 2 * d + c] = 0.f; }
            continue;
        }
        for (int64_t i = lane; i < T; i += 32) {
            float sdot = 0.f, gdot = 0.f;
            for (int c = 0; c < dk; ++c) {
                sdot = fmaf(Qs[i * st + c], Ks[j * st + c], sdot);
                gdot = fmaf(Gs[i * st + c], Vs[j * st + c], gdot);
            }
            const float p = expf(sdot * scale - sm_m[i]) * sm_l[i];
            pb[i] = p;
            db[i] = p * (gdot - sm_d[i]) * scale;
        }
        __syncwarp();
        for (int c = lane; c < dk; c += 32) {
            float ak = 0.f, av = 0.f;
            for (int64_t i = 0; i < T; ++i) {
                ak = fmaf(db[i], Qs[i * st + c], ak);
                av = fmaf(pb[i], Gs[i * st + c], av);
            }
            dbase[j * 3 * d + d + c] = ak;
            dbase[j * 3 * d + 2 * d + c] = av;
        }
        __syncwarp();
    }
}

int mha_bwd(int64_t B, int64_t T, int d, int heads, const float* QKV, const int64_t* lens, const float* dO, float* dQKV,
            cudaStream_t s) {
    if (B <= 0 || T <= 0) return INTEL_OK;
    INTEL_REQUIRE(heads > 0 && d % heads == 0, INTEL_ERR_ARG, "mha: d=%d not divisible by heads=%d", d, heads);
    const size_t smem = mha_bwd_smem(T, d / heads);
    INTEL_REQUIRE(smem <= kMaxSmem, INTEL_ERR_UNSUPPORTED, "mha_bwd: list length %lld too long for one SM", (long long)T);
    if (smem > 48 * 1024) cudaFuncSetAttribute(mha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LAUNCH(mha_bwd_kernel, dim3((unsigned)(B * heads)), dim3(MHA_WARPS * 32), smem, s, T, d, heads, QKV, lens, dO, dQKV);
    return check_launch("mha_bwd", 28.0 * B * T * d, 16.0 * B * T * T * d);
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
    for (int64_t j = lane; j < L; j += 32) {
        float a = 0.f;
        for (int c = 0; c < d; ++c) a = fmaf(x[j * d + c], q[c], a);
        a *= scale;
        pb[j] = a;
        mx = fmaxf(mx, a);             // row max over ALL slots, pads included (attention.py:57)
    }
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
    for (int64_t j = lane; j < n; j += 32) {
        float a = 0.f;
        for (int c = 0; c < d; ++c) a = fmaf(x[j * d + c], g[c], a);
        ab[j] = a;                                // dp_j
        delta = fmaf(p[b * L + j], a, delta);
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
// Row softmax (pred_layer(...).softmax, IntEL.py:153) and its backward; one warp per row.
__global__ void __launch_bounds__(256) softmax_rows_kernel(int64_t R, int64_t N, const float* __restrict__ Z,
                                                           float* __restrict__ P) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < R; r += nwarps) {
        const float* z = Z + r * N;
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
int softmax_rows(int64_t R, int64_t N, const float* Z, float* P, cudaStream_t s) {
    if (R <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(R, 8), 8);
    LAUNCH(softmax_rows_kernel, dim3(grid), dim3(256), 0, s, R, N, Z, P);
    return check_launch("softmax_rows", 8.0 * R * N, 4.0 * R * N);
}

__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(int64_t R, int64_t N, const float* __restrict__ P,
                                                               const float* __restrict__ dP,
                                                               const float* __restrict__ dP2, float* __restrict__ dZ) {
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
            dZ[r * N + c] = P[r * N + c] * (g - dot);
        }
    }
}
int softmax_rows_bwd(int64_t R, int64_t N, const float* P, const float* dP, const float* dP2, float* dZ,
                     cudaStream_t s) {
    if (R <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(R, 8), 8);
    LAUNCH(softmax_rows_bwd_kernel, dim3(grid), dim3(256), 0, s, R, N, P, dP, dP2, dZ);
    return check_launch("softmax_rows_bwd", 12.0 * R * N, 4.0 * R * N);
}

}  // namespace intel
