// Listwise / pairwise / pointwise ensemble losses with fused gradients, the "diversity" regulariser
// and the intent CE+KL loss.  One warp per session; the reference's [B,L,L] and [B,L,L,K] float64
// temporaries (Listloss.py:33-40, BPRloss.py:45-53) never exist: pair terms are formed in registers,
// only for the positive rows (r_i > 0), from s / r staged in shared memory.
#include "kernels.h"
#include "../../include/intel_b200.h"

namespace intel {

static const int LOSS_WARPS = 4;
static const int LOSS_MAX_K = 16;

struct LossArgs {
    int64_t B, L;
    int K;
    const float* ens;
    const float* weights;
    const double* scores;
    const int64_t* ranking;
    const int64_t* lens;
    int cal_div;
    float alpha;
    double* out;          // out[0] += loss terms
    float* d_ens;
    float* d_weights;
    const float* noise;   // BPR only
    uint64_t seed;
};

__device__ __forceinline__ void stage_session(const LossArgs& a, int64_t b, int lane, float* s, int* r, float* ds,
                                              int64_t& n, int& npos) {
    n = a.lens[b];
    if (n > a.L) n = a.L;
    int cnt = 0;
    for (int64_t j = lane; j < a.L; j += 32) {
        s[j] = a.ens[b * a.L + j];
        int64_t rv = a.ranking[b * a.L + j];
        r[j] = rv > 0 ? (int)rv : 0;            // torch.clamp(ranking, 0, max)
        ds[j] = 0.f;
        cnt += (rv > 0);
    }
    npos = warp_sum_i(cnt);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Plackett-Luce style list loss (Listloss.py:12-43), O(L K) per session instead of the reference's [L, L, K] pair tensors:
//   E_i = sum_{j valid, r_j < r_i} exp(s_j - s_i),  l_i = log(1 + E_i) for r_i > 0
//   F_ik = sum_j exp(s_j - s_i) (u_ik - u_jk),  u = x - s,   div_i = sum_k w_ik F_ik^2 / (2 (1+E_i)^2)
//   loss = mean_b( sum_i l_i / #pos ) - alpha * mean_b( sum_i div_i / #pos )
// The clamped rankings take four values, so every "sum over the items ranked below i" is a prefix over three bucket sums
// (SURVEY.md 8a-7):  A_rho = sum_{r_j = rho} e^{s_j - m},  C_rho,k = sum_{r_j = rho} e^{s_j - m} u_jk  (m = max s, the
// shift the reference does not apply: same value, no overflow until s differs by ~88 inside one session)
//   E_i = e^{m - s_i} A_{<r_i},   F_ik = u_ik E_i - e^{m - s_i} C_{<r_i,k}
// and the gradient that item j receives from all positives ranked above it is a suffix over the buckets of
//   a_i = e^{m - s_i} (c1_i + sum_k q_ik (u_ik + 1)),  b_ik = e^{m - s_i} q_ik,   c1_i = cl/(1+E_i) - cd G_i/(1+E_i)^3,
//   q_ik = cd w_ik F_ik / (1+E_i)^2:   d s_j += e^{s_j - m} (a_{>r_j} - sum_k u_jk b_{>r_j,k}).
// pair form (one pass over the lower-ranked items per positive): exact for any number of rank levels; the bucket kernel
// below falls back to it for a session whose clamped rankings exceed the three levels the reference's datasets emit
__device__ __noinline__ void pl_pairs_session(const LossArgs& a, int64_t b, int lane, const float* s, const int* r, float* ds, int64_t n,
                                             int npos) {
    const int64_t L = a.L;
    const int K = a.K;
    const float inv_pos = 1.0f / (float)npos;      // 0 positives -> inf -> NaN loss, as the reference
    const float cl = inv_pos / (float)a.B;
    const float cd = -a.alpha * cl;
    double loss_acc = 0.0;
    const double* xb = a.scores + b * L * K;
    const float* wb = a.weights ? a.weights + b * L * K : nullptr;
    float* dwb = a.d_weights ? a.d_weights + b * L * K : nullptr;
    if (dwb) for (int64_t e = lane; e < L * K; e += 32) dwb[e] = 0.f;
    __syncwarp();
    for (int64_t i = 0; i < L; ++i) {
        const int ri = r[i];
        if (ri <= 0 || i >= n) continue;             // invalid positive rows: log(1+0) = 0
        const float si = s[i];
        float xi[LOSS_MAX_K], wi[LOSS_MAX_K], F[LOSS_MAX_K];
#pragma unroll
        for (int k = 0; k < LOSS_MAX_K; ++k) {
            xi[k] = (a.cal_div && k < K) ? (float)xb[i * K + k] : 0.f;
            wi[k] = (a.cal_div && k < K) ? wb[i * K + k] : 0.f;
            F[k] = 0.f;
        }
        float E = 0.f;
        for (int64_t j = lane; j < n; j += 32) {
            if (r[j] < ri) {
                const float D = si - s[j];
                const float e = expf(-D);
                E += e;
                if (a.cal_div) {
#pragma unroll
                    for (int k = 0; k < LOSS_MAX_K; ++k)
                        if (k < K) F[k] = fmaf(e, (xi[k] - (float)xb[j * K + k]) - D, F[k]);
                }
            }
        }
        E = warp_sum(E);
        const float opE = 1.f + E;
        float G = 0.f, H = 0.f;      // G = sum_k w F^2, H = sum_k w F (-F - E)
        if (a.cal_div) {
#pragma unroll
            for (int k = 0; k < LOSS_MAX_K; ++k) {
                if (k < K) {
                    F[k] = warp_sum(F[k]);
                    G = fmaf(wi[k] * F[k], F[k], G);
                    H = fmaf(wi[k] * F[k], -F[k] - E, H);
                }
            }
        }
        const float inv1 = 1.f / opE, inv2 = inv1 * inv1, inv3 = inv2 * inv1;
        loss_acc += (double)(logf(opE) * inv_pos);
        float dsi = cl * (-E * inv1);
        if (a.cal_div) {
            loss_acc += (double)(-a.alpha * inv_pos * G * 0.5f * inv2);
            dsi += cd * (H * inv2 + G * E * inv3);
            if (dwb) {
                float fl = 0.f;        // F[lane] without dynamic register indexing
#pragma unroll
                for (int k = 0; k < LOSS_MAX_K; ++k) fl = (k == lane) ? F[k] : fl;
                if (lane < K) dwb[i * K + lane] = cd * fl * fl * 0.5f * inv2;
            }
        }
        for (int64_t j = lane; j < n; j += 32) {
            if (r[j] < ri) {
                const float D = si - s[j];
                const float e = expf(-D);
                float coef = cl * inv1;
                if (a.cal_div) {
                    float t = 0.f;
#pragma unroll
                    for (int k = 0; k < LOSS_MAX_K; ++k)
                        if (k < K) t = fmaf(wi[k] * F[k], (xi[k] - (float)xb[j * K + k]) - D + 1.f, t);
                    coef += cd * (t * inv2 - G * inv3);
                }
                ds[j] = fmaf(e, coef, ds[j]);
            }
        }
        if (lane == (int)(i & 31)) ds[i] += dsi;
        __syncwarp();
    }
    __syncwarp();
    for (int64_t j = lane; j < L; j += 32) a.d_ens[b * L + j] = ds[j];
    if (npos == 0) loss_acc = (double)NAN;      // 0 / 0 positives, as the reference
    if (lane == 0) atomicAdd(a.out, loss_acc / (double)a.B);
}

static const int PL_LEVELS = 3;         // rank levels that can sit below a positive: 0, 1, 2

// KM = compile-time bound on model_num: the bucket sums are KM-wide register arrays (K <= 4 covers the reference's datasets
// at a third of the registers of the general instance, i.e. twice the sessions in flight per SM)
template <int KM>
__global__ void __launch_bounds__(LOSS_WARPS * 32) loss_pl_kernel(LossArgs a) {
    DYN_SMEM(float, sm);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * LOSS_WARPS + w;
    if (b >= a.B) return;
    const int64_t L = a.L;
    const int K = a.K;
    float* s = sm + (size_t)w * 3 * L;
    int* r = reinterpret_cast<int*>(s + L);
    float* ds = s + 2 * L;
    int64_t n;
    int npos;
    stage_session(a, b, lane, s, r, ds, n, npos);
    {
        int mr = 0;
        for (int64_t j = lane; j < n; j += 32) mr = r[j] > mr ? r[j] : mr;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, mr, o); mr = t > mr ? t : mr; }
        if (mr > PL_LEVELS) {                      // more rank levels than pay / fav / click: pair form
            pl_pairs_session(a, b, lane, s, r, ds, n, npos);
            return;
        }
    }
    const float inv_pos = 1.0f / (float)npos;      // 0 positives -> inf -> NaN loss, as the reference
    const float cl = inv_pos / (float)a.B;
    const float cd = -a.alpha * cl;
    const double* xb = a.scores + b * L * K;
    const float* wb = a.weights ? a.weights + b * L * K : nullptr;
    float* dwb = a.d_weights ? a.d_weights + b * L * K : nullptr;
    const bool div = a.cal_div != 0;
    // ---- pass A: bucket sums over the valid items ----
    float m = -INFINITY;
    for (int64_t j = lane; j < n; j += 32) m = fmaxf(m, s[j]);
    m = warp_max(m);
    float A[PL_LEVELS], C[PL_LEVELS][KM];
#pragma unroll
    for (int q = 0; q < PL_LEVELS; ++q) {
        A[q] = 0.f;
#pragma unroll
        for (int k = 0; k < KM; ++k) C[q][k] = 0.f;
    }
    for (int64_t j = lane; j < n; j += 32) {
        const int rj = r[j];
        if (rj >= PL_LEVELS) continue;
        const float e = expf(s[j] - m);
#pragma unroll
        for (int q = 0; q < PL_LEVELS; ++q) A[q] += (rj == q) ? e : 0.f;
        if (div) {
#pragma unroll
            for (int k = 0; k < KM; ++k) {
                if (k < K) {
                    const float eu = e * ((float)xb[j * K + k] - s[j]);
#pragma unroll
                    for (int q = 0; q < PL_LEVELS; ++q) C[q][k] += (rj == q) ? eu : 0.f;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < PL_LEVELS; ++q) {
        A[q] = warp_sum(A[q]);
        if (div) {
#pragma unroll
            for (int k = 0; k < KM; ++k)
                if (k < K) C[q][k] = warp_sum(C[q][k]);
        }
    }
    // prefixes: slot q holds the sum over the levels < q + 1, i.e. what a positive of rank q + 1 sees below it
#pragma unroll
    for (int q = 1; q < PL_LEVELS; ++q) {
        A[q] += A[q - 1];
#pragma unroll
        for (int k = 0; k < KM; ++k) C[q][k] += C[q - 1][k];
    }
    // ---- pass B: the positives (lanes over items), bucket sums of their gradient coefficients by rank ----
    float Ga[PL_LEVELS], Gb[PL_LEVELS][KM];      // level q: positives of rank q + 1
#pragma unroll
    for (int q = 0; q < PL_LEVELS; ++q) {
        Ga[q] = 0.f;
#pragma unroll
        for (int k = 0; k < KM; ++k) Gb[q][k] = 0.f;
    }
    double loss_acc = 0.0;
    for (int64_t i = lane; i < L; i += 32) {
        const int ri = r[i];
        const bool pos = ri > 0 && i < n;
        if (dwb && !pos)
            for (int k = 0; k < K; ++k) dwb[i * K + k] = 0.f;
        if (!pos) continue;
        const int lv = ri - 1;
        float Alo = 0.f;
#pragma unroll
        for (int q = 0; q < PL_LEVELS; ++q) Alo = (lv == q) ? A[q] : Alo;
        const float si = s[i], ei = expf(m - si);
        const float E = ei * Alo, opE = 1.f + E;
        const float inv1 = 1.f / opE, inv2 = inv1 * inv1, inv3 = inv2 * inv1;
        loss_acc += (double)(logf(opE) * inv_pos);
        float dsi = cl * (-E * inv1);
        float alpha_i = cl * inv1;
        if (div) {
            float G = 0.f, H = 0.f, t1 = 0.f;
            float F[KM], wi[KM], ui[KM];
#pragma unroll
            for (int k = 0; k < KM; ++k) {
                F[k] = 0.f; wi[k] = 0.f; ui[k] = 0.f;
                if (k < K) {
                    float Clo = 0.f;
#pragma unroll
                    for (int q = 0; q < PL_LEVELS; ++q) Clo = (lv == q) ? C[q][k] : Clo;
                    ui[k] = (float)xb[i * K + k] - si;
                    wi[k] = wb[i * K + k];
                    F[k] = ui[k] * E - ei * Clo;
                    G = fmaf(wi[k] * F[k], F[k], G);
                    H = fmaf(wi[k] * F[k], -F[k] - E, H);
                }
            }
            loss_acc += (double)(-a.alpha * inv_pos * G * 0.5f * inv2);
            dsi += cd * (H * inv2 + G * E * inv3);
            alpha_i -= cd * G * inv3;
#pragma unroll
            for (int k = 0; k < KM; ++k) {
                if (k < K) {
                    const float q_ik = cd * inv2 * wi[k] * F[k];
                    t1 = fmaf(q_ik, ui[k] + 1.f, t1);
                    const float bik = ei * q_ik;
#pragma unroll
                    for (int q = 0; q < PL_LEVELS; ++q) Gb[q][k] += (lv == q) ? bik : 0.f;
                    if (dwb) dwb[i * K + k] = cd * F[k] * F[k] * 0.5f * inv2;
                }
            }
            alpha_i += t1;
        }
        alpha_i *= ei;
#pragma unroll
        for (int q = 0; q < PL_LEVELS; ++q) Ga[q] += (lv == q) ? alpha_i : 0.f;
        ds[i] = dsi;                                    // each slot is owned by one lane
    }
#pragma unroll
    for (int q = 0; q < PL_LEVELS; ++q) {
        Ga[q] = warp_sum(Ga[q]);
        if (div) {
#pragma unroll
            for (int k = 0; k < KM; ++k)
                if (k < K) Gb[q][k] = warp_sum(Gb[q][k]);
        }
    }
    // suffixes: slot q = sum over the positives of rank > q, i.e. what an item of level q receives
#pragma unroll
    for (int q = PL_LEVELS - 2; q >= 0; --q) {
        Ga[q] += Ga[q + 1];
#pragma unroll
        for (int k = 0; k < KM; ++k) Gb[q][k] += Gb[q + 1][k];
    }
    __syncwarp();
    // ---- pass C: what every valid item receives from the positives ranked above it ----
    for (int64_t j = lane; j < L; j += 32) {
        float g = ds[j];
        const int rj = r[j];
        if (j < n && rj < PL_LEVELS) {
            float ga = 0.f;
#pragma unroll
            for (int q = 0; q < PL_LEVELS; ++q) ga = (rj == q) ? Ga[q] : ga;
            float acc = ga;
            if (div) {
#pragma unroll
                for (int k = 0; k < KM; ++k) {
                    if (k < K) {
                        float gb = 0.f;
#pragma unroll
                        for (int q = 0; q < PL_LEVELS; ++q) gb = (rj == q) ? Gb[q][k] : gb;
                        acc = fmaf(-((float)xb[j * K + k] - s[j]), gb, acc);
                    }
                }
            }
            g = fmaf(expf(s[j] - m), acc, g);
        }
        a.d_ens[b * L + j] = g;
    }
    loss_acc = warp_sum_d(loss_acc);
    if (npos == 0) loss_acc = (double)NAN;      // 0 / 0 positives, as the reference
    if (lane == 0) atomicAdd(a.out, loss_acc / (double)a.B);
}

// ------------------------------------------------------------------------------------------------
// BPR (BPRloss.py:12-56): for each positive i pick one j among the valid items of the closest lower
// rank (argmax of the injected noise, BPRloss.py:26-28); with no candidate the argmax falls on pure
// noise over all L slots.  l_i = -log sigmoid(s_i - s_j);  div_i = sum_k w_ik sig'(D) ((x_ik-x_jk) - D)^2
__device__ __forceinline__ uint32_t bpr_hash(uint64_t seed, uint64_t idx) { return hash_u32(seed, idx); }

__global__ void __launch_bounds__(LOSS_WARPS * 32) loss_bpr_kernel(LossArgs a) {
    DYN_SMEM(float, sm);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * LOSS_WARPS + w;
    if (b >= a.B) return;
    const int64_t L = a.L;
    const int K = a.K;
    float* s = sm + (size_t)w * 3 * L;
    int* r = reinterpret_cast<int*>(s + L);
    float* ds = s + 2 * L;
    int64_t n;
    int npos;
    stage_session(a, b, lane, s, r, ds, n, npos);
    const float inv_pos = 1.0f / (float)npos;
    const float cl = inv_pos / (float)a.B;
    const float cd = -a.alpha * cl;
    double loss_acc = 0.0;
    const double* xb = a.scores + b * L * K;
    const float* wb = a.weights ? a.weights + b * L * K : nullptr;
    float* dwb = a.d_weights ? a.d_weights + b * L * K : nullptr;
    if (dwb) for (int64_t e = lane; e < L * K; e += 32) dwb[e] = 0.f;
    __syncwarp();
    for (int64_t i = 0; i < L; ++i) {
        const int ri = r[i];
        if (ri <= 0) continue;
        const bool vi = i < n;
        // closest lower rank among valid items
        int lower = -1;
        if (vi)
            for (int64_t j = lane; j < n; j += 32)
                if (r[j] < ri && r[j] > lower) lower = r[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, lower, o); lower = t > lower ? t : lower; }
        // argmax of the noise over the candidates (or over every slot when there is none)
        const int64_t jend = (lower >= 0) ? n : L;
        float best = -1.f;
        int64_t bj = L;
        for (int64_t j = lane; j < jend; j += 32) {
            if (lower >= 0 && r[j] != lower) continue;
            float u;
            if (a.noise) u = a.noise[(b * L + i) * L + j];
            else u = (float)bpr_hash(a.seed, (uint64_t)((b * L + i) * L + j)) * (1.0f / 4294967296.0f);
            if (u > best) { best = u; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int64_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        const int64_t sel = bj < L ? bj : 0;
        const float D = s[i] - s[sel];
        const float sg = sigmoidf_(D);
        // -log(sigmoid(D)) evaluated as softplus(-D) (no underflow for very negative D)
        const float li = (D > 0.f) ? log1pf(expf(-D)) : (-D + log1pf(expf(D)));
        float dD = cl * (sg - 1.f);
        if (lane == 0) loss_acc += (double)(li * inv_pos);
        if (a.cal_div) {
            const float dsg = sg * (1.f - sg), ddsg = dsg * (1.f - 2.f * sg);
            float div = 0.f, dd = 0.f;
            for (int k = 0; k < K; ++k) {
                const float u = ((float)xb[i * K + k] - (float)xb[sel * K + k]) - D;
                const float wk = wb[i * K + k];
                div = fmaf(wk * dsg, u * u, div);
                dd += wk * (ddsg * u * u - 2.f * dsg * u);
                if (dwb && lane == 0) dwb[i * K + k] = cd * dsg * u * u;
            }
            if (lane == 0) loss_acc += (double)(-a.alpha * inv_pos * div);
            dD += cd * dd;
        }
        if (lane == 0) { ds[i] += dD; ds[sel] -= dD; }
        __syncwarp();
    }
    __syncwarp();
    for (int64_t j = lane; j < L; j += 32) a.d_ens[b * L + j] = ds[j];
    if (npos == 0) loss_acc = (double)NAN;      // 0 / 0 positives, as the reference
    if (lane == 0) atomicAdd(a.out, loss_acc / (double)a.B);
}

// ------------------------------------------------------------------------------------------------
// MSE (MSEloss.py:12-30): sum_valid (s - r)^2 / n ; diversity -alpha * sum_valid sum_k w (x - s)^2 / n
__global__ void __launch_bounds__(LOSS_WARPS * 32) loss_mse_kernel(LossArgs a) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * LOSS_WARPS + w;
    if (b >= a.B) return;
    const int64_t L = a.L;
    const int K = a.K;
    int64_t n = a.lens[b];
    if (n > L) n = L;
    const float inv_n = 1.0f / (float)n;
    const float cl = inv_n / (float)a.B;
    float acc = 0.f;
    for (int64_t j = lane; j < L; j += 32) {
        float g = 0.f;
        const bool valid = j < n;
        if (valid) {
            const float sj = a.ens[b * L + j];
            const int64_t rv = a.ranking[b * L + j];
            const float diff = sj - (float)(rv > 0 ? rv : 0);
            acc = fmaf(diff, diff, acc);
            g = 2.f * diff * cl;
            if (a.cal_div) {
                for (int k = 0; k < K; ++k) {
                    const float u = (float)a.scores[(b * L + j) * K + k] - sj;
                    const float wk = a.weights[(b * L + j) * K + k];
                    acc -= a.alpha * wk * u * u;
                    g += a.alpha * cl * 2.f * wk * u;
                    if (a.d_weights) a.d_weights[(b * L + j) * K + k] = -a.alpha * cl * u * u;
                }
            }
        } else if (a.cal_div && a.d_weights) {
            for (int k = 0; k < K; ++k) a.d_weights[(b * L + j) * K + k] = 0.f;
        }
        a.d_ens[b * L + j] = g;
    }
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(a.out, (double)(acc * inv_n) / (double)a.B);
}

static int launch_loss(int which, LossArgs a, cudaStream_t s) {
    if (a.B <= 0) return INTEL_OK;
    INTEL_REQUIRE(a.K <= LOSS_MAX_K, INTEL_ERR_UNSUPPORTED, "model_num %d > %d", a.K, LOSS_MAX_K);
    INTEL_REQUIRE(a.ens && a.scores && a.ranking && a.lens && a.out && a.d_ens, INTEL_ERR_ARG, "loss: null pointer");
    INTEL_REQUIRE(!a.cal_div || a.weights, INTEL_ERR_ARG, "loss: cal_diversity needs the weights tensor");
    INTEL_TRY(fill_zero(a.out, sizeof(double), s));
    const size_t smem = (size_t)LOSS_WARPS * 3 * a.L * 4;
    INTEL_REQUIRE(smem <= 200 * 1024, INTEL_ERR_UNSUPPORTED, "loss: list length %lld too long", (long long)a.L);
    dim3 grid((unsigned)ceil_div(a.B, LOSS_WARPS)), block(LOSS_WARPS * 32);
    if (which == 0) {
        if (a.K <= 4) {
            ensure_smem(loss_pl_kernel<4>, smem);
            LAUNCH(loss_pl_kernel<4>, grid, block, smem, s, a);
        } else {
            ensure_smem(loss_pl_kernel<LOSS_MAX_K>, smem);
            LAUNCH(loss_pl_kernel<LOSS_MAX_K>, grid, block, smem, s, a);
        }
    } else if (which == 1) {
        ensure_smem(loss_bpr_kernel, smem);
        LAUNCH(loss_bpr_kernel, grid, block, smem, s, a);
    } else {
        LAUNCH(loss_mse_kernel, grid, block, 0, s, a);
    }
    return check_launch(which == 0 ? "loss_pl" : (which == 1 ? "loss_bpr" : "loss_mse"),
                        (double)a.B * a.L * (4.0 + 8.0 + 4.0 + (a.cal_div ? 16.0 * a.K : 0.0)), 0.0);
}

// ------------------------------------------------------------------------------------------------
// Intent loss (BaseIntloss.py:30-67).  pass 1: global min of the predictions decides the
// "make soft" branch (BaseIntloss.py:32,47) on the device; pass 2: one warp per session.
__device__ __forceinline__ int float_order_key(float f) {
    int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}

__global__ void __launch_bounds__(256) intent_min_kernel(int64_t n, const float* __restrict__ p, int* key_out) {
    int best = 0x7fffffff;
#pragma unroll 8
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = float_order_key(p[e]);
        best = k < best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, best, o); best = t < best ? t : best; }
    if ((threadIdx.x & 31) == 0) atomicMin(key_out, best);
}

__global__ void __launch_bounds__(256) intent_loss_kernel(int64_t B, int64_t I, const float* __restrict__ pred,
                                                          const double* __restrict__ truth, float kw, float T2,
                                                          const int* __restrict__ min_key, double* out,
                                                          float* __restrict__ d_pred) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool soften = (*min_key == float_order_key(0.0f)) || (*min_key == float_order_key(-0.0f));
    const float invB = 1.0f / (float)B;
    for (int64_t b = warp; b < B; b += nwarps) {
        const float* p = pred + b * I;
        const double* t = truth + b * I;
        float S = 1.f;
        if (soften) {
            float acc = 0.f;
            for (int64_t c = lane; c < I; c += 32) acc += p[c] + 1e-6f;
            S = warp_sum(acc);
        }
        double ce = 0.0, kl = 0.0;
        float gdot = 0.f;
        // four elements per lane in flight (the loop is otherwise one 4 + 8 byte load deep and latency bound); each lane
        // still accumulates its elements in ascending order
        for (int64_t c0 = lane; c0 < I; c0 += 128) {
            float pv[4];
            double tv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t c = c0 + 32 * u;
                pv[u] = c < I ? p[c] : 1.f;
                tv[u] = c < I ? t[c] : -1.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t c = c0 + 32 * u;
                if (c >= I) break;
                const float q = soften ? (pv[u] + 1e-6f) / S : pv[u];
                const double tc = tv[u];
                const float lq = logf(q);
                float g = 0.f;
                if (tc > 0.0) {
                    ce -= tc * (double)lq;
                    const float t32 = (float)tc;
                    kl += (double)(t32 * logf(t32) - t32 * lq);
                    g = -((1.f - kw) * (float)tc + kw * T2 * t32) / q;
                } else if (tc == 0.0) {
                    ce -= (double)logf(1.f - q);
                    g = (1.f - kw) / (1.f - q);
                }
                g *= invB;
                gdot = fmaf(g, q, gdot);
                d_pred[b * I + c] = g;
            }
        }
        ce = warp_sum_d(ce);
        kl = warp_sum_d(kl) * (double)T2;
        if (soften) {
            gdot = warp_sum(gdot);
            for (int64_t c = lane; c < I; c += 32) d_pred[b * I + c] = (d_pred[b * I + c] - gdot) / S;
        }
        if (lane == 0) {
            atomicAdd(out + 0, (ce * (1.0 - (double)kw) + kl * (double)kw) / (double)B);
            atomicAdd(out + 1, ce / (double)B);
            atomicAdd(out + 2, kl / (double)B);
        }
    }
}

__global__ void __launch_bounds__(256) scale_kernel(int64_t n, const float* __restrict__ x, const double* a, double ca,
                                                    const double* b, double cb, float* __restrict__ y) {
    const float f = (float)((a ? ca * *a : 0.0) + (b ? cb * *b : 0.0));
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        y[e] = x[e] * f;
}

// ------------------------------------------------------------------------------------------------
// LambdaRank lambdas (helpers/LambdaRankRunner.py:315-344, compute_lambda_new): for t = clamp(ranking, 0), list slot
// discounts d_j = 1/log2(j+2) (the SLOT of the item in the list, not its current rank - kept as the reference has it),
// gains g = 2^t - 1 and IDCG over the valid slots of the label-sorted list,
//   Delta_ij = |g_i d_j + g_j d_i - g_i d_i - g_j d_j| / IDCG,   Rho_ij = 1 / (1 + exp(s_i - s_j)),
//   Lambda_i = sum_{j: t_i > t_j} Delta_ij Rho_ij  -  sum_{j: t_i < t_j} Delta_ji Rho_ji      (valid i, j only).
// A session without a positive item has IDCG = 0 and the reference's 0/0 turns its whole row into NaN: kept.
// One warp per session; s and g of the session plus the block's discount table live in shared memory.
static const int LR_WARPS = 4;

__global__ void __launch_bounds__(LR_WARPS * 32) lambdarank_kernel(int64_t B, int64_t L, const int64_t* __restrict__ ranking,
                                                                   const float* __restrict__ scores,
                                                                   const int64_t* __restrict__ lens, float* __restrict__ lambdas) {
    DYN_SMEM(float, sm);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * LR_WARPS + w;
    float* disc = sm;                                   // [L], shared by the block's warps
    float* s = sm + L + (size_t)w * 2 * L;              // [L] scores of this warp's session
    float* g = s + L;                                   // [L] gains 2^t - 1 (0 on pad slots)
    for (int64_t j = threadIdx.x; j < L; j += LR_WARPS * 32) disc[j] = 1.0f / log2f((float)j + 2.0f);
    int64_t n = 0;
    if (b < B) {
        n = lens[b];
        if (n > L) n = L;
        for (int64_t j = lane; j < L; j += 32) {
            int64_t t = ranking[b * L + j];
            t = t > 0 ? (t < 62 ? t : 62) : 0;
            s[j] = scores[b * L + j];
            g[j] = (float)(((int64_t)1 << t) - 1);
        }
    }
    __syncthreads();
    if (b >= B) return;
    // IDCG: slot of item i in the label-sorted list = #{t_j > t_i} + #{j < i, t_j == t_i}; slots >= n are masked
    double idcg_part = 0.0;
    for (int64_t i = lane; i < L; i += 32) {
        const float gi = g[i];
        if (gi <= 0.f) continue;
        int64_t pos = 0;
        for (int64_t j = 0; j < L; ++j) pos += (g[j] > gi) || (g[j] == gi && j < i);
        if (pos < n) idcg_part += (double)(gi * disc[pos]);
    }
    const float idcg = (float)warp_sum_d(idcg_part);
    if (!(idcg > 0.f)) {
        for (int64_t i = lane; i < L; i += 32) lambdas[b * L + i] = NAN;
        return;
    }
    for (int64_t i = lane; i < L; i += 32) {
        float up = 0.f, down = 0.f;
        if (i < n) {
            const float gi = g[i], si = s[i], di = disc[i];
            const float single_i = __fmul_rn(gi, di);
            for (int64_t j = 0; j < n; ++j) {
                const float gj = g[j];
                if (gj == gi) continue;
                const float dj = disc[j];
                // (pair_ij + pair_ji - single_i - single_j) in the reference's order of operations
                const float delta = fabsf(__fsub_rn(__fsub_rn(__fadd_rn(__fmul_rn(gi, dj), __fmul_rn(gj, di)), single_i),
                                                    __fmul_rn(gj, dj))) / idcg;
                if (gi > gj) up += delta * (1.0f / (1.0f + expf(si - s[j])));
                else down += delta * (1.0f / (1.0f + expf(s[j] - si)));
            }
        }
        lambdas[b * L + i] = up - down;
    }
}

int lambdarank_lambdas(int64_t B, int64_t L, const int64_t* ranking, const float* scores, const int64_t* lens, float* lambdas,
                       cudaStream_t st) {
    if (B <= 0 || L <= 0) return INTEL_OK;
    INTEL_REQUIRE(ranking && scores && lens && lambdas, INTEL_ERR_ARG, "lambdarank: null pointer");
    const size_t smem = (size_t)(1 + 2 * LR_WARPS) * L * 4;
    INTEL_REQUIRE(smem <= 200 * 1024, INTEL_ERR_UNSUPPORTED, "lambdarank: list length %lld too long", (long long)L);
    ensure_smem(lambdarank_kernel, smem);
    LAUNCH(lambdarank_kernel, dim3((unsigned)ceil_div(B, LR_WARPS)), dim3(LR_WARPS * 32), smem, st, B, L, ranking, scores, lens,
           lambdas);
    return check_launch("lambdarank", (double)B * L * 16.0 + (double)B * 8.0, 0.0);
}

}  // namespace intel

using namespace intel;

extern "C" {

int intel_loss_pl_fwd_bwd(int64_t B, int64_t L, int64_t K, const float* ens, const float* weights,
                          const double* scores, const int64_t* ranking, const int64_t* session_len, int cal_diversity,
                          double alpha, double* loss_out, float* d_ens, float* d_weights, intel_stream_t stream) {
    LossArgs a{B, L, (int)K, ens, weights, scores, ranking, session_len, cal_diversity, (float)alpha, loss_out, d_ens,
               cal_diversity ? d_weights : nullptr, nullptr, 0};
    return launch_loss(0, a, (cudaStream_t)stream);
}

int intel_loss_bpr_fwd_bwd(int64_t B, int64_t L, int64_t K, const float* ens, const float* weights,
                           const double* scores, const int64_t* ranking, const int64_t* session_len, const float* noise,
                           uint64_t seed, int cal_diversity, double alpha, double* loss_out, float* d_ens,
                           float* d_weights, intel_stream_t stream) {
    LossArgs a{B, L, (int)K, ens, weights, scores, ranking, session_len, cal_diversity, (float)alpha, loss_out, d_ens,
               cal_diversity ? d_weights : nullptr, noise, seed};
    return launch_loss(1, a, (cudaStream_t)stream);
}

int intel_loss_mse_fwd_bwd(int64_t B, int64_t L, int64_t K, const float* ens, const float* weights,
                           const double* scores, const int64_t* ranking, const int64_t* session_len, int cal_diversity,
                           double alpha, double* loss_out, float* d_ens, float* d_weights, intel_stream_t stream) {
    LossArgs a{B, L, (int)K, ens, weights, scores, ranking, session_len, cal_diversity, (float)alpha, loss_out, d_ens,
               cal_diversity ? d_weights : nullptr, nullptr, 0};
    return launch_loss(2, a, (cudaStream_t)stream);
}

int intel_intent_loss_fwd_bwd(int64_t B, int64_t I, const float* pred, const double* true_intents, double kl_weight,
                              double kl_temp, double* out, float* d_pred, void* scratch, intel_stream_t stream) {
    if (B <= 0) return INTEL_OK;
    INTEL_REQUIRE(pred && true_intents && out && d_pred && scratch, INTEL_ERR_ARG, "intent_loss: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    INTEL_TRY(fill_zero(out, 3 * sizeof(double), s));
    // 0x7f7f7f7f is a large positive int: a valid +inf-like start for the ordered-key minimum
    cudaError_t e = cudaMemsetAsync(scratch, 0x7f, sizeof(int), s);
    INTEL_REQUIRE(e == cudaSuccess, INTEL_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    unsigned g1 = stream_grid(ceil_div(B * I, 256 * 4), 4);
    LAUNCH(intent_min_kernel, dim3(g1), dim3(256), 0, s, B * I, pred, (int*)scratch);
    INTEL_TRY(check_launch("intent_min"));
    unsigned g2 = stream_grid(ceil_div(B, 8), 8);
    LAUNCH(intent_loss_kernel, dim3(g2), dim3(256), 0, s, B, I, pred, true_intents, (float)kl_weight,
           (float)(kl_temp * kl_temp), (const int*)scratch, out, d_pred);
    return check_launch("intent_loss", (double)B * I * 16.0, 0.0);
}

int intel_scale_by_device_scalar(int64_t n, const float* x, const double* a, double ca, const double* b, double cb,
                                 float* y, intel_stream_t stream) {
    if (n <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(n, 256 * 4), 8);
    LAUNCH(scale_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, n, x, a, ca, b, cb, y);
    return check_launch("scale");
}

int intel_lambdarank_lambdas(int64_t B, int64_t L, const int64_t* ranking, const float* scores, const int64_t* session_len,
                             float* lambdas, intel_stream_t stream) {
    return lambdarank_lambdas(B, L, ranking, scores, session_len, lambdas, (cudaStream_t)stream);
}

}  // extern "C"
