// Segmented per-session top-k + NDCG@k / HR@k evaluation (BaseRunner.evaluate_method / evaluate_intents)
// and the fixed-weight list-fusion baselines.  HBM-bound: every session is read once (4 B score + 8 B
// ranking per candidate), one warp per session, top-k by k rounds of a warp arg-max (k <= 64).
//
// Tie rule (the reference sorts with numpy's unstable argsort, so exact ties are undefined there): the
// literal operations of BaseRunner.py:66-126 with every argsort made stable -
//   column order  = ranking descending, original index descending; zero-score pad columns (ranking -2) last
//   sorted order  = score descending; among equal scores the later column first, i.e. lower raw ranking
//                   first (pads before real items), then lower original index.
#include "kernels.h"
#include "../../include/intel_b200.h"

namespace intel {

static const int EV_WARPS = 4;
static const int EV_MAX_K = 64;       // largest k in topk
static const int EV_OUT = 7;          // per k: ndcg, pay_hr, pay_ndcg, fav_hr, fav_ndcg, click_hr, click_ndcg

struct TopkList { int k[INTEL_MAX_TOPK]; int n; int kmax; };

__device__ __forceinline__ double disc_at(int p) { return 1.0 / log2((double)p + 2.0); }

// Deterministic block reduction of per-warp partial sums: warp w of block blk wrote part[w][c];
// thread c sums the EV_WARPS rows in fixed order and stores the block partial.
// Deterministic reduction of the per-block partial rows [nblocks][stride]: 32 row groups run in parallel (group g sums
// rows g, g + 32, ... in order, coalesced 8-byte loads), then the 32 group sums of a column are added in a fixed order.
// Columns [0, ncols) go to `out`, columns [ncols, ncols + ncols2) to `out2` (nullable): one launch for sums and counts.
static const int RP_GROUPS = 32;
__global__ void __launch_bounds__(1024) reduce_partials_kernel(int64_t nblocks, int stride, int ncols, const double* __restrict__ part,
                                                               double* __restrict__ out, int ncols2, double* __restrict__ out2) {
    __shared__ double grp[RP_GROUPS][33];
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = ncols + (out2 ? ncols2 : 0);
    for (int c0 = 0; c0 < total; c0 += 32) {
        const int c = c0 + lane;
        double acc = 0.0;
        if (c < total)
            for (int64_t i = g; i < nblocks; i += RP_GROUPS) acc += part[i * stride + c];
        grp[g][lane] = acc;
        __syncthreads();
        if (g == 0 && c < total) {
            double t = 0.0;
#pragma unroll 4
            for (int k = 0; k < RP_GROUPS; ++k) t += grp[k][lane];
            if (c < ncols) out[c] = t;
            else out2[c - ncols] = t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(EV_WARPS * 32) ndcg_kernel(int64_t N, int64_t ld, const float* __restrict__ pred,
                                                             const int64_t* __restrict__ ranking,
                                                             const int64_t* __restrict__ lens,
                                                             const int64_t* __restrict__ pay,
                                                             const int64_t* __restrict__ fav,
                                                             const int64_t* __restrict__ click, int64_t max_len,
                                                             TopkList tk, double* __restrict__ partial) {
    DYN_SMEM(float, sm);
    __shared__ double wsum[EV_WARPS][INTEL_MAX_TOPK * EV_OUT + 4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ncols = tk.n * EV_OUT + 4;
    float* sc = sm + (size_t)w * 2 * ld;                 // scores; taken items are overwritten with NaN
    int* rk = reinterpret_cast<int*>(sc + ld);           // raw ranking
    for (int c = lane; c < ncols; c += 32) wsum[w][c] = 0.0;
    __syncwarp();
    const int64_t sessions_per_block = EV_WARPS;
    for (int64_t base = (int64_t)blockIdx.x * sessions_per_block; base < N; base += (int64_t)gridDim.x * sessions_per_block) {
        const int64_t sidx = base + w;
        if (sidx >= N) continue;
        int64_t n = lens[sidx];
        if (n > ld) n = ld;
        const int64_t npad = max_len - n;
        for (int64_t j = lane; j < n; j += 32) {
            sc[j] = pred[sidx * ld + j];
            rk[j] = (int)ranking[sidx * ld + j];
        }
        __syncwarp();
        const int64_t apos[3] = {pay[sidx], fav[sidx], pay[sidx] + fav[sidx] + click[sidx]};
        // ---- top-kmax by score ----
        int gain[EV_MAX_K];          // per sorted position (uniform across the warp)
        unsigned inpos[3] = {0u, 0u, 0u};   // bit p: item at sorted position p < 32 is inside the first all_pos columns
        unsigned inpos_hi[3] = {0u, 0u, 0u};
        int64_t pads_used = 0;
        const int kmax = tk.kmax;
#pragma unroll 1
        for (int p = 0; p < kmax; ++p) {
            // best remaining real item: score desc, raw ranking asc, index asc
            float bs = -INFINITY;
            int br = 0x7fffffff;
            int64_t bj = -1;
            for (int64_t j = lane; j < n; j += 32) {
                const float v = sc[j];
                if (v != v) continue;                       // taken
                const int rj = rk[j];
                if (bj < 0 || v > bs || (v == bs && (rj < br || (rj == br && j < bj)))) { bs = v; br = rj; bj = j; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, bs, o);
                const int orr = __shfl_xor_sync(0xffffffffu, br, o);
                const int64_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (oj >= 0 && (bj < 0 || os > bs || (os == bs && (orr < br || (orr == br && oj < bj))))) { bs = os; br = orr; bj = oj; }
            }
            const bool pad_left = pads_used < npad;
            int g = 0;
            int64_t col = -1;
            if (bj >= 0 && !(pad_left && bs <= 0.f)) {
                // a real item wins this position; its column in the ranking-descending order
                int cnt = 0;
                for (int64_t j = lane; j < n; j += 32) cnt += (rk[j] > br) || (rk[j] == br && j > bj);
                col = warp_sum_i(cnt);
                g = br > 0 ? br : 0;
                if (lane == 0) sc[bj] = NAN;
                __syncwarp();
            } else if (pad_left) {
                col = max_len - 1 - pads_used;              // pads: later column first
                pads_used++;
            }
            if (p < EV_MAX_K) gain[p] = g;
            if (col >= 0) {
#pragma unroll
                for (int x = 0; x < 3; ++x) {
                    if (col < apos[x]) { if (p < 32) inpos[x] |= 1u << p; else inpos_hi[x] |= 1u << (p - 32); }
                }
            }
        }
        // ---- ideal ordering of the gains (only the multiset matters) ----
        int ideal[EV_MAX_K];
        {
            int filled = 0, prev = 0x7fffffff;
            while (filled < kmax) {
                int g = 0;
                for (int64_t j = lane; j < n; j += 32) { const int v = rk[j]; if (v < prev && v > g) g = v; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, g, o); g = t > g ? t : g; }
                if (g <= 0) break;
                int c = 0;
                for (int64_t j = lane; j < n; j += 32) c += (rk[j] == g);
                c = warp_sum_i(c);
                for (int q = 0; q < c && filled < kmax; ++q) ideal[filled++] = g;
                prev = g;
            }
            for (; filled < kmax; ++filled) ideal[filled] = 0;
        }
        if (lane == 0) {
            for (int t = 0; t < tk.n; ++t) {
                const int k = tk.k[t];
                double dcg = 0.0, idcg = 0.0;
                for (int p = 0; p < k; ++p) { dcg += gain[p] * disc_at(p); idcg += ideal[p] * disc_at(p); }
                wsum[w][t * EV_OUT + 0] += dcg / idcg;               // 0/0 -> NaN like the reference
                const int mk = (int64_t)k < max_len ? k : (int)max_len;
                for (int x = 0; x < 3; ++x) {
                    if (apos[x] <= 0) continue;
                    double bd = 0.0, bi = 0.0;
                    bool hit = false;
                    for (int p = 0; p < mk; ++p) {
                        const bool in = p < 32 ? ((inpos[x] >> p) & 1u) : ((inpos_hi[x] >> (p - 32)) & 1u);
                        if (in) { bd += disc_at(p); hit = true; }
                        if ((int64_t)p < apos[x]) bi += disc_at(p);
                    }
                    wsum[w][t * EV_OUT + 1 + 2 * x] += hit ? 1.0 : 0.0;
                    wsum[w][t * EV_OUT + 2 + 2 * x] += bd / bi;
                }
            }
            wsum[w][tk.n * EV_OUT + 0] += 1.0;
            for (int x = 0; x < 3; ++x) wsum[w][tk.n * EV_OUT + 1 + x] += apos[x] > 0 ? 1.0 : 0.0;
        }
        __syncwarp();
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < EV_WARPS; ++i) acc += wsum[i][c];
        partial[(int64_t)blockIdx.x * ncols + c] = acc;
    }
}

static unsigned ndcg_grid(int64_t N) { return stream_grid(ceil_div(N, EV_WARPS), 8); }

// ------------------------------------------------------------------------------------------------
// evaluate_intents: one warp per session over the I intent classes.
__global__ void __launch_bounds__(EV_WARPS * 32) intent_topk_kernel(int64_t N, int64_t I,
                                                                    const double* __restrict__ truth,
                                                                    const float* __restrict__ pred, TopkList tk,
                                                                    double* __restrict__ partial) {
    DYN_SMEM(double, smd);
    __shared__ double wsum[EV_WARPS][INTEL_MAX_TOPK * 2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ncols = tk.n * 2;
    double* tv = smd + (size_t)w * 2 * I;        // true values (a copy that the ideal pass consumes)
    double* pv = tv + I;                         // predictions (as double; taken -> NaN)
    for (int c = lane; c < ncols; c += 32) wsum[w][c] = 0.0;
    __syncwarp();
    for (int64_t base = (int64_t)blockIdx.x * EV_WARPS; base < N; base += (int64_t)gridDim.x * EV_WARPS) {
        const int64_t sidx = base + w;
        if (sidx >= N) continue;
        // label = first arg-max of the true distribution (np.argmax)
        double bt = -INFINITY;
        int64_t bl = I;
        for (int64_t c = lane; c < I; c += 32) {
            const double t = truth[sidx * I + c];
            tv[c] = t;
            pv[c] = (double)pred[sidx * I + c];
            if (t > bt) { bt = t; bl = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ot = __shfl_xor_sync(0xffffffffu, bt, o);
            const int64_t ol = __shfl_xor_sync(0xffffffffu, bl, o);
            if (ot > bt || (ot == bt && ol < bl)) { bt = ot; bl = ol; }
        }
        __syncwarp();
        const int kmax = tk.kmax < (int)I ? tk.kmax : (int)I;
        double got[EV_MAX_K], ideal[EV_MAX_K];
        int label_pos = 0x7fffffff;
#pragma unroll 1
        for (int p = 0; p < kmax; ++p) {
            // predicted order: value desc, ties -> larger index first (stable ascending sort reversed)
            double bs = -INFINITY;
            int64_t bj = -1;
            for (int64_t c = lane; c < I; c += 32) {
                const double v = pv[c];
                if (v != v) continue;
                if (bj < 0 || v > bs || (v == bs && c > bj)) { bs = v; bj = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double os = __shfl_xor_sync(0xffffffffu, bs, o);
                const int64_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (oj >= 0 && (bj < 0 || os > bs || (os == bs && oj > bj))) { bs = os; bj = oj; }
            }
            got[p] = bj >= 0 ? truth[sidx * I + bj] : 0.0;
            if (bj == bl) label_pos = p;
            if (lane == 0 && bj >= 0) pv[bj] = NAN;
            __syncwarp();
            // ideal order of the true values
            double bi = -INFINITY;
            int64_t bk = -1;
            for (int64_t c = lane; c < I; c += 32) {
                const double v = tv[c];
                if (v != v) continue;
                if (bk < 0 || v > bi || (v == bi && c < bk)) { bi = v; bk = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const int64_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
                if (ok >= 0 && (bk < 0 || oi > bi || (oi == bi && ok < bk))) { bi = oi; bk = ok; }
            }
            ideal[p] = bk >= 0 ? bi : 0.0;
            if (lane == 0 && bk >= 0) tv[bk] = NAN;
            __syncwarp();
        }
        if (lane == 0) {
            for (int t = 0; t < tk.n; ++t) {
                const int k = tk.k[t] < kmax ? tk.k[t] : kmax;
                double dcg = 0.0, idcg = 0.0;
                for (int p = 0; p < k; ++p) { dcg += got[p] * disc_at(p); idcg += ideal[p] * disc_at(p); }
                wsum[w][2 * t] += dcg / idcg;
                wsum[w][2 * t + 1] += label_pos < k ? 1.0 : 0.0;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < EV_WARPS; ++i) acc += wsum[i][c];
        partial[(int64_t)blockIdx.x * ncols + c] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Borda (Borda.py:23-30): ascending rank of every slot inside each basic list over the padded length
// (stable: ties by index), mean over the K lists.  One warp per session, lists staged in shared memory.
__global__ void __launch_bounds__(EV_WARPS * 32) rank_lists_kernel(int64_t B, int64_t L, int K,
                                                                   const double* __restrict__ scores,
                                                                   float* __restrict__ ens) {
    DYN_SMEM(float, sm);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * EV_WARPS + w;
    if (b >= B) return;
    float* x = sm + (size_t)w * L * K;
    for (int64_t e = lane; e < L * K; e += 32) x[e] = (float)scores[b * L * K + e];
    __syncwarp();
    const float wk = 1.0f / (float)K;
    for (int64_t l = lane; l < L; l += 32) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            const float v = x[l * K + k];
            int rank = 0;
            for (int64_t j = 0; j < L; ++j) {
                const float u = x[j * K + k];
                rank += (u < v) || (u == v && j < l);
            }
            acc = __fadd_rn(acc, __fmul_rn(wk, (float)rank));   // torch: (w * rank).sum(2), no fma contraction
        }
        ens[b * L + l] = acc;
    }
}

__global__ void __launch_bounds__(256) select_list_kernel(int64_t R, int K, const double* __restrict__ scores, int column,
                                                          float* __restrict__ ens) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x)
        ens[r] = (float)scores[r * K + column];
}

static int make_topk(const int32_t* topk, int n_topk, TopkList& tk) {
    INTEL_REQUIRE(topk && n_topk > 0 && n_topk <= INTEL_MAX_TOPK, INTEL_ERR_ARG, "topk: need 1..%d entries", INTEL_MAX_TOPK);
    tk.n = n_topk;
    tk.kmax = 0;
    for (int i = 0; i < n_topk; ++i) {
        INTEL_REQUIRE(topk[i] >= 1 && topk[i] <= EV_MAX_K, INTEL_ERR_UNSUPPORTED, "topk value %d outside 1..%d", topk[i], EV_MAX_K);
        tk.k[i] = topk[i];
        if (topk[i] > tk.kmax) tk.kmax = topk[i];
    }
    for (int i = n_topk; i < INTEL_MAX_TOPK; ++i) tk.k[i] = 0;
    return INTEL_OK;
}

}  // namespace intel

using namespace intel;

extern "C" {

size_t intel_ndcg_workspace_bytes(int64_t N, int n_topk) {
    return (size_t)ndcg_grid(N) * (size_t)(n_topk * EV_OUT + 4) * sizeof(double) + 256;
}

int intel_ndcg_topk(int64_t N, int64_t ld, const float* pred, const int64_t* ranking, const int64_t* session_len,
                    const int64_t* pay, const int64_t* fav, const int64_t* click, int64_t max_len, const int32_t* topk,
                    int n_topk, double* sums, double* counts, void* workspace, size_t workspace_bytes,
                    intel_stream_t stream) {
    if (N <= 0) return INTEL_OK;
    TopkList tk;
    INTEL_TRY(make_topk(topk, n_topk, tk));
    INTEL_REQUIRE(pred && ranking && session_len && pay && fav && click && sums && counts, INTEL_ERR_ARG, "ndcg: null pointer");
    INTEL_REQUIRE(max_len >= tk.kmax, INTEL_ERR_ARG, "ndcg: max_len %lld < max(topk) %d", (long long)max_len, tk.kmax);
    INTEL_REQUIRE(workspace && workspace_bytes >= intel_ndcg_workspace_bytes(N, n_topk), INTEL_ERR_WORKSPACE, "ndcg: workspace too small");
    const size_t smem = (size_t)EV_WARPS * 2 * ld * 4;
    INTEL_REQUIRE(smem <= 200 * 1024, INTEL_ERR_UNSUPPORTED, "ndcg: row length %lld too long", (long long)ld);
    cudaStream_t s = (cudaStream_t)stream;
    ensure_smem(ndcg_kernel, smem);
    const unsigned grid = ndcg_grid(N);
    const int ncols = n_topk * EV_OUT + 4;
    double* partial = reinterpret_cast<double*>(workspace);
    LAUNCH(ndcg_kernel, dim3(grid), dim3(EV_WARPS * 32), smem, s, N, ld, pred, ranking, session_len, pay, fav, click,
           max_len, tk, partial);
    INTEL_TRY(check_launch("ndcg", (double)N * ld * 12.0, 0.0));
    // fixed-order (deterministic) reduction of the per-block partial rows
    LAUNCH(reduce_partials_kernel, dim3(1), dim3(1024), 0, s, (int64_t)grid, ncols, n_topk * EV_OUT, partial, sums, 4, counts);
    return check_launch("ndcg_reduce", (double)grid * ncols * 8.0, 0.0);
}

size_t intel_intent_topk_workspace_bytes(int64_t N, int n_topk) {
    return (size_t)ndcg_grid(N) * (size_t)(n_topk * 2) * sizeof(double) + 256;
}

int intel_intent_topk(int64_t N, int64_t I, const double* true_intents, const float* pred_intents, const int32_t* topk,
                      int n_topk, double* sums, void* workspace, size_t workspace_bytes, intel_stream_t stream) {
    if (N <= 0) return INTEL_OK;
    TopkList tk;
    INTEL_TRY(make_topk(topk, n_topk, tk));
    INTEL_REQUIRE(true_intents && pred_intents && sums, INTEL_ERR_ARG, "intent_topk: null pointer");
    INTEL_REQUIRE(workspace && workspace_bytes >= intel_intent_topk_workspace_bytes(N, n_topk), INTEL_ERR_WORKSPACE,
                  "intent_topk: workspace too small");
    const size_t smem = (size_t)EV_WARPS * 2 * I * 8;
    INTEL_REQUIRE(smem <= 200 * 1024, INTEL_ERR_UNSUPPORTED, "intent_topk: intent_num %lld too large", (long long)I);
    cudaStream_t s = (cudaStream_t)stream;
    ensure_smem(intent_topk_kernel, smem);
    const unsigned grid = ndcg_grid(N);
    double* partial = reinterpret_cast<double*>(workspace);
    LAUNCH(intent_topk_kernel, dim3(grid), dim3(EV_WARPS * 32), smem, s, N, I, true_intents, pred_intents, tk, partial);
    INTEL_TRY(check_launch("intent_topk"));
    LAUNCH(reduce_partials_kernel, dim3(1), dim3(1024), 0, s, (int64_t)grid, n_topk * 2, n_topk * 2, partial, sums, 0,
           (double*)nullptr);
    return check_launch("intent_topk_reduce");
}

int intel_fuse_fwd(int64_t B, int64_t L, int64_t K, const float* weights, const double* scores, float* ens_out,
                   intel_stream_t stream) {
    INTEL_REQUIRE(weights && scores && ens_out, INTEL_ERR_ARG, "fuse: null pointer");
    return item_fuse_fwd(B * L, (int)K, weights, scores, ens_out, (cudaStream_t)stream);
}

int intel_select_list(int64_t B, int64_t L, int64_t K, const double* scores, int column, float* ens_out,
                      intel_stream_t stream) {
    if (B * L <= 0) return INTEL_OK;
    INTEL_REQUIRE(scores && ens_out && column >= 0 && column < K, INTEL_ERR_ARG, "select_list: bad column %d", column);
    unsigned grid = stream_grid(ceil_div(B * L, 256), 8);
    LAUNCH(select_list_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, B * L, (int)K, scores, column, ens_out);
    return check_launch("select_list");
}

int intel_rank_lists(int64_t B, int64_t L, int64_t K, const double* scores, float* ens_out, intel_stream_t stream) {
    if (B * L <= 0) return INTEL_OK;
    INTEL_REQUIRE(scores && ens_out, INTEL_ERR_ARG, "rank_lists: null pointer");
    const size_t smem = (size_t)EV_WARPS * L * K * 4;
    INTEL_REQUIRE(smem <= 200 * 1024, INTEL_ERR_UNSUPPORTED, "rank_lists: list too long");
    ensure_smem(rank_lists_kernel, smem);
    LAUNCH(rank_lists_kernel, dim3((unsigned)ceil_div(B, EV_WARPS)), dim3(EV_WARPS * 32), smem, (cudaStream_t)stream, B, L,
           (int)K, scores, ens_out);
    return check_launch("rank_lists");
}

}  // extern "C"
