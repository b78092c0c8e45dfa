// Device-side batch construction: what BaseModel.Dataset._get_feed_dict + collate_batch do per session on
// the host (BaseModel.py:121-197, GeneralSeq.py:35-54, IntEL.py:220-239), as one gather kernel over a columnar corpus that
// lives in HBM.  One CTA per session of the batch: per-session scalars, the (optionally permuted) item list with its
// classes / rankings / normalised scores right-padded with zeros to the batch's L, the session history (context ids +
// intent vectors in the compact index / value form) and the item history (ids + one-hot intent index), the dense
// float64 true-intent row the intent loss reads.  Pure HBM streaming: ~ (24 + 8 K) L + 12 nz H + 8 I bytes per session.
#include "kernels.h"
#include "intel_b200.h"

namespace intel {

__global__ void __launch_bounds__(128) batch_build_kernel(intel_corpus_t c, int64_t B, const int64_t* __restrict__ rows,
                                                          const int32_t* __restrict__ perm, int64_t L, int64_t H1, int64_t H2,
                                                          intel_built_batch_t o) {
    const int64_t b = blockIdx.x;
    if (b >= B) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t row = rows[b];
    const int64_t n = c.session_len[row], uid = c.u_id[row];
    const int64_t pos = c.position[row], ipos = c.item_position[row];
    // histories: the last max_his entries before this session; none -> one all-zero row (GeneralSeq.py:50-52, IntEL.py:236-238)
    const int64_t h1 = pos > 0 ? (c.max_his > 0 && pos > c.max_his ? c.max_his : pos) : 1;
    const int64_t h2 = ipos > 0 ? (c.max_his > 0 && ipos > c.max_his ? c.max_his : ipos) : 1;
    if (tid == 0) {
        o.u_id[b] = uid; o.c_id[b] = c.c_id[row]; o.context_mh[b] = c.context_mh[row]; o.user_mh[b] = c.user_mh[row];
        o.pay[b] = c.pay[row]; o.fav[b] = c.fav[row]; o.click[b] = c.click[row];
        o.session_len[b] = n; o.position[b] = pos; o.history_len[b] = h1; o.history_item_len[b] = h2;
    }
    // ---- the list: slot l of the batch row takes slot perm[b, l] of the stored list ----
    const int64_t base = c.item_off[row];
    const int K = (int)c.K;
    for (int64_t l = tid; l < L; l += nt) {
        const bool live = l < n;
        int64_t src = live ? (perm ? (int64_t)perm[b * L + l] : l) : 0;
        if (src < 0 || src >= n) src = live ? l : 0;
        const int64_t e = base + src;
        o.i_id[b * L + l] = live ? c.item_id[e] : 0;
        o.i_class[b * L + l] = live ? c.item_class[e] : 0;
        o.ranking[b * L + l] = live ? c.ranking[e] : 0;
        for (int k = 0; k < K; ++k) o.scores[(b * L + l) * K + k] = live ? c.scores[e * K + k] : 0.0;
    }
    // ---- session history ----
    const int nz = c.nz1;
    const int64_t ub = c.uhis_off[uid] + (pos > 0 ? pos - h1 : 0);
    for (int64_t h = tid; h < H1; h += nt) o.his_context[b * H1 + h] = (pos > 0 && h < h1) ? c.uhis_ctx[ub + h] : 0;
    for (int64_t e = tid; e < H1 * nz; e += nt) {
        const int64_t h = e / nz;
        const int j = (int)(e % nz);
        int32_t ix = 0;
        float v = 0.f;
        if (pos > 0 && h < h1) {
            const int64_t r = c.uhis_row[ub + h];
            const int64_t s0 = c.int_off[r], cnt = c.int_off[r + 1] - s0;
            if (j < cnt) { ix = c.int_idx[s0 + j]; v = (float)c.int_val[s0 + j]; }
        }
        o.his_intents_idx[(b * H1) * nz + e] = ix;
        o.his_intents_val[(b * H1) * nz + e] = v;
    }
    // ---- item history ----
    const int64_t ib = c.uitem_off[uid] + (ipos > 0 ? ipos - h2 : 0);
    for (int64_t h = tid; h < H2; h += nt) {
        const bool on = ipos > 0 && h < h2;
        o.his_item_id[b * H2 + h] = on ? c.uitem_id[ib + h] : 0;
        o.his_item_int_idx[b * H2 + h] = on ? c.uitem_int[ib + h] : 0;
        o.his_item_int_val[b * H2 + h] = on ? 1.0f : 0.0f;
    }
    // ---- dense true intents (corpus.intents.get(c_id_c, zero_int), BaseModel.py:175) ----
    double* dst = o.intents + b * c.I;
    for (int64_t i = tid; i < c.I; i += nt) dst[i] = 0.0;
    __syncthreads();
    const int64_t r = c.intent_row[row];
    const int64_t s0 = c.int_off[r], cnt = c.int_off[r + 1] - s0;
    for (int64_t j = tid; j < cnt; j += nt) dst[c.int_idx[s0 + j]] = c.int_val[s0 + j];
}

}  // namespace intel

extern "C" int intel_batch_build(const intel_corpus_t* corpus, int64_t B, const int64_t* rows, const int32_t* perm, int64_t L,
                                 int64_t H1, int64_t H2, const intel_built_batch_t* out, intel_stream_t stream) {
    using namespace intel;
    INTEL_REQUIRE(corpus && rows && out, INTEL_ERR_ARG, "intel_batch_build: null argument");
    INTEL_REQUIRE(L >= 1 && H1 >= 1 && H2 >= 1 && corpus->nz1 >= 1, INTEL_ERR_ARG, "intel_batch_build: empty shape");
    if (B <= 0) return INTEL_OK;
    cudaStream_t s = (cudaStream_t)stream;
    LAUNCH(batch_build_kernel, dim3((unsigned)B), dim3(128), 0, s, *corpus, B, rows, perm, L, H1, H2, *out);
    const double bytes = (double)B * ((32.0 + 16.0 * corpus->K) * L + 12.0 * corpus->nz1 * H1 + 24.0 * (H1 + H2) + 8.0 * corpus->I + 96.0);
    return check_launch("batch_build", bytes, 0.0);
}
