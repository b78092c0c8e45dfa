// BERT4RecEncoder forward (GeneralSeq.py:80-106, layers.py:62-88) in one kernel for the width the scripts use (d = 32,
// history of at most 24 slots): learned positions, N post-LN blocks (q|k|v projections with bias, key-masked multi-head
// attention, residual + LayerNorm, d -> d -> d feed-forward, residual + LayerNorm) and the pick of the state at len - 1.
// The staged path (api_model.cu: bert_fwd) runs the same math as ~25 launches per encoder over [B*T, 32] tensors, each of
// them at the floor of a dependent launch; this is what BASELINE.json configs[3] (eval only, default BERT4Rec sizes) spends
// 45 % of its step in.
//
// One warp per session, lane = channel.  Activations of the session live in shared memory in two layouts: channel-major
// [32][24] (token contiguous: a float4 is four tokens of one channel, read as a broadcast) for everything that is an A
// operand, token-major [24][33] for V.  A projection out[t][n] = b[n] + sum_k x[t][k] W[n][k] keeps all T outputs of lane n
// in registers: per k one shared-memory word of W and six broadcast float4 of x feed 24 FMAs.  Attention: lane = key for the
// scores (q broadcast, k channel-major), lane = channel for P V.  All arithmetic is fp32 FMA (the parity budget of the 3xTF32
// kernels is kept with margin).  In training mode the kernel leaves exactly the activations bert_bwd reads.
#include "kernels.h"

namespace intel {

namespace {
constexpr int BF_D = 32, BF_TMAX = 24;
constexpr int BF_WS = BF_D + 1;                 // weight row stride: lane n reads W[n][k], 33 n + k is conflict free
constexpr int BF_LAYER = 5 * BF_D * BF_WS + 9 * BF_D;      // qw kw vw l1w l2w | qb kb vb ln1w ln1b l1b l2b ln2w ln2b
// per warp: x | q | k channel-major tiles [32][TP] and v token-major [TP][33]; the probabilities of both heads ([2][TP][TP])
// take the place of q | k once the scores of both heads sit in registers (2 TP^2 <= 64 TP)
constexpr int bf_warp_floats(int TP) { return 3 * BF_D * TP + TP * BF_WS; }

struct BertFusedArgs {
    int64_t B;
    int T, layers, heads, save, vecw;            // vecw: every weight matrix is 16-byte aligned
    const int64_t* lens;
    const float* pos;
    float* seq;                                  // [B, T, 32]: token embeddings in, X[0] (positions added) out when save
    intel_bert_layer_t L[INTEL_MAX_BERT_LAYERS];
    float *X[INTEL_MAX_BERT_LAYERS + 1], *QKV[INTEL_MAX_BERT_LAYERS], *Z1[INTEL_MAX_BERT_LAYERS], *st1[INTEL_MAX_BERT_LAYERS],
        *C[INTEL_MAX_BERT_LAYERS], *F[INTEL_MAX_BERT_LAYERS], *Z2[INTEL_MAX_BERT_LAYERS], *st2[INTEL_MAX_BERT_LAYERS];
    float* out; int64_t ld_out;
};

// acc[t] = bias + sum_k x[k][t] W[lane][k] for t < 24 (x: channel-major tile, W: [32][33]); relu_in: x -> max(x, 0)
template <int TP>
__device__ __forceinline__ void project(float (&acc)[TP], const float* __restrict__ x, const float* __restrict__ W, float bias,
                                        int lane, bool relu_in) {
#pragma unroll
    for (int t = 0; t < TP; ++t) acc[t] = bias;
#pragma unroll 4
    for (int k = 0; k < BF_D; ++k) {
        const float w = W[lane * BF_WS + k];
#pragma unroll
        for (int g = 0; g < TP / 4; ++g) {
            float4 v = *reinterpret_cast<const float4*>(x + k * TP + 4 * g);
            if (relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            acc[4 * g] = fmaf(v.x, w, acc[4 * g]);
            acc[4 * g + 1] = fmaf(v.y, w, acc[4 * g + 1]);
            acc[4 * g + 2] = fmaf(v.z, w, acc[4 * g + 2]);
            acc[4 * g + 3] = fmaf(v.w, w, acc[4 * g + 3]);
        }
    }
}
// lane's 24 values -> row `lane` of a channel-major tile
template <int TP>
__device__ __forceinline__ void put_cm(float* tile, const float (&v)[TP], int lane) {
#pragma unroll
    for (int g = 0; g < TP / 4; ++g)
        *reinterpret_cast<float4*>(tile + lane * TP + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
}
// lane's values of tokens t < T -> column `lane` of a [B*T, ld] global tensor (coalesced: 32 lanes = 32 channels of a row)
template <int TP>
__device__ __forceinline__ void put_global(float* dst, int64_t ld, const float (&v)[TP], int T, int lane) {
#pragma unroll
    for (int t = 0; t < TP; ++t)
        if (t < T) dst[(int64_t)t * ld + lane] = v[t];
}
// z -> LayerNorm(z) per token (over the 32 lanes); stats (mean, rstd) to st[2 t] when given
template <int TP>
__device__ __forceinline__ void layer_norm(float (&z)[TP], float gamma, float beta, float* st, int T, int lane) {
#pragma unroll
    for (int t = 0; t < TP; ++t) {
        const float mean = warp_sum(z[t]) * (1.0f / BF_D);
        const float dlt = z[t] - mean;
        const float rstd = rsqrtf(warp_sum(dlt * dlt) * (1.0f / BF_D) + 1e-5f);
        if (st && lane == 0 && t < T) { st[2 * t] = mean; st[2 * t + 1] = rstd; }
        z[t] = dlt * rstd * gamma + beta;
    }
}
}  // namespace

template <int TP>
__global__ void __launch_bounds__(512) bert_fused_fwd_kernel(BertFusedArgs a) {
    DYN_SMEM(float, sm);
    constexpr int BF_CM = BF_D * TP;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int T = a.T, heads = a.heads, dk = BF_D / heads;
    float* Wsm = sm;                                                   // layers x BF_LAYER
    float* mine = sm + a.layers * BF_LAYER + w * bf_warp_floats(TP);
    float* xs = mine;                                                  // channel-major tiles
    float* qs = xs + BF_CM;
    float* ks = qs + BF_CM;
    float* vs = ks + BF_CM;                                            // token-major [TP][33]
    float* ps = qs;                                                    // [head][key][TP queries], over q | k (see above)
    // weights -> shared memory: one flat loop over (layer, matrix, 16-byte chunk) so that every thread has several
    // independent global loads in flight (ten short loops in sequence cost 15 % of the kernel: one latency each)
    {
        const int chunks = a.layers * 5 * (BF_D * BF_D / 4);
#pragma unroll 4
        for (int e = threadIdx.x; e < chunks; e += blockDim.x) {
            const int l = e / (5 * 256), m = (e / 256) % 5, c = e % 256;       // chunk c = row c / 8, columns 4 (c % 8) ..
            const intel_bert_layer_t& q = a.L[l];
            const float* src = m == 0 ? q.qw : (m == 1 ? q.kw : (m == 2 ? q.vw : (m == 3 ? q.l1w : q.l2w)));
            const float4 v = a.vecw ? *reinterpret_cast<const float4*>(src + 4 * c)
                                    : make_float4(src[4 * c], src[4 * c + 1], src[4 * c + 2], src[4 * c + 3]);
            float* dst = Wsm + l * BF_LAYER + m * BF_D * BF_WS + (c >> 3) * BF_WS + 4 * (c & 7);
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
        for (int e = threadIdx.x; e < a.layers * 9 * BF_D; e += blockDim.x) {
            const int l = e / (9 * BF_D), m = (e / BF_D) % 9, c = e % BF_D;
            const intel_bert_layer_t& q = a.L[l];
            const float* src = m == 0 ? q.qb : (m == 1 ? q.kb : (m == 2 ? q.vb : (m == 3 ? q.ln1w : (m == 4 ? q.ln1b : (m == 5 ? q.l1b :
                               (m == 6 ? q.l2b : (m == 7 ? q.ln2w : q.ln2b)))))));
            Wsm[l * BF_LAYER + 5 * BF_D * BF_WS + m * BF_D + c] = src[c];
        }
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)dk);
    const int wpb = blockDim.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    for (int64_t b = (int64_t)blockIdx.x * wpb + w; b < a.B; b += nwarps) {
        const int64_t len64 = a.lens[b];
        const int nk = (int)(len64 < 0 ? 0 : (len64 < T ? len64 : T));           // keys j >= len are masked (GeneralSeq.py:100-101)
        const int64_t row0 = b * T;
        float x[TP];                                                           // lane = channel, x[t]
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            x[t] = 0.f;
            if (t < T) x[t] = a.seq[(row0 + t) * BF_D + lane] + a.pos[(int64_t)((t < len64) ? t : 0) * BF_D + lane];
        }
        if (a.save) put_global(a.seq + row0 * BF_D, BF_D, x, T, lane);            // X[0] for the backward pass
        for (int l = 0; l < a.layers; ++l) {
            const float* Wl = Wsm + l * BF_LAYER;
            const float* vec = Wl + 5 * BF_D * BF_WS;
            put_cm(xs, x, lane);
            __syncwarp();
            float acc[TP];
            // ---- q | k | v ----
            project(acc, xs, Wl, vec[lane], lane, false);
            put_cm(qs, acc, lane);
            if (a.save) put_global(a.QKV[l] + row0 * 3 * BF_D, 3 * BF_D, acc, T, lane);
            project(acc, xs, Wl + BF_D * BF_WS, vec[BF_D + lane], lane, false);
            put_cm(ks, acc, lane);
            if (a.save) put_global(a.QKV[l] + row0 * 3 * BF_D + BF_D, 3 * BF_D, acc, T, lane);
            project(acc, xs, Wl + 2 * BF_D * BF_WS, vec[2 * BF_D + lane], lane, false);
#pragma unroll
            for (int t = 0; t < TP; ++t) vs[t * BF_WS + lane] = acc[t];
            if (a.save) put_global(a.QKV[l] + row0 * 3 * BF_D + 2 * BF_D, 3 * BF_D, acc, T, lane);
            __syncwarp();
            // ---- scores: lane = key j, the queries t of both heads in registers; softmax over the lanes j < nk.  The
            //      probabilities are written over the q | k tiles once every lane is done reading them ----
            {
                float s0[TP], s1[TP];
#pragma unroll
                for (int t = 0; t < TP; ++t) { s0[t] = 0.f; s1[t] = 0.f; }
                const int jj = lane < TP ? lane : 0;
                for (int c = 0; c < dk; ++c) {
                    const float kv = ks[c * TP + jj];
#pragma unroll
                    for (int g = 0; g < TP / 4; ++g) {
                        const float4 qv = *reinterpret_cast<const float4*>(qs + c * TP + 4 * g);
                        s0[4 * g] = fmaf(qv.x, kv, s0[4 * g]);
                        s0[4 * g + 1] = fmaf(qv.y, kv, s0[4 * g + 1]);
                        s0[4 * g + 2] = fmaf(qv.z, kv, s0[4 * g + 2]);
                        s0[4 * g + 3] = fmaf(qv.w, kv, s0[4 * g + 3]);
                    }
                }
                if (heads == 2) {
                    for (int c = dk; c < 2 * dk; ++c) {
                        const float kv = ks[c * TP + jj];
#pragma unroll
                        for (int g = 0; g < TP / 4; ++g) {
                            const float4 qv = *reinterpret_cast<const float4*>(qs + c * TP + 4 * g);
                            s1[4 * g] = fmaf(qv.x, kv, s1[4 * g]);
                            s1[4 * g + 1] = fmaf(qv.y, kv, s1[4 * g + 1]);
                            s1[4 * g + 2] = fmaf(qv.z, kv, s1[4 * g + 2]);
                            s1[4 * g + 3] = fmaf(qv.w, kv, s1[4 * g + 3]);
                        }
                    }
                }
                const bool on = lane < nk;
#pragma unroll
                for (int t = 0; t < TP; ++t) {
                    const float v = on ? s0[t] * scale : -INFINITY;
                    const float mx = warp_max(v);
                    const float e = on ? expf(v - mx) : 0.f;
                    const float sum = warp_sum(e);
                    s0[t] = (nk > 0 && t < T) ? e / sum : 0.f;
                }
                if (heads == 2) {
#pragma unroll
                    for (int t = 0; t < TP; ++t) {
                        const float v = on ? s1[t] * scale : -INFINITY;
                        const float mx = warp_max(v);
                        const float e = on ? expf(v - mx) : 0.f;
                        const float sum = warp_sum(e);
                        s1[t] = (nk > 0 && t < T) ? e / sum : 0.f;
                    }
                }
                __syncwarp();                                                      // every lane has read q | k
                if (lane < TP) {
#pragma unroll
                    for (int g = 0; g < TP / 4; ++g) {
                        *reinterpret_cast<float4*>(ps + lane * TP + 4 * g) = make_float4(s0[4 * g], s0[4 * g + 1], s0[4 * g + 2], s0[4 * g + 3]);
                        if (heads == 2)
                            *reinterpret_cast<float4*>(ps + (TP + lane) * TP + 4 * g) = make_float4(s1[4 * g], s1[4 * g + 1], s1[4 * g + 2], s1[4 * g + 3]);
                    }
                }
            }
            __syncwarp();
            // ---- O = P V (lane = channel, its head's probabilities), + residual, LayerNorm 1 ----
            {
                const float* ph = ps + (lane / dk) * TP * TP;
#pragma unroll
                for (int t = 0; t < TP; ++t) acc[t] = 0.f;
                for (int j = 0; j < nk; ++j) {
                    const float vv = vs[j * BF_WS + lane];
#pragma unroll
                    for (int g = 0; g < TP / 4; ++g) {
                        const float4 pv = *reinterpret_cast<const float4*>(ph + j * TP + 4 * g);
                        acc[4 * g] = fmaf(pv.x, vv, acc[4 * g]);
                        acc[4 * g + 1] = fmaf(pv.y, vv, acc[4 * g + 1]);
                        acc[4 * g + 2] = fmaf(pv.z, vv, acc[4 * g + 2]);
                        acc[4 * g + 3] = fmaf(pv.w, vv, acc[4 * g + 3]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < TP; ++t) acc[t] += x[t];
            if (a.save) put_global(a.Z1[l] + row0 * BF_D, BF_D, acc, T, lane);
            layer_norm(acc, vec[3 * BF_D + lane], vec[4 * BF_D + lane], a.save ? a.st1[l] + row0 * 2 : nullptr, T, lane);
            if (a.save) put_global(a.C[l] + row0 * BF_D, BF_D, acc, T, lane);
            // ---- feed-forward: F = C W1^T + b1, Z2 = relu(F) W2^T + b2 + C, LayerNorm 2 ----
            __syncwarp();
            put_cm(xs, acc, lane);                                                 // C (the x tile is free: x[] is in registers)
            __syncwarp();
            float f[TP];
            project(f, xs, Wl + 3 * BF_D * BF_WS, vec[5 * BF_D + lane], lane, false);
            if (a.save) put_global(a.F[l] + row0 * BF_D, BF_D, f, T, lane);
            put_cm(qs, f, lane);                                                   // the q | k tiles (P since the scores) are free
            __syncwarp();
            project(f, qs, Wl + 4 * BF_D * BF_WS, vec[6 * BF_D + lane], lane, true);
#pragma unroll
            for (int t = 0; t < TP; ++t) x[t] = f[t] + acc[t];
            if (a.save) put_global(a.Z2[l] + row0 * BF_D, BF_D, x, T, lane);
            layer_norm(x, vec[7 * BF_D + lane], vec[8 * BF_D + lane], a.save ? a.st2[l] + row0 * 2 : nullptr, T, lane);
            if (a.save) put_global(a.X[l + 1] + row0 * BF_D, BF_D, x, T, lane);
            __syncwarp();
        }
        // his_vector = state at len - 1 (GeneralSeq.py:105), clamped like take_last
        int tl = (int)(len64 - 1 < 0 ? 0 : (len64 - 1 >= T ? T - 1 : len64 - 1));
        float o = 0.f;
#pragma unroll
        for (int t = 0; t < TP; ++t) o = (t == tl) ? x[t] : o;
        a.out[b * a.ld_out + lane] = o;
    }
}

static int g_bert_fused = 1;
void bert_debug_use_fused(int on) { g_bert_fused = on ? 1 : 0; }

bool bert_fused_ok(int64_t T, int d, int heads, int layers) {
    return g_bert_fused && d == BF_D && T >= 1 && T <= BF_TMAX && (heads == 1 || heads == 2) && layers >= 1 && layers <= INTEL_MAX_BERT_LAYERS;
}

int bert_fused_fwd(int64_t B, int64_t T, int heads, int layers, const int64_t* lens, const intel_encoder_t& p, float* seq,
                   float* const* X, float* const* QKV, float* const* Z1, float* const* st1, float* const* C, float* const* F,
                   float* const* Z2, float* const* st2, bool save, float* out, int64_t ld_out, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    BertFusedArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.T = (int)T; a.layers = layers; a.heads = heads; a.save = save ? 1 : 0;
    a.lens = lens; a.pos = p.pos; a.seq = seq; a.out = out; a.ld_out = ld_out;
    a.vecw = 1;
    for (int l = 0; l < layers; ++l) {
        a.L[l] = p.layer[l];
        const float* ws[5] = {p.layer[l].qw, p.layer[l].kw, p.layer[l].vw, p.layer[l].l1w, p.layer[l].l2w};
        for (int m = 0; m < 5; ++m)
            if ((uintptr_t)ws[m] % 16) a.vecw = 0;
        a.QKV[l] = QKV[l]; a.Z1[l] = Z1[l]; a.st1[l] = st1[l]; a.C[l] = C[l]; a.F[l] = F[l]; a.Z2[l] = Z2[l]; a.st2[l] = st2[l];
        a.X[l + 1] = X[l + 1];
    }
    // token tile = T rounded up to a multiple of four (12 | 16 | 20 | 24); as many warps per CTA as fit beside the weights
    // (one CTA per SM): with the scripts' 20 history slots that is 16 warps, 2368 sessions in flight on the 148 SMs
    const int TP = T <= 12 ? 12 : (T <= 16 ? 16 : (T <= 20 ? 20 : 24));
    const size_t per_warp = (size_t)bf_warp_floats(TP) * 4, wbytes = (size_t)layers * BF_LAYER * 4;
    int warps = (int)((226 * 1024 - wbytes) / per_warp);
    warps = warps > 16 ? 16 : (warps < 1 ? 1 : warps);
    const size_t smem = wbytes + warps * per_warp;
    const unsigned grid = stream_grid(ceil_div(B, warps), 1);
#define BF_LAUNCH(TPV)                                                                    \
    do {                                                                                  \
        auto k = bert_fused_fwd_kernel<TPV>;                                              \
        ensure_smem(k, smem);                                                             \
        LAUNCH(k, dim3(grid), dim3(warps * 32), smem, s, a);                              \
    } while (0)
    if (TP == 12) BF_LAUNCH(12);
    else if (TP == 16) BF_LAUNCH(16);
    else if (TP == 20) BF_LAUNCH(20);
    else BF_LAUNCH(24);
#undef BF_LAUNCH
    const double tok = (double)B * T;
    return check_launch("bert_fused_fwd", tok * 4.0 * BF_D * (1 + (save ? 1 + layers * (3 + 5) : 0)),
                        tok * layers * (2.0 * 5 * BF_D * BF_D + 4.0 * T * BF_D));
}

}  // namespace intel
