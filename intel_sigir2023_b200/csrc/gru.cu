// GRU4RecEncoder recurrence (GeneralSeq.py:58-78 -> torch.nn.GRU, gate order r,z,n), one masked time
// step per launch; the two projections per step (W_ih x, W_hh h) are GEMMs issued by the caller.
// Packed-sequence semantics: a session stops updating once t >= len, so h_all[:, T] is its last state.
#include "kernels.h"

namespace intel {

__global__ void __launch_bounds__(256) gru_step_fwd_kernel(int64_t B, int64_t T, int h, int t,
                                                           const int64_t* __restrict__ lens,
                                                           const float* __restrict__ gi, const float* __restrict__ gh,
                                                           float* h_all, float* __restrict__ gates) {
    const int64_t total = B * h;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / h;
        const int c = (int)(e % h);
        const float hp = h_all[(b * (T + 1) + t) * h + c];
        float hn = hp;
        if (t < lens[b]) {
            const float* gib = gi + (b * T + t) * 3 * h;
            const float* ghb = gh + b * 3 * h;
            const float r = sigmoidf_(gib[c] + ghb[c]);
            const float z = sigmoidf_(gib[h + c] + ghb[h + c]);
            const float ghn = ghb[2 * h + c];
            const float n = tanhf(gib[2 * h + c] + r * ghn);
            hn = (1.f - z) * n + z * hp;
            float* gt = gates + (b * T + t) * 4 * h;
            gt[c] = r; gt[h + c] = z; gt[2 * h + c] = n; gt[3 * h + c] = ghn;
        }
        h_all[(b * (T + 1) + t + 1) * h + c] = hn;
    }
}

int gru_step_fwd(int64_t B, int64_t T, int h, int t, const int64_t* lens, const float* gi, const float* gh,
                 float* h_all, float* gates, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(B * h, 256), 8);
    LAUNCH(gru_step_fwd_kernel, dim3(grid), dim3(256), 0, s, B, T, h, t, lens, gi, gh, h_all, gates);
    return check_launch("gru_step_fwd");
}

__global__ void __launch_bounds__(256) gru_step_bwd_kernel(int64_t B, int64_t T, int h, int t,
                                                           const int64_t* __restrict__ lens,
                                                           const float* __restrict__ h_all,
                                                           const float* __restrict__ gates, float* dh,
                                                           float* __restrict__ dgi, float* __restrict__ dgh_all) {
    const int64_t total = B * h;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / h;
        const int c = (int)(e % h);
        float* di = dgi + (b * T + t) * 3 * h;
        float* dg = dgh_all + (b * (T + 1) + t) * 3 * h;   // [B, T+1, 3h]: row layout of h_all (slot T stays 0)
        if (t < lens[b]) {
            const float* gt = gates + (b * T + t) * 4 * h;
            const float r = gt[c], z = gt[h + c], n = gt[2 * h + c], ghn = gt[3 * h + c];
            const float hp = h_all[(b * (T + 1) + t) * h + c];
            const float g = dh[e];
            const float dn = g * (1.f - z) * (1.f - n * n);
            const float dz = g * (hp - n) * z * (1.f - z);
            const float dr = dn * ghn * r * (1.f - r);
            di[c] = dr; di[h + c] = dz; di[2 * h + c] = dn;
            dg[c] = dr; dg[h + c] = dz; dg[2 * h + c] = dn * r;
            dh[e] = g * z;
        } else {
            di[c] = 0.f; di[h + c] = 0.f; di[2 * h + c] = 0.f;
            dg[c] = 0.f; dg[h + c] = 0.f; dg[2 * h + c] = 0.f;
        }
    }
}

int gru_step_bwd(int64_t B, int64_t T, int h, int t, const int64_t* lens, const float* h_all, const float* gates,
                 float* dh, float* dgi, float* dgh_all, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    unsigned grid = stream_grid(ceil_div(B * h, 256), 8);
    LAUNCH(gru_step_bwd_kernel, dim3(grid), dim3(256), 0, s, B, T, h, t, lens, h_all, gates, dh, dgi, dgh_all);
    return check_launch("gru_step_bwd");
}

}  // namespace intel

// =================================================================================================
// Fused recurrence for hidden size 128 (the reference hard-codes GRU4RecEncoder(hidden_size=128),
// IntEL.py:105-106).  The per-step formulation above needs 2 launches per time step and runs its
// [B,128]x[128,384] products on a few hundred CTAs; here ONE persistent CTA owns a tile of sessions for all
// T steps: W_hh (384x128 fp32 = 192 KB) stays in shared memory for the whole kernel, the hidden state of the
// tile lives in shared memory / registers, gh = W_hh h is formed by 3xTF32 tensor-core MMAs and the gate
// math runs in the MMA epilogue on the accumulator fragments (each warp owns 16 hidden units: its r, z and n
// columns of a unit land in the same lane).  HBM traffic is just gi in, h_all / gates out.
// =================================================================================================
#include "mma.cuh"

namespace intel {

static const int GH = 128;              // hidden size
static const int GW = GH + 4;           // smem row stride of W_hh / h tiles (conflict-free fragment reads)
static const int GF_SB = 32;            // sessions per CTA, forward
static const int GB_SB = 16;            // sessions per CTA, backward
static const int GD = 3 * GH + 4;       // smem row stride of the dgh tile

__device__ __forceinline__ void stage_whh(float* Ws, const float* __restrict__ w_hh) {
    for (int e = threadIdx.x; e < 3 * GH * (GH / 4); e += blockDim.x) {
        const int n = e / (GH / 4), k4 = (e % (GH / 4)) * 4;
        *reinterpret_cast<float4*>(Ws + n * GW + k4) = *reinterpret_cast<const float4*>(w_hh + n * GH + k4);
    }
}

__global__ void __launch_bounds__(256, 1) gru_seq_fwd_kernel(int64_t B, int64_t T, const int64_t* __restrict__ lens,
                                                             const float* __restrict__ gi,
                                                             const float* __restrict__ w_hh,
                                                             const float* __restrict__ b_hh, float* __restrict__ h_all,
                                                             float* __restrict__ gates) {
    DYN_SMEM(float, sm);
    float* Ws = sm;                          // [384][GW]
    float* hs = Ws + 3 * GH * GW;            // [GF_SB][GW]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int64_t b0 = (int64_t)blockIdx.x * GF_SB;
    stage_whh(Ws, w_hh);
    for (int e = threadIdx.x; e < GF_SB * GW; e += blockDim.x) hs[e] = 0.f;
    // rows of this lane: r = i*16 + hh*8 + gq (i: m-tile, hh: half); columns: c = 16*warp + 8*ct + 2*tq + e
    int64_t len_r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t b = b0 + (q >> 1) * 16 + (q & 1) * 8 + gq;
        len_r[q] = (b < B) ? lens[b] : 0;
    }
    float bias[3][2][2];
#pragma unroll
    for (int gte = 0; gte < 3; ++gte)
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) bias[gte][ct][e] = b_hh[gte * GH + 16 * warp + 8 * ct + 2 * tq + e];
    __syncthreads();

    for (int64_t t = 0; t < T; ++t) {
        // this step's input pre-activations: requested before the recurrent product so that the HBM latency is hidden
        float2 gin[2][2][2][3];             // [m-tile][half][ct][gate]
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int64_t b = b0 + i * 16 + hh * 8 + gq;
                const bool live = (b < B) && (t < len_r[i * 2 + hh]);
#pragma unroll
                for (int ct = 0; ct < 2; ++ct)
#pragma unroll
                    for (int gte = 0; gte < 3; ++gte)
                        gin[i][hh][ct][gte] = live ? *reinterpret_cast<const float2*>(gi + (b * T + t) * 3 * GH + gte * GH + 16 * warp + 8 * ct + 2 * tq)
                                                   : make_float2(0.f, 0.f);
            }
        float acc[2][6][4];                 // [m-tile][gate*2 + ct][frag]
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
#pragma unroll 2
        for (int ks = 0; ks < GH / 8; ++ks) {
            const int k0 = ks * 8 + tq;
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                split_tf32(hs[(i * 16 + gq) * GW + k0], ah[i][0], al[i][0]);
                split_tf32(hs[(i * 16 + gq + 8) * GW + k0], ah[i][1], al[i][1]);
                split_tf32(hs[(i * 16 + gq) * GW + k0 + 4], ah[i][2], al[i][2]);
                split_tf32(hs[(i * 16 + gq + 8) * GW + k0 + 4], ah[i][3], al[i][3]);
            }
            uint32_t bh[6][2], bl[6][2];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int n = (j >> 1) * GH + 16 * warp + 8 * (j & 1) + gq;
                split_tf32(Ws[n * GW + k0], bh[j][0], bl[j][0]);
                split_tf32(Ws[n * GW + k0 + 4], bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) mma_tf32(acc[i][j], al[i], bh[j]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) mma_tf32(acc[i][j], ah[i], bl[j]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) mma_tf32(acc[i][j], ah[i], bh[j]);
        }
        // ---- gates on the accumulator fragments ----
        float hn[2][2][2][2];               // [m-tile][half][ct][e]
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int row = i * 16 + hh * 8 + gq;
                const int64_t b = b0 + row;
                const bool live = (b < B) && (t < len_r[i * 2 + hh]);
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) {
                    const int c = 16 * warp + 8 * ct + 2 * tq;
                    const float hp0 = hs[row * GW + c], hp1 = hs[row * GW + c + 1];
                    float o0 = hp0, o1 = hp1;
                    if (live) {
                        const float2 ir = gin[i][hh][ct][0], iz = gin[i][hh][ct][1], in = gin[i][hh][ct][2];
                        const float r0 = sigmoidf_(ir.x + acc[i][0 + ct][2 * hh] + bias[0][ct][0]);
                        const float r1 = sigmoidf_(ir.y + acc[i][0 + ct][2 * hh + 1] + bias[0][ct][1]);
                        const float z0 = sigmoidf_(iz.x + acc[i][2 + ct][2 * hh] + bias[1][ct][0]);
                        const float z1 = sigmoidf_(iz.y + acc[i][2 + ct][2 * hh + 1] + bias[1][ct][1]);
                        const float g0 = acc[i][4 + ct][2 * hh] + bias[2][ct][0];
                        const float g1 = acc[i][4 + ct][2 * hh + 1] + bias[2][ct][1];
                        const float n0 = tanhf(in.x + r0 * g0), n1 = tanhf(in.y + r1 * g1);
                        o0 = (1.f - z0) * n0 + z0 * hp0;
                        o1 = (1.f - z1) * n1 + z1 * hp1;
                        float* gt = gates + (b * T + t) * 4 * GH;
                        *reinterpret_cast<float2*>(gt + c) = make_float2(r0, r1);
                        *reinterpret_cast<float2*>(gt + GH + c) = make_float2(z0, z1);
                        *reinterpret_cast<float2*>(gt + 2 * GH + c) = make_float2(n0, n1);
                        *reinterpret_cast<float2*>(gt + 3 * GH + c) = make_float2(g0, g1);
                    }
                    hn[i][hh][ct][0] = o0;
                    hn[i][hh][ct][1] = o1;
                    if (b < B) *reinterpret_cast<float2*>(h_all + (b * (T + 1) + t + 1) * GH + c) = make_float2(o0, o1);
                }
            }
        __syncthreads();                    // every warp has finished reading the old state
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int ct = 0; ct < 2; ++ct) {
                    const int row = i * 16 + hh * 8 + gq, c = 16 * warp + 8 * ct + 2 * tq;
                    hs[row * GW + c] = hn[i][hh][ct][0];
                    hs[row * GW + c + 1] = hn[i][hh][ct][1];
                }
        __syncthreads();
    }
}

// h_all[:, 0] must be zero (the caller clears it); h_all [B, T+1, 128], gates [B, T, 512]
int gru_seq_fwd(int64_t B, int64_t T, int h, const int64_t* lens, const float* gi, const float* w_hh,
                const float* b_hh, float* h_all, float* gates, cudaStream_t s, bool save_gates) {
    if (B <= 0 || T <= 0) return INTEL_OK;
    INTEL_REQUIRE(h == GH, INTEL_ERR_UNSUPPORTED, "fused GRU needs hidden size 128");
    if (gru_tc_supported(h)) return gru_tc_fwd(B, T, lens, gi, w_hh, b_hh, h_all, gates, s, save_gates);
    const size_t smem = (size_t)(3 * GH * GW + GF_SB * GW) * 4;
    ensure_smem(gru_seq_fwd_kernel, smem);
    LAUNCH(gru_seq_fwd_kernel, dim3((unsigned)ceil_div(B, GF_SB)), dim3(256), smem, s, B, T, lens, gi, w_hh, b_hh, h_all,
           gates);
    return check_launch("gru_seq_fwd", (double)B * T * (3 + 4 + 1) * GH * 4.0, 2.0 * B * T * 3 * GH * GH);
}

// ------------------------------------------------------------------------------------------------
// Output projections of the two GRU encoders (GeneralSeq.py:76-77: out = Linear(128 -> d, no bias) of the last state) and
// their input gradients, both encoders in one launch: a [4096 x 48..64 x 128] product is far too small for a GEMM launch of
// its own (four launches of 22 us each in round 2).  One warp per session; W [n][128] sits in shared memory with row stride 129.
namespace {
constexpr int OP_H = 128, OP_WS = OP_H + 1, OP_WARPS = 8, OP_NMAX = 64;
struct OutProjArgs {
    int64_t B;
    int n[2];
    const float* W[2];
    const float* x[2]; int64_t ldx[2];      // forward: the last states; backward: d(out) rows
    float* y[2]; int64_t ldy[2];            // forward: out rows; backward: d(last state) [B, 128]
};
}  // namespace

__global__ void __launch_bounds__(OP_WARPS * 32) gru_outproj_fwd_kernel(OutProjArgs a) {
    DYN_SMEM(float, sm);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* Ws[2] = {sm, sm + OP_NMAX * OP_WS};
    float* xb = sm + 2 * OP_NMAX * OP_WS + w * OP_H;
    // weights -> shared memory with independent 16-byte loads (nn.Linear weights are 16-byte aligned allocations)
    for (int i = 0; i < 2; ++i) {
#pragma unroll 8
        for (int e = threadIdx.x; e < a.n[i] * (OP_H / 4); e += blockDim.x) {
            const float4 v = *reinterpret_cast<const float4*>(a.W[i] + 4 * e);
            float* dst = Ws[i] + (e >> 5) * OP_WS + 4 * (e & 31);
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
    }
    __syncthreads();
    const int64_t nwarps = (int64_t)gridDim.x * OP_WARPS;
    for (int64_t b = (int64_t)blockIdx.x * OP_WARPS + w; b < a.B; b += nwarps) {
        for (int i = 0; i < 2; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(a.x[i] + b * a.ldx[i] + 4 * lane);
            __syncwarp();
            *reinterpret_cast<float4*>(xb + 4 * lane) = v;
            __syncwarp();
            const int j0 = lane, j1 = lane + 32;
            const bool on0 = j0 < a.n[i], on1 = j1 < a.n[i];
            const float* w0 = Ws[i] + (on0 ? j0 : 0) * OP_WS;
            const float* w1 = Ws[i] + (on1 ? j1 : 0) * OP_WS;
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
            for (int u = 0; u < OP_H; u += 4) {
                const float4 h4 = *reinterpret_cast<const float4*>(xb + u);
                acc0 = fmaf(h4.x, w0[u], fmaf(h4.y, w0[u + 1], fmaf(h4.z, w0[u + 2], fmaf(h4.w, w0[u + 3], acc0))));
                acc1 = fmaf(h4.x, w1[u], fmaf(h4.y, w1[u + 1], fmaf(h4.z, w1[u + 2], fmaf(h4.w, w1[u + 3], acc1))));
            }
            if (on0) a.y[i][b * a.ldy[i] + j0] = acc0;
            if (on1) a.y[i][b * a.ldy[i] + j1] = acc1;
        }
    }
}

// d(last state)[b, u] = sum_j d(out)[b, j] W[j][u]
__global__ void __launch_bounds__(OP_WARPS * 32) gru_outproj_dx_kernel(OutProjArgs a) {
    DYN_SMEM(float, sm);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* Ws[2] = {sm, sm + OP_NMAX * OP_WS};
    float* gb = sm + 2 * OP_NMAX * OP_WS + w * OP_NMAX;
    // weights -> shared memory with independent 16-byte loads (nn.Linear weights are 16-byte aligned allocations)
    for (int i = 0; i < 2; ++i) {
#pragma unroll 8
        for (int e = threadIdx.x; e < a.n[i] * (OP_H / 4); e += blockDim.x) {
            const float4 v = *reinterpret_cast<const float4*>(a.W[i] + 4 * e);
            float* dst = Ws[i] + (e >> 5) * OP_WS + 4 * (e & 31);
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
    }
    __syncthreads();
    const int64_t nwarps = (int64_t)gridDim.x * OP_WARPS;
    for (int64_t b = (int64_t)blockIdx.x * OP_WARPS + w; b < a.B; b += nwarps) {
        for (int i = 0; i < 2; ++i) {
            __syncwarp();
            gb[lane] = lane < a.n[i] ? a.x[i][b * a.ldx[i] + lane] : 0.f;
            gb[lane + 32] = lane + 32 < a.n[i] ? a.x[i][b * a.ldx[i] + lane + 32] : 0.f;
            __syncwarp();
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < a.n[i]; ++j) {
                const float g = gb[j];
                const float* wr = Ws[i] + j * OP_WS + lane;
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = fmaf(g, wr[32 * q], acc[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) a.y[i][b * a.ldy[i] + lane + 32 * q] = acc[q];
        }
    }
}

bool gru_outproj_pair_ok(int h, int n0, int n1) { return h == OP_H && n0 >= 1 && n0 <= OP_NMAX && n1 >= 1 && n1 <= OP_NMAX; }

static int gru_outproj_launch(bool dx, int64_t B, const int* n, const float* const* W, const float* const* x, const int64_t* ldx,
                              float* const* y, const int64_t* ldy, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    OutProjArgs a;
    a.B = B;
    for (int i = 0; i < 2; ++i) {
        a.n[i] = n[i]; a.W[i] = W[i]; a.x[i] = x[i]; a.ldx[i] = ldx[i]; a.y[i] = y[i]; a.ldy[i] = ldy[i];
        INTEL_REQUIRE(dx || (ldx[i] % 4 == 0 && (uintptr_t)x[i] % 16 == 0), INTEL_ERR_ARG, "gru_outproj: state rows must be 16-byte aligned");
        INTEL_REQUIRE((uintptr_t)W[i] % 16 == 0, INTEL_ERR_ARG, "gru_outproj: weights must be 16-byte aligned");
    }
    const size_t smem = (size_t)(2 * OP_NMAX * OP_WS + OP_WARPS * OP_H) * 4;
    const unsigned grid = stream_grid(ceil_div(B, OP_WARPS), 2);
    if (dx) {
        ensure_smem(gru_outproj_dx_kernel, smem);
        LAUNCH(gru_outproj_dx_kernel, dim3(grid), dim3(OP_WARPS * 32), smem, s, a);
    } else {
        ensure_smem(gru_outproj_fwd_kernel, smem);
        LAUNCH(gru_outproj_fwd_kernel, dim3(grid), dim3(OP_WARPS * 32), smem, s, a);
    }
    return check_launch(dx ? "gru_outproj_dx" : "gru_outproj_fwd", (double)B * 4.0 * (2 * OP_H + n[0] + n[1]), 2.0 * B * OP_H * (n[0] + n[1]));
}
int gru_outproj_pair_fwd(int64_t B, const int* n, const float* const* W, const float* const* h_last, const int64_t* ldh,
                         float* const* out, const int64_t* ldo, cudaStream_t s) {
    return gru_outproj_launch(false, B, n, W, h_last, ldh, out, ldo, s);
}
int gru_outproj_pair_dx(int64_t B, const int* n, const float* const* W, const float* const* dout, const int64_t* ldd,
                        float* const* dh, const int64_t* ldh, cudaStream_t s) {
    return gru_outproj_launch(true, B, n, W, dout, ldd, dh, ldh, s);
}

// Row numbers of the live (session, step) pairs of a padded [B, T] history, in (b, t) order.  Block k owns sessions
// [1024 k, 1024 k + 1024): it sums the lengths before them, scans its own and writes its rows; the last block writes the count.
__device__ void gru_live_rows_body(int blk, int nblk, int64_t B, int T, const int64_t* __restrict__ lens, int32_t* __restrict__ rows_t,
                                   int32_t* __restrict__ rows_t1, int32_t* __restrict__ count) {
    __shared__ int warp_tot[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto len_of = [&](int64_t e) { const int64_t v = lens[e]; return (int)(v < 0 ? 0 : (v > T ? T : v)); };
    const int64_t first = (int64_t)blk * 1024;
    // lengths before this block
    int part = 0;
    for (int64_t e = tid; e < first; e += 1024) part += len_of(e);
    part = warp_sum_i(part);
    if (lane == 0) warp_tot[warp] = part;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < 32; ++w) t += warp_tot[w];
        base_s = t;
    }
    __syncthreads();
    const int base = base_s;
    __syncthreads();
    // exclusive scan of the block's own lengths
    const int64_t b = first + tid;
    const int n = b < B ? len_of(b) : 0;
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    const int off = base + before + incl - n;
    // a warp writes the rows of its 32 sessions one session at a time: lane t takes step t (contiguous stores)
    for (int i = 0; i < 32; ++i) {
        const int o = __shfl_sync(0xffffffffu, off, i), m = __shfl_sync(0xffffffffu, n, i);
        const int64_t bb = first + warp * 32 + i;
        for (int t = lane; t < m; t += 32) {
            rows_t[o + t] = (int32_t)(bb * T + t);
            rows_t1[o + t] = (int32_t)(bb * (T + 1) + t);
        }
    }
    if (blk == nblk - 1 && tid == 1023) *count = off + n;
}
// Sessions in order of decreasing length (stable), the way pack_padded_sequence orders them (GeneralSeq.py:64-71): a tile
// of consecutive sessions of that order then shares one loop bound, and tiles that end early make room for the next ones.
// One block; a chunk of 1024 sessions per pass, rank inside a chunk from warp ballots, so the order is deterministic.
__device__ void gru_order_body(int64_t B, int T, const int64_t* __restrict__ lens, int32_t* __restrict__ order) {
    __shared__ int cnt[64], start[64], run[64], wcnt[32][64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) { cnt[tid] = 0; run[tid] = 0; }
    if (tid == 0) order[B] = 0;              // tile counter of the backward recurrence, kept behind the B entries
    __syncthreads();
    auto key = [&](int64_t e) { const int64_t v = lens[e]; return (int)(v < 0 ? 0 : (v > T ? T : v)); };
    for (int64_t c0 = 0; c0 < B; c0 += 1024) {           // histogram: one shared-memory atomic per warp and length
        const int64_t e = c0 + tid;
        const int v = e < B ? key(e) : -1;
        for (int u = 0; u <= T; ++u) {
            const unsigned m = __ballot_sync(0xffffffffu, v == u);
            if (lane == 0 && m) atomicAdd(&cnt[u], __popc(m));
        }
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int v = T; v >= 0; --v) { start[v] = acc; acc += cnt[v]; }
    }
    __syncthreads();
    for (int64_t c0 = 0; c0 < B; c0 += 1024) {
        const int64_t e = c0 + tid;
        const int v = e < B ? key(e) : -1;
        int rank = 0;
        for (int u = 0; u <= T; ++u) {
            const unsigned m = __ballot_sync(0xffffffffu, v == u);
            if (lane == 0) wcnt[warp][u] = __popc(m);
            if (v == u) rank = __popc(m & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (v >= 0) {
            int pre = 0;
            for (int w = 0; w < warp; ++w) pre += wcnt[w][v];
            order[start[v] + run[v] + pre + rank] = (int32_t)e;
        }
        __syncthreads();
        if (tid <= T) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += wcnt[w][tid];
            run[tid] += tot;
        }
        __syncthreads();
    }
}
// One launch prepares both encoders: blocks [0, nb) of row y list the live rows of encoder y, block nb sorts its sessions.
struct GruPrepArgs {
    struct One { int64_t B; int T; const int64_t* lens; int32_t *rows_t, *rows_t1, *count, *order; } e[2];
    int n;
};
__global__ void __launch_bounds__(1024) gru_prep_kernel(GruPrepArgs a) {
    const GruPrepArgs::One& E = a.e[blockIdx.y];
    const int nb = (int)((E.B + 1023) / 1024);
    if ((int)blockIdx.x < nb) {
        if (E.rows_t) gru_live_rows_body((int)blockIdx.x, nb, E.B, E.T, E.lens, E.rows_t, E.rows_t1, E.count);
    } else if ((int)blockIdx.x == nb) {
        if (E.order) gru_order_body(E.B, E.T, E.lens, E.order);
    }
}
int gru_prep(int n, const int64_t* B, const int64_t* T, const int64_t* const* lens, int32_t* const* rows_t, int32_t* const* rows_t1,
             int32_t* const* count, int32_t* const* order, cudaStream_t s) {
    GruPrepArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    int64_t nb = 0;
    for (int i = 0; i < n; ++i) {
        INTEL_REQUIRE(B[i] * (T[i] + 1) < (1LL << 31) || !rows_t[i], INTEL_ERR_UNSUPPORTED, "gru_prep: more than 2^31 history rows");
        INTEL_REQUIRE(!order[i] || (T[i] >= 1 && T[i] <= 63), INTEL_ERR_UNSUPPORTED, "gru_prep: sorting needs T in [1, 63]");
        a.e[i].B = B[i]; a.e[i].T = (int)T[i]; a.e[i].lens = lens[i];
        a.e[i].rows_t = rows_t[i]; a.e[i].rows_t1 = rows_t1[i]; a.e[i].count = count[i]; a.e[i].order = order[i];
        const int64_t x = ceil_div(B[i], 1024);
        nb = x > nb ? x : nb;
    }
    if (n <= 0 || nb <= 0) return INTEL_OK;
    LAUNCH(gru_prep_kernel, dim3((unsigned)(nb + 1), (unsigned)n), dim3(1024), 0, s, a);
    return check_launch("gru_prep");
}
int gru_live_rows(int64_t B, int64_t T, const int64_t* lens, int32_t* rows_t, int32_t* rows_t1, int32_t* count, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    int32_t* none = nullptr;
    return gru_prep(1, &B, &T, &lens, &rows_t, &rows_t1, &count, &none, s);
}
int gru_order_by_len(int64_t B, int64_t T, const int64_t* lens, int32_t* order, cudaStream_t s) {
    if (B <= 0) return INTEL_OK;
    int32_t* none = nullptr;
    return gru_prep(1, &B, &T, &lens, &none, &none, &none, &order, s);
}

// Backward recurrence.  dh [B,128] holds d(loss)/d h_T on entry.  Writes dgi [B,T,384] and
// dgh_all [B,T+1,384] (slot T untouched: the caller clears the buffer) for the weight-gradient GEMMs.
__global__ void __launch_bounds__(256, 1) gru_seq_bwd_kernel(int64_t B, int64_t T, const int64_t* __restrict__ lens,
                                                             const float* __restrict__ w_hh,
                                                             const float* __restrict__ h_all,
                                                             const float* __restrict__ gates,
                                                             const float* __restrict__ dh_in, float* __restrict__ dgi,
                                                             float* __restrict__ dgh_all, float* __restrict__ db_ih,
                                                             float* __restrict__ db_hh, const int32_t* __restrict__ order,
                                                             int32_t* __restrict__ counter) {
    DYN_SMEM(float, sm);
    __shared__ int s_tile;
    float* Ws = sm;                          // [384][GW]   W_hh[k = gate column][n = hidden]
    float* ds = Ws + 3 * GH * GW;            // [GB_SB][GD] dgh of the current step
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    stage_whh(Ws, w_hh);                     // once per CTA: the CTA is persistent and walks tiles of GB_SB sessions
    float sum_i[2][3][2], sum_h[2][2];        // bias-gradient partial sums: [ct][gate][e] of dgi, [ct][e] of the n-gate of dgh
#pragma unroll
    for (int ct = 0; ct < 2; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            sum_h[ct][e] = 0.f;
#pragma unroll
            for (int q = 0; q < 3; ++q) sum_i[ct][q][e] = 0.f;
        }
    const int64_t ntiles = (B + GB_SB - 1) / GB_SB;
    for (int64_t it = 0;; ++it) {
    // next tile: with the length-sorted order the tiles are handed out longest first from a device counter (the CTAs that
    // drew short tiles come back for more), otherwise round robin
    int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
    if (counter) {
        if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1);
        __syncthreads();
        tile = s_tile;
    }
    if (tile >= ntiles) break;
    const int64_t b0 = tile * GB_SB;
    // the tile's sessions: consecutive entries of the length-sorted order (longest first), so the first one bounds the loop;
    // session slots beyond B are marked with B (every access below is guarded by b < B)
    auto sess = [&](int64_t i) -> int64_t { return i < B ? (order ? (int64_t)order[i] : i) : B; };
    int64_t tmax = T;
    if (order) {
        const int64_t l0 = lens[sess(b0)];
        tmax = l0 < 0 ? 0 : (l0 > T ? T : l0);
    }
    // gradient rows behind the tile's last live step are zero (the weight-gradient products read every row)
    if (tmax < T) {
        const int64_t per = (T - tmax) * (3 * GH / 4);
        for (int64_t e = threadIdx.x; e < (int64_t)GB_SB * per; e += blockDim.x) {
            const int64_t b = sess(b0 + e / per), r = e % per;
            if (b < B) {
                const int64_t t = tmax + r / (3 * GH / 4), c4 = (r % (3 * GH / 4)) * 4;
                *reinterpret_cast<float4*>(dgi + (b * T + t) * 3 * GH + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(dgh_all + (b * (T + 1) + t) * 3 * GH + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    int64_t len_r[2], b_r[2];
    float dh[2][2][2];                       // [half][ct][e]: rows gq / gq+8, cols 16*warp + 8*ct + 2*tq + e
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int64_t b = sess(b0 + hh * 8 + gq);
        b_r[hh] = b;
        len_r[hh] = (b < B) ? lens[b] : 0;
#pragma unroll
        for (int ct = 0; ct < 2; ++ct) {
            const int c = 16 * warp + 8 * ct + 2 * tq;
            float2 v = make_float2(0.f, 0.f);
            if (b < B) v = *reinterpret_cast<const float2*>(dh_in + b * GH + c);
            dh[hh][ct][0] = v.x;
            dh[hh][ct][1] = v.y;
        }
    }
    __syncthreads();
    // saved gate values r | z | n | W_hn h + b_hn and the previous state of one step, for the elements this lane owns;
    // the values of step t - 1 are requested before the recurrent product of step t (latency hidden behind the MMAs)
    float2 sv[2][2][5];
    auto fetch = [&](int64_t t, float2 (&v)[2][2][5]) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int64_t b = b_r[hh];
            const bool live = (b < B) && (t >= 0) && (t < len_r[hh]);
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
                const int c = 16 * warp + 8 * ct + 2 * tq;
#pragma unroll
                for (int q = 0; q < 5; ++q) v[hh][ct][q] = make_float2(0.f, 0.f);
                if (live) {
                    const float* gt = gates + (b * T + t) * 4 * GH;
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[hh][ct][q] = *reinterpret_cast<const float2*>(gt + q * GH + c);
                    v[hh][ct][4] = *reinterpret_cast<const float2*>(h_all + (b * (T + 1) + t) * GH + c);
                }
            }
        }
    };
    fetch(tmax - 1, sv);
    for (int64_t t = tmax - 1; t >= 0; --t) {
        // ---- gate derivatives for the elements this lane owns ----
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int row = hh * 8 + gq;
            const int64_t b = b_r[hh];
            const bool live = (b < B) && (t < len_r[hh]);
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
                const int c = 16 * warp + 8 * ct + 2 * tq;
                float2 dr = make_float2(0.f, 0.f), dz = dr, dn = dr, dnr = dr;
                if (live) {
                    const float2 r = sv[hh][ct][0], z = sv[hh][ct][1], n = sv[hh][ct][2], gn = sv[hh][ct][3], hp = sv[hh][ct][4];
                    const float g0 = dh[hh][ct][0], g1 = dh[hh][ct][1];
                    dn.x = g0 * (1.f - z.x) * (1.f - n.x * n.x);
                    dn.y = g1 * (1.f - z.y) * (1.f - n.y * n.y);
                    dz.x = g0 * (hp.x - n.x) * z.x * (1.f - z.x);
                    dz.y = g1 * (hp.y - n.y) * z.y * (1.f - z.y);
                    dr.x = dn.x * gn.x * r.x * (1.f - r.x);
                    dr.y = dn.y * gn.y * r.y * (1.f - r.y);
                    dnr.x = dn.x * r.x;
                    dnr.y = dn.y * r.y;
                    dh[hh][ct][0] = g0 * z.x;
                    dh[hh][ct][1] = g1 * z.y;
                }
                if (b < B) {
                    float* di = dgi + (b * T + t) * 3 * GH;
                    float* dg = dgh_all + (b * (T + 1) + t) * 3 * GH;
                    *reinterpret_cast<float2*>(di + c) = dr;
                    *reinterpret_cast<float2*>(di + GH + c) = dz;
                    *reinterpret_cast<float2*>(di + 2 * GH + c) = dn;
                    *reinterpret_cast<float2*>(dg + c) = dr;
                    *reinterpret_cast<float2*>(dg + GH + c) = dz;
                    *reinterpret_cast<float2*>(dg + 2 * GH + c) = dnr;
                }
                sum_i[ct][0][0] += dr.x; sum_i[ct][0][1] += dr.y;
                sum_i[ct][1][0] += dz.x; sum_i[ct][1][1] += dz.y;
                sum_i[ct][2][0] += dn.x; sum_i[ct][2][1] += dn.y;
                sum_h[ct][0] += dnr.x;   sum_h[ct][1] += dnr.y;
                ds[row * GD + c] = dr.x;            ds[row * GD + c + 1] = dr.y;
                ds[row * GD + GH + c] = dz.x;       ds[row * GD + GH + c + 1] = dz.y;
                ds[row * GD + 2 * GH + c] = dnr.x;  ds[row * GD + 2 * GH + c + 1] = dnr.y;
            }
        }
        __syncthreads();
        float2 nx[2][2][5];
        fetch(t - 1, nx);
        // ---- dh_prev = dh * z + dgh W_hh : [16 x 384] x [384 x 128], this warp owns 16 output columns ----
        float acc[2][4];
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[ct][c] = 0.f;
#pragma unroll 4
        for (int ks = 0; ks < 3 * GH / 8; ++ks) {
            const int k0 = ks * 8 + tq;
            uint32_t ah[4], al[4];
            split_tf32(ds[gq * GD + k0], ah[0], al[0]);
            split_tf32(ds[(gq + 8) * GD + k0], ah[1], al[1]);
            split_tf32(ds[gq * GD + k0 + 4], ah[2], al[2]);
            split_tf32(ds[(gq + 8) * GD + k0 + 4], ah[3], al[3]);
            uint32_t bh[2][2], bl[2][2];
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
                const int n = 16 * warp + 8 * ct + gq;
                split_tf32(Ws[k0 * GW + n], bh[ct][0], bl[ct][0]);
                split_tf32(Ws[(k0 + 4) * GW + n], bh[ct][1], bl[ct][1]);
            }
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) mma_tf32(acc[ct], al, bh[ct]);
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) mma_tf32(acc[ct], ah, bl[ct]);
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) mma_tf32(acc[ct], ah, bh[ct]);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
            for (int ct = 0; ct < 2; ++ct) {
                dh[hh][ct][0] += acc[ct][2 * hh];
                dh[hh][ct][1] += acc[ct][2 * hh + 1];
#pragma unroll
                for (int q = 0; q < 5; ++q) sv[hh][ct][q] = nx[hh][ct][q];
            }
        __syncthreads();                    // the dgh tile is rewritten by the next step
    }
    }   // tile loop
    // ---- bias gradients: db_ih = column sums of dgi, db_hh = column sums of dgh (r and z parts are shared) ----
    if (db_ih != nullptr || db_hh != nullptr) {
#pragma unroll
        for (int ct = 0; ct < 2; ++ct)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float v[4] = {sum_i[ct][0][e], sum_i[ct][1][e], sum_i[ct][2][e], sum_h[ct][e]};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 4);
                    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 8);
                    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
                }
                if (gq == 0) {
                    const int c = 16 * warp + 8 * ct + 2 * tq + e;
                    if (db_ih) { atomicAdd(db_ih + c, v[0]); atomicAdd(db_ih + GH + c, v[1]); atomicAdd(db_ih + 2 * GH + c, v[2]); }
                    if (db_hh) { atomicAdd(db_hh + c, v[0]); atomicAdd(db_hh + GH + c, v[1]); atomicAdd(db_hh + 2 * GH + c, v[3]); }
                }
            }
    }
}

int gru_seq_bwd(int64_t B, int64_t T, int h, const int64_t* lens, const float* w_hh, const float* h_all,
                const float* gates, const float* dh_in, float* dgi, float* dgh_all, float* db_ih, float* db_hh, cudaStream_t s,
                int32_t* order) {
    if (B <= 0 || T <= 0) return INTEL_OK;
    INTEL_REQUIRE(h == GH, INTEL_ERR_UNSUPPORTED, "fused GRU needs hidden size 128");
    const size_t smem = (size_t)(3 * GH * GW + GB_SB * GD) * 4;
    ensure_smem(gru_seq_bwd_kernel, smem);
    // persistent: one CTA per SM (W_hh fills its shared memory), tiles drawn from the counter behind `order` (order[B])
    const unsigned grid = stream_grid(ceil_div(B, GB_SB), 1);
    LAUNCH(gru_seq_bwd_kernel, dim3(grid), dim3(256), smem, s, B, T, lens, w_hh, h_all, gates,
           dh_in, dgi, dgh_all, db_ih, db_hh, (const int32_t*)order, order ? order + B : (int32_t*)nullptr);
    return check_launch("gru_seq_bwd", (double)B * T * (4 + 1 + 3 + 3) * GH * 4.0, 2.0 * B * T * 3 * GH * GH);
}

}  // namespace intel
