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
