// fp32 FFMA tiled GEMM for the dense contractions of the IntEL path (nn.Linear forward,
// input gradients and weight gradients).  fp32 parity (1e-5) rules out plain TF32/BF16
// tensor-core math, and every contraction here has a tiny inner
// dimension (32..384) or a tiny output, so the kernel is a register-tiled FFMA GEMM:
// 256 threads, BK = 16, each thread owns a (BM/16) x (BN/16) accumulator tile.
#include "kernels.h"

namespace intel {

struct GemmDev {
    int64_t M, N, K;
    const float* A; int64_t lda;
    const float* B; int64_t ldb;
    float* C; int64_t ldc;
    const float* bias;
    const float* add; int64_t ldadd;
    const float* mask; int64_t ldmask;
    int relu_a, relu_b, relu_out, accumulate, splits;
    int64_t kchunk;
};

static const int BK = 16;

template <int BM, int BN, bool AT, bool BT>
__global__ void __launch_bounds__(256) gemm_kernel(GemmDev g) {
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int SA = BM + 4, SB = BN + 4;
    __shared__ float As[BK * SA];
    __shared__ float Bs[BK * SB];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int64_t n0 = (int64_t)blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * g.kchunk;
    const int64_t kend = (kbeg + g.kchunk < g.K) ? kbeg + g.kchunk : g.K;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        // ---- stage the A tile: As[kk][mm] ----
#pragma unroll
        for (int p = 0; p < (BM * BK) / 256; ++p) {
            int e = tid + p * 256;
            int kk, mm;
            if (AT) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
            int64_t gm = m0 + mm, gk = k0 + kk;
            float v = 0.f;
            if (gm < g.M && gk < kend) v = AT ? g.A[gk * g.lda + gm] : g.A[gm * g.lda + gk];
            if (g.relu_a) v = fmaxf(v, 0.f);
            As[kk * SA + mm] = v;
        }
        // ---- stage the B tile: Bs[kk][nn] ----
#pragma unroll
        for (int p = 0; p < (BN * BK) / 256; ++p) {
            int e = tid + p * 256;
            int kk, nn;
            if (BT) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
            int64_t gn = n0 + nn, gk = k0 + kk;
            float v = 0.f;
            if (gn < g.N && gk < kend) v = BT ? g.B[gk * g.ldb + gn] : g.B[gn * g.ldb + gk];
            if (g.relu_b) v = fmaxf(v, 0.f);
            Bs[kk * SB + nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk * SA + ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[kk * SB + tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const bool first = (blockIdx.z == 0);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int64_t gm = m0 + ty * TM + i;
        if (gm >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int64_t gn = n0 + tx * TN + j;
            if (gn >= g.N) continue;
            float v = acc[i][j];
            if (first) {
                if (g.bias) v += g.bias[gn];
                if (g.add) v += g.add[gm * g.ldadd + gn];
            }
            if (g.relu_out) v = fmaxf(v, 0.f);
            if (g.mask) v = (g.mask[gm * g.ldmask + gn] > 0.f) ? v : 0.f;
            float* c = g.C + gm * g.ldc + gn;
            if (g.accumulate == 0) *c = v;
            else if (g.accumulate == 1) *c += v;
            else atomicAdd(c, v);
        }
    }
}

template <int BM, int BN>
static int launch_tile(const GemmDev& d, bool at, bool bt, cudaStream_t s) {
    dim3 grid((unsigned)ceil_div(d.M, BM), (unsigned)ceil_div(d.N, BN), (unsigned)d.splits);
    dim3 block(256);
    if (!at && !bt) { auto k = gemm_kernel<BM, BN, false, false>; LAUNCH(k, grid, block, 0, s, d); }
    else if (!at && bt) { auto k = gemm_kernel<BM, BN, false, true>; LAUNCH(k, grid, block, 0, s, d); }
    else if (at && bt) { auto k = gemm_kernel<BM, BN, true, true>; LAUNCH(k, grid, block, 0, s, d); }
    else { auto k = gemm_kernel<BM, BN, true, false>; LAUNCH(k, grid, block, 0, s, d); }
    const double bytes = 4.0 * ((double)d.M * d.K + (double)d.N * d.K + (double)d.M * d.N * (1 + (d.add != nullptr) + (d.mask != nullptr)));
    return check_launch(at ? "gemm_wgrad" : (bt ? "gemm_dgrad" : "gemm_fwd"), bytes, 2.0 * d.M * d.N * d.K);
}

int gemm(const Gemm& g, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0) return INTEL_OK;
    INTEL_REQUIRE(g.A && g.B && g.C, INTEL_ERR_ARG, "gemm: null operand");
    INTEL_REQUIRE(g.M <= 0x7fffffffLL * 32 && g.N < (65535LL * 32), INTEL_ERR_ARG, "gemm: shape too large");
    GemmDev d;
    d.M = g.M; d.N = g.N; d.K = g.K;
    d.A = g.A; d.lda = g.lda; d.B = g.B; d.ldb = g.ldb; d.C = g.C; d.ldc = g.ldc;
    d.bias = g.bias; d.add = g.add; d.ldadd = g.ldadd; d.mask = g.mask; d.ldmask = g.ldmask;
    d.relu_a = g.relu_a; d.relu_b = g.relu_b; d.relu_out = g.relu_out; d.accumulate = g.accumulate;
    const bool small_m = g.M <= 32, small_n = g.N <= 32;
    const int bm = small_m ? 32 : (small_n ? 128 : 64);
    const int bn = small_n ? 32 : 64;
    int splits = g.splits;
    if (splits <= 0) {
        // weight-gradient shape: few output tiles, very long inner dimension -> fill ~2 waves
        int64_t tiles = ceil_div(g.M, bm) * ceil_div(g.N, bn);
        int64_t want = ceil_div(2 * kNumSMs, tiles);
        int64_t maxs = ceil_div(g.K, 8 * BK);
        splits = (int)(want < 1 ? 1 : (want > maxs ? maxs : want));
        if (splits < 1) splits = 1;
    }
    INTEL_REQUIRE(splits == 1 || (g.accumulate == 2 && !g.relu_out && !g.mask), INTEL_ERR_ARG,
                  "gemm: split-K needs atomic accumulation and a linear epilogue");
    INTEL_REQUIRE(splits <= 65535, INTEL_ERR_ARG, "gemm: too many splits");
    d.splits = splits;
    d.kchunk = ceil_div(ceil_div(g.K, splits), BK) * BK;
    if (d.kchunk <= 0) d.kchunk = BK;
    if (bm == 32 && bn == 32) return launch_tile<32, 32>(d, g.a_t, g.b_t, s);
    if (bm == 32 && bn == 64) return launch_tile<32, 64>(d, g.a_t, g.b_t, s);
    if (bm == 128) return launch_tile<128, 32>(d, g.a_t, g.b_t, s);
    return launch_tile<64, 64>(d, g.a_t, g.b_t, s);
}

int linear(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* W, int64_t ldw,
           const float* bias, float* C, int64_t ldc, cudaStream_t s, bool relu_a, bool relu_out,
           const float* add, int64_t ldadd) {
    Gemm g;
    g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = W; g.ldb = ldw; g.C = C; g.ldc = ldc;
    g.bias = bias; g.relu_a = relu_a; g.relu_out = relu_out; g.add = add; g.ldadd = ldadd;
    return gemm(g, s);
}

int linear_dx(int64_t M, int64_t N, int64_t K, const float* dY, int64_t lddy, const float* W, int64_t ldw,
              float* dX, int64_t lddx, cudaStream_t s, int accumulate, const float* mask, int64_t ldmask) {
    Gemm g;   // dX[M,K] = dY[M,N] * W[N,K]: inner dimension N, B operand stored [inner][out]
    g.M = M; g.N = K; g.K = N; g.A = dY; g.lda = lddy; g.B = W; g.ldb = ldw; g.b_t = true;
    g.C = dX; g.ldc = lddx; g.accumulate = accumulate; g.mask = mask; g.ldmask = ldmask;
    return gemm(g, s);
}

int linear_dw(int64_t M, int64_t N, int64_t K, const float* dY, int64_t lddy, const float* X, int64_t ldx,
              float* dW, int64_t lddw, float* db, cudaStream_t s, bool relu_x) {
    if (dW) {
        Gemm g;   // dW[N,K] += sum_m dY[m,n] X[m,k]: both operands stored [inner][out]
        g.M = N; g.N = K; g.K = M; g.A = dY; g.lda = lddy; g.a_t = true; g.B = X; g.ldb = ldx; g.b_t = true;
        g.C = dW; g.ldc = lddw; g.accumulate = 2; g.splits = 0; g.relu_b = relu_x;
        INTEL_TRY(gemm(g, s));
    }
    if (db) INTEL_TRY(colsum(M, N, dY, lddy, db, s));
    return INTEL_OK;
}

// ---- column sums (bias gradients): out[n] += sum_m X[m, n] ----
__global__ void __launch_bounds__(256) colsum_kernel(int64_t M, int64_t N, const float* X, int64_t ld, float* out,
                                                     int64_t rows_per_block) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t n = (int64_t)blockIdx.x * 32 + tx;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
    float acc = 0.f;
    if (n < N)
        for (int64_t r = r0 + ty; r < r1; r += 8) acc += X[r * ld + n];
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        atomicAdd(out + n, t);
    }
}

int colsum(int64_t M, int64_t N, const float* X, int64_t ld, float* out, cudaStream_t s) {
    if (M <= 0 || N <= 0) return INTEL_OK;
    int64_t col_blocks = ceil_div(N, 32);
    int64_t want_rows = ceil_div(4 * kNumSMs, col_blocks);
    int64_t row_blocks = ceil_div(M, 64);
    if (row_blocks > want_rows) row_blocks = want_rows;
    if (row_blocks < 1) row_blocks = 1;
    int64_t rpb = ceil_div(M, row_blocks);
    row_blocks = ceil_div(M, rpb);
    dim3 grid((unsigned)col_blocks, (unsigned)row_blocks);
    LAUNCH(colsum_kernel, grid, dim3(256), 0, s, M, N, X, ld, out, rpb);
    return check_launch("colsum", 4.0 * M * N, (double)M * N);
}

}  // namespace intel
