// K1: G = Y^T Y + lambda I for the fixed-side factors (wmf_model.py:215,244,258,332).
//
// FP32 CUDA cores. The flop count is n*f^2*2 (4.5 GFLOP for 138k x 128), far below a
// half-step's, and the read of Y (n*f*4 B) is the algorithmic traffic, so this kernel is
// HBM-bound by design: each CTA streams a contiguous chunk of rows once through shared
// memory and keeps a 64x64 output block per (bi,bj) in registers (4x4 per thread).
// Accuracy: with all-positive factors (the U[0,1) initialisation) G is a huge rank-one term
// plus a small well-conditioned part, and the half-step solution is ~1e2..1e3 times more
// sensitive to relative errors in G than to anything else. So FP32 FMAs only run over one
// staged tile of 32 rows; tiles are summed in double, partials are stored and reduced in
// double in a fixed order (bit-reproducible), and G is rounded to fp32 once.
#include "common.cuh"

namespace wmf {

constexpr int GB = 64;        // output block edge
constexpr int GR = 32;        // rows staged per iteration
constexpr int G_THREADS = 256;

__global__ __launch_bounds__(G_THREADS) void gram_partial_kernel(const float* __restrict__ Y, int64_t n, int f,
                                                                 int64_t ldy, int ones_col0, int rows_per_cta,
                                                                 double* __restrict__ partial) {
    __shared__ float sa[GR][GB + 4];
    __shared__ float sb[GR][GB + 4];
    const int nblk = (f + GB - 1) / GB;
    const int bi = blockIdx.y / nblk, bj = blockIdx.y % nblk;
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;
    double dacc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dacc[a][b] = 0.0;

    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    int64_t r1 = r0 + rows_per_cta;
    if (r1 > n) r1 = n;
    for (int64_t base = r0; base < r1; base += GR) {
        // stage GR rows x 64 columns of both column blocks
        for (int e = tid; e < GR * GB; e += G_THREADS) {
            int rr = e / GB, c = e % GB;
            int64_t row = base + rr;
            float va = 0.f, vb = 0.f;
            if (row < r1) {
                int ca = bi * GB + c, cb = bj * GB + c;
                if (ca < f) va = (ones_col0 && ca == 0) ? 1.0f : Y[row * ldy + ca];
                if (cb < f) vb = (ones_col0 && cb == 0) ? 1.0f : Y[row * ldy + cb];
            }
            sa[rr][c] = va;
            sb[rr][c] = vb;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < GR; ++rr) {
            float4 a4 = *reinterpret_cast<const float4*>(&sa[rr][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&sb[rr][tx * 4]);
            float av[4] = {a4.x, a4.y, a4.z, a4.w};
            float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dacc[a][b] += (double)acc[a][b];
        __syncthreads();
    }
    double* out = partial + (size_t)blockIdx.x * f * f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = bi * GB + ty * 4 + a, j = bj * GB + tx * 4 + b;
            if (i < f && j < f) out[(size_t)i * f + j] = dacc[a][b];
        }
}

__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nparts, int f, float lambda,
                                   float* __restrict__ G) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= f * f) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * f * f + e];
    // np.dot(Y.T, Y) is rounded to fp32 before lambda*eye is added (wmf_model.py:215)
    float g = (float)s;
    if (e / f == e % f) g = __fadd_rn(g, lambda);
    G[e] = g;
}

static int gram_parts(int64_t n) {
    int parts = sm_count() * 2;
    int64_t max_parts = (n + GR - 1) / GR;
    if (max_parts < 1) max_parts = 1;
    if (parts > max_parts) parts = (int)max_parts;
    return parts;
}

}  // namespace wmf

using namespace wmf;

extern "C" {

size_t wmf_gram_workspace_bytes(int64_t n, int f) {
    if (n < 0 || f <= 0) return 0;
    return (size_t)gram_parts(n) * f * f * sizeof(double);
}

int wmf_gram(const float* Y, int64_t n, int f, int64_t ldy, float lambda, int ones_col0, float* G, void* ws,
             size_t ws_bytes, void* stream) {
    WMF_REQUIRE(f > 0 && f <= WMF_MAX_F, "wmf_gram: f=%d outside 1..%d", f, WMF_MAX_F);
    WMF_REQUIRE(n >= 0 && G != nullptr && (Y != nullptr || n == 0) && ldy >= f, "wmf_gram: bad arguments");
    size_t need = wmf_gram_workspace_bytes(n, f);
    if (ws_bytes < need || ws == nullptr) {
        set_error("wmf_gram: workspace %zu < %zu", ws_bytes, need);
        return WMF_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int parts = gram_parts(n);
    int rows_per_cta = (int)((n + parts - 1) / parts);
    rows_per_cta = (rows_per_cta + GR - 1) / GR * GR;
    if (rows_per_cta == 0) rows_per_cta = GR;
    int nblk = (f + GB - 1) / GB;
    dim3 grid(parts, nblk * nblk);
    gram_partial_kernel<<<grid, G_THREADS, 0, st>>>(Y, n, f, ldy, ones_col0, rows_per_cta, (double*)ws);
    WMF_LAUNCH_CHECK("gram_partial_kernel");
    gram_reduce_kernel<<<(f * f + 255) / 256, 256, 0, st>>>((const double*)ws, parts, f, lambda, G);
    WMF_LAUNCH_CHECK("gram_reduce_kernel");
    return WMF_OK;
}

}  // extern "C"
