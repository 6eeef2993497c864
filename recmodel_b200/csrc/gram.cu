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
// The decomposition into row blocks depends on n alone (wmf_gram_block_rows), never on the device or on
// who computes a block: a rank of a row-sharded run computes the blocks of its own rows
// (wmf_gram_partials), the block partials are exchanged, and every rank reduces all blocks in block order
// (wmf_gram_reduce) to the same bits a single GPU gets. Only the blocks on and below the diagonal of the
// 64x64 tiling are computed (G is symmetric); the reduction mirrors them.
#include "common.cuh"

namespace wmf {

constexpr int GB = 64;        // output block edge
constexpr int GR = 32;        // rows staged per iteration
constexpr int G_THREADS = 256;
constexpr int G_TARGET_BLOCKS = 128;  // row blocks per Gram (a multiple of GR rows each): 16.8 MB of partials at f = 128

static int64_t gram_block_rows(int64_t n) {
    int64_t b = (n + G_TARGET_BLOCKS - 1) / G_TARGET_BLOCKS;
    b = (b + GR - 1) / GR * GR;
    return b < GR ? GR : b;
}
static int64_t gram_blocks(int64_t n) {
    const int64_t b = gram_block_rows(n);
    return n <= 0 ? 1 : (n + b - 1) / b;
}

// CTA (x = local row block, y = tile pair bi >= bj): partial[(block0 + x)][bi*64.., bj*64..] over the rows
// [x*B, min((x+1)*B, nloc)) of the local slice
__global__ __launch_bounds__(G_THREADS) void gram_partial_kernel(const float* __restrict__ Y, int64_t nloc, int f,
                                                                 int64_t ldy, int ones_col0, int64_t block_rows,
                                                                 int64_t block0, double* __restrict__ partial) {
    __shared__ float sa[GR][GB + 4];
    __shared__ float sb[GR][GB + 4];
    int bi = 0, rem = blockIdx.y;  // pair index -> (bi, bj), bj <= bi
    while (rem > bi) { rem -= bi + 1; ++bi; }
    const int bj = rem;
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;
    double dacc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dacc[a][b] = 0.0;

    const int64_t r0 = (int64_t)blockIdx.x * block_rows;
    int64_t r1 = r0 + block_rows;
    if (r1 > nloc) r1 = nloc;
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
    double* out = partial + (size_t)(block0 + blockIdx.x) * f * f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int i = bi * GB + ty * 4 + a, j = bj * GB + tx * 4 + b;
            if (i < f && j < f) out[(size_t)i * f + j] = dacc[a][b];
        }
}

// G[i][j] = sum over blocks, in block order, of the stored tile entry ((i,j) if tile(i) >= tile(j), else (j,i))
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nparts, int f, float lambda,
                                   float* __restrict__ G) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= f * f) return;
    const int i = e / f, j = e % f;
    const int src = (i / GB >= j / GB) ? e : j * f + i;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * f * f + src];
    // np.dot(Y.T, Y) is rounded to fp32 before lambda*eye is added (wmf_model.py:215)
    float g = (float)s;
    if (i == j) g = __fadd_rn(g, lambda);
    G[e] = g;
}

}  // namespace wmf

using namespace wmf;

extern "C" {

int64_t wmf_gram_block_rows(int64_t n) { return gram_block_rows(n); }
int64_t wmf_gram_blocks(int64_t n) { return gram_blocks(n); }

size_t wmf_gram_workspace_bytes(int64_t n, int f) {
    if (n < 0 || f <= 0) return 0;
    return (size_t)gram_blocks(n) * f * f * sizeof(double);
}

int wmf_gram_partials(const float* Y, int64_t row0, int64_t nloc, int64_t n, int f, int64_t ldy, int ones_col0,
                      void* partials, size_t partials_bytes, void* stream) {
    WMF_REQUIRE(f > 0 && f <= WMF_MAX_F, "wmf_gram_partials: f=%d outside 1..%d", f, WMF_MAX_F);
    WMF_REQUIRE(n >= 0 && nloc >= 0 && row0 >= 0 && row0 + nloc <= n && partials != nullptr && (Y != nullptr || nloc == 0) &&
                    ldy >= f, "wmf_gram_partials: bad arguments");
    if (nloc == 0) return WMF_OK;
    const int64_t B = gram_block_rows(n);
    WMF_REQUIRE(row0 % B == 0 && (row0 + nloc == n || nloc % B == 0),
                "wmf_gram_partials: the slice [%lld, %lld) is not made of whole %lld-row blocks", (long long)row0,
                (long long)(row0 + nloc), (long long)B);
    if (partials_bytes < wmf_gram_workspace_bytes(n, f)) {
        set_error("wmf_gram_partials: buffer %zu < %zu", partials_bytes, wmf_gram_workspace_bytes(n, f));
        return WMF_ERR_WORKSPACE;
    }
    const int nblk = (f + GB - 1) / GB;
    dim3 grid((unsigned)((nloc + B - 1) / B), nblk * (nblk + 1) / 2);
    gram_partial_kernel<<<grid, G_THREADS, 0, (cudaStream_t)stream>>>(Y, nloc, f, ldy, ones_col0, B, row0 / B,
                                                                      (double*)partials);
    WMF_LAUNCH_CHECK("gram_partial_kernel");
    return WMF_OK;
}

int wmf_gram_reduce(const void* partials, int64_t n, int f, float lambda, float* G, void* stream) {
    WMF_REQUIRE(f > 0 && f <= WMF_MAX_F && n >= 0 && partials != nullptr && G != nullptr, "wmf_gram_reduce: bad arguments");
    gram_reduce_kernel<<<(f * f + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const double*)partials, (int)gram_blocks(n), f,
                                                                              lambda, G);
    WMF_LAUNCH_CHECK("gram_reduce_kernel");
    return WMF_OK;
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
    if (n == 0) WMF_CUDA(cudaMemsetAsync(ws, 0, need, (cudaStream_t)stream));  // one empty block
    int rc = wmf_gram_partials(Y, 0, n, n, f, ldy, ones_col0, ws, ws_bytes, stream);
    if (rc != WMF_OK) return rc;
    return wmf_gram_reduce(ws, n, f, lambda, G, stream);
}

}  // extern "C"
