// K3: fused predict + error reduction (predict: wmf_model.py:205-211; eval_prec:
// base_model.py:163-176), and R8: element-wise pair scores (WMF.predict).
//
// HBM/L2-bound integer+float streaming: per stored entry one user row and one item row
// (2*4f B) plus 12 B of index/value. Every prediction is formed in NumPy's exact rounding
// order (common.cuh np_score) by a group of 8 lanes, so yhat is bit-identical to the
// reference's and only the final mean differs (we sum in double, deterministically:
// per-CTA partials reduced in a fixed order).
#include "common.cuh"

namespace wmf {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_CHUNK = 4096;  // stored entries per CTA

__global__ __launch_bounds__(LOSS_THREADS) void sddmm_loss_kernel(const int64_t* __restrict__ indptr,
                                                                  const int32_t* __restrict__ indices,
                                                                  const float* __restrict__ data, int64_t rows,
                                                                  int64_t nnz, const float* __restrict__ U, int64_t ldu,
                                                                  const float* __restrict__ V, int64_t ldv, int f,
                                                                  int bias, double* __restrict__ partial) {
    __shared__ int64_t s_rlo, s_rhi;
    __shared__ double s_red[3][LOSS_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    const int gl = lane & 7;
    const unsigned gmask = 0xFFu << (lane & 24);
    const int64_t e0 = (int64_t)blockIdx.x * LOSS_CHUNK;
    int64_t e1 = e0 + LOSS_CHUNK;
    if (e1 > nnz) e1 = nnz;
    if (tid == 0) {
        // first row whose range contains e0 / e1-1  (largest r with indptr[r] <= e)
        int64_t lo = 0, hi = rows;
        while (hi - lo > 1) { int64_t m = (lo + hi) >> 1; if (indptr[m] <= e0) lo = m; else hi = m; }
        s_rlo = lo;
        int64_t lo2 = lo, hi2 = rows;
        while (hi2 - lo2 > 1) { int64_t m = (lo2 + hi2) >> 1; if (indptr[m] <= e1 - 1) lo2 = m; else hi2 = m; }
        s_rhi = lo2;
    }
    __syncthreads();
    const int64_t rlo = s_rlo, rhi = s_rhi;
    double sq = 0.0, ab = 0.0, cnt = 0.0;
    const int group = tid >> 3;
    for (int64_t e = e0 + group; e < e1; e += LOSS_THREADS / 8) {
        const float r = data[e];
        if (r == 0.0f) continue;  // nonzero() drops explicit zeros (base_model.py:163); group-uniform
        int64_t lo = rlo, hi = rhi + 1;
        while (hi - lo > 1) { int64_t m = (lo + hi) >> 1; if (indptr[m] <= e) lo = m; else hi = m; }
        const float* u = U + lo * ldu;
        const float* v = V + (int64_t)indices[e] * ldv;
        const float yhat = np_score(u, v, f, bias, gl, gmask);
        if (gl == 0) {
            const float d = __fsub_rn(r, yhat);  // float32 difference like the reference
            sq += (double)d * (double)d;
            ab += fabs((double)d);
            cnt += 1.0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        ab += __shfl_xor_sync(0xffffffffu, ab, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { s_red[0][tid >> 5] = sq; s_red[1][tid >> 5] = ab; s_red[2][tid >> 5] = cnt; }
    __syncthreads();
    if (tid < 3) {
        double s = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) s += s_red[tid][w];
        partial[(size_t)blockIdx.x * 3 + tid] = s;
    }
}

__global__ void loss_reduce_kernel(const double* __restrict__ partial, int64_t nparts, double* __restrict__ out3) {
    // one warp per output; fixed lane-strided order then a fixed shuffle tree
    const int which = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int64_t p = lane; p < nparts; p += 32) s += partial[p * 3 + which];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out3[which] = s;
}

__global__ __launch_bounds__(256) void predict_pairs_kernel(const int64_t* __restrict__ users, int64_t user_stride,
                                                            const int64_t* __restrict__ items, int64_t n,
                                                            const float* __restrict__ U, int64_t ldu,
                                                            const float* __restrict__ V, int64_t ldv, int f, int bias,
                                                            float* __restrict__ out) {
    const int lane = threadIdx.x & 31, gl = lane & 7;
    const unsigned gmask = 0xFFu << (lane & 24);
    int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 3;
    for (; g < n; g += stride) {
        const float* u = U + users[g * user_stride] * ldu;
        const float* v = V + items[g] * ldv;
        const float s = np_score(u, v, f, bias, gl, gmask);
        if (gl == 0) out[g] = s;
    }
}

}  // namespace wmf

using namespace wmf;

extern "C" {

size_t wmf_sddmm_loss_workspace_bytes(int64_t nnz) {
    if (nnz <= 0) return 3 * sizeof(double);
    return (size_t)((nnz + LOSS_CHUNK - 1) / LOSS_CHUNK) * 3 * sizeof(double);
}

int wmf_sddmm_loss(const int64_t* indptr, const int32_t* indices, const float* data, int64_t rows, int64_t nnz,
                   const float* U, int64_t ldu, const float* V, int64_t ldv, int f, int bias, double* out3, void* ws,
                   size_t ws_bytes, void* stream) {
    WMF_REQUIRE(f > 0 && f <= WMF_MAX_F && (!bias || f >= 2), "wmf_sddmm_loss: f=%d out of range", f);
    WMF_REQUIRE(rows >= 0 && nnz >= 0 && indptr && out3, "wmf_sddmm_loss: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (nnz == 0 || rows == 0) {
        WMF_CUDA(cudaMemsetAsync(out3, 0, 3 * sizeof(double), st));
        return WMF_OK;
    }
    WMF_REQUIRE(indices && data && U && V, "wmf_sddmm_loss: null argument");
    int64_t parts = (nnz + LOSS_CHUNK - 1) / LOSS_CHUNK;
    if (ws == nullptr || ws_bytes < (size_t)parts * 3 * sizeof(double)) {
        set_error("wmf_sddmm_loss: workspace %zu < %zu", ws_bytes, (size_t)parts * 3 * sizeof(double));
        return WMF_ERR_WORKSPACE;
    }
    sddmm_loss_kernel<<<(unsigned)parts, LOSS_THREADS, 0, st>>>(indptr, indices, data, rows, nnz, U, ldu, V, ldv, f,
                                                                bias, (double*)ws);
    WMF_LAUNCH_CHECK("sddmm_loss_kernel");
    loss_reduce_kernel<<<1, 96, 0, st>>>((const double*)ws, parts, out3);
    WMF_LAUNCH_CHECK("loss_reduce_kernel");
    return WMF_OK;
}

int wmf_predict_pairs(const int64_t* users, int64_t user_stride, const int64_t* items, int64_t n, const float* U,
                      int64_t ldu, const float* V, int64_t ldv, int f, int bias, float* out, void* stream) {
    WMF_REQUIRE(f > 0 && f <= WMF_MAX_F && (!bias || f >= 2), "wmf_predict_pairs: f=%d out of range", f);
    WMF_REQUIRE(n >= 0 && (n == 0 || (users && items && out && U && V)), "wmf_predict_pairs: null argument");
    WMF_REQUIRE(user_stride == 0 || user_stride == 1, "wmf_predict_pairs: user_stride must be 0 or 1");
    if (n == 0) return WMF_OK;
    int64_t blocks = (n * 8 + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    predict_pairs_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(users, user_stride, items, n, U, ldu, V,
                                                                            ldv, f, bias, out);
    WMF_LAUNCH_CHECK("predict_pairs_kernel");
    return WMF_OK;
}

}  // extern "C"
