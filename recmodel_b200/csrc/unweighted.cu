// K7: the unweighted (closed-form) half-step  X = R * (inv(Y^T Y + gamma I) * Y^T)^T
// (wmf_model.py:85, :88): a small dense inverse, a skinny dense product and an SpMM.
#include "common.cuh"

namespace wmf {

// Gauss-Jordan with partial pivoting on [G | I] in a global-memory slab, one CTA.
// f <= 320 -> the slab is <= 0.8 MB and stays in L2; the work is f^3 flops (negligible).
__global__ __launch_bounds__(1024) void inverse_kernel(const float* __restrict__ G, int f, float* __restrict__ M,
                                                       float* __restrict__ out) {
    __shared__ int s_piv;
    __shared__ float s_best[32];
    __shared__ int s_idx[32];
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int w = 2 * f;
    for (int e = tid; e < f * w; e += NT) {
        int i = e / w, j = e % w;
        M[e] = j < f ? G[i * f + j] : (j - f == i ? 1.0f : 0.0f);
    }
    __syncthreads();
    for (int k = 0; k < f; ++k) {
        float best = -1.0f;
        int bi = k;
        for (int i = k + tid; i < f; i += NT) {
            float v = fabsf(M[i * w + k]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { s_best[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            float b = s_best[0];
            int bi2 = s_idx[0];
            for (int q = 1; q < NT / 32; ++q)
                if (s_best[q] > b || (s_best[q] == b && s_idx[q] < bi2)) { b = s_best[q]; bi2 = s_idx[q]; }
            s_piv = bi2;
        }
        __syncthreads();
        const int p = s_piv;
        if (p != k)
            for (int j = tid; j < w; j += NT) { float t = M[k * w + j]; M[k * w + j] = M[p * w + j]; M[p * w + j] = t; }
        __syncthreads();
        const float inv = 1.0f / M[k * w + k];
        __syncthreads();
        for (int j = tid; j < w; j += NT) M[k * w + j] *= inv;
        __syncthreads();
        // eliminate column k from every other row; column k itself is read-only this phase
        for (int e = tid; e < f * w; e += NT) {
            int i = e / w, j = e % w;
            if (i != k && j != k) M[e] = fmaf(-M[i * w + k], M[k * w + j], M[e]);
        }
        __syncthreads();
        for (int i = tid; i < f; i += NT) if (i != k) M[i * w + k] = 0.0f;
        __syncthreads();
    }
    for (int e = tid; e < f * f; e += NT) out[e] = M[(e / f) * w + f + (e % f)];
}

// W[r][i] = sum_c Y[r][c] * M[i][c]
__global__ __launch_bounds__(256) void right_multiply_kernel(const float* __restrict__ Y, int64_t n, int64_t ldy,
                                                             const float* __restrict__ M, int f,
                                                             float* __restrict__ W, int64_t ldw) {
    __shared__ float sy[64][33];
    __shared__ float sm[64][33];
    const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    const int i0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int c0 = 0; c0 < f; c0 += 32) {
        for (int e = tid; e < 64 * 32; e += 256) {
            int rr = e / 32, c = e % 32;
            sy[rr][c] = (r0 + rr < n && c0 + c < f) ? Y[(r0 + rr) * ldy + c0 + c] : 0.f;
            sm[rr][c] = (i0 + rr < f && c0 + c < f) ? M[(size_t)(i0 + rr) * f + c0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(sy[ty * 4 + a][c], sm[tx * 4 + b][c], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int64_t r = r0 + ty * 4 + a;
            int i = i0 + tx * 4 + b;
            if (r < n && i < f) W[r * ldw + i] = acc[a][b];
        }
}

// X[r][c] = sum_j data_j * W[indices_j][c], stored order, separate multiply and add like
// SciPy's csr_matvecs axpy (no FMA). One warp per row.
__global__ __launch_bounds__(256) void spmm_kernel(const int64_t* __restrict__ indptr,
                                                   const int32_t* __restrict__ indices,
                                                   const float* __restrict__ data, int64_t rows,
                                                   const float* __restrict__ W, int64_t ldw, int f,
                                                   float* __restrict__ X, int64_t ldx) {
    const int lane = threadIdx.x & 31;
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; r < rows; r += stride) {
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        for (int c0 = 0; c0 < f; c0 += 32 * 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int64_t e = lo; e < hi; ++e) {
                const float d = data[e];
                const float* w = W + (int64_t)indices[e] * ldw;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int c = c0 + q * 32 + lane;
                    if (c < f) acc[q] = __fadd_rn(acc[q], __fmul_rn(d, w[c]));
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int c = c0 + q * 32 + lane;
                if (c < f) X[r * ldx + c] = acc[q];
            }
        }
    }
}

}  // namespace wmf

using namespace wmf;

extern "C" {

size_t wmf_inverse_workspace_bytes(int f) { return f > 0 ? (size_t)f * 2 * f * sizeof(float) : 0; }

int wmf_inverse(const float* G, int f, float* Ginv, void* ws, size_t ws_bytes, void* stream) {
    WMF_REQUIRE(f > 0 && f <= WMF_MAX_F && G && Ginv, "wmf_inverse: bad arguments");
    if (ws == nullptr || ws_bytes < wmf_inverse_workspace_bytes(f)) {
        set_error("wmf_inverse: workspace %zu < %zu", ws_bytes, wmf_inverse_workspace_bytes(f));
        return WMF_ERR_WORKSPACE;
    }
    inverse_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(G, f, (float*)ws, Ginv);
    WMF_LAUNCH_CHECK("inverse_kernel");
    return WMF_OK;
}

int wmf_dense_right_multiply(const float* Y, int64_t n, int64_t ldy, const float* M, int f, float* W, int64_t ldw,
                             void* stream) {
    WMF_REQUIRE(f > 0 && f <= WMF_MAX_F && n >= 0 && M && (n == 0 || (Y && W)), "wmf_dense_right_multiply: bad arguments");
    if (n == 0) return WMF_OK;
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((f + 63) / 64));
    right_multiply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Y, n, ldy, M, f, W, ldw);
    WMF_LAUNCH_CHECK("right_multiply_kernel");
    return WMF_OK;
}

int wmf_spmm(const int64_t* indptr, const int32_t* indices, const float* data, int64_t rows, const float* W,
             int64_t ldw, int f, float* X, int64_t ldx, void* stream) {
    WMF_REQUIRE(f > 0 && f <= WMF_MAX_F && rows >= 0 && (rows == 0 || (indptr && X)), "wmf_spmm: bad arguments");
    if (rows == 0) return WMF_OK;
    int64_t blocks = (rows * 32 + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    spmm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(indptr, indices, data, rows, W, ldw, f, X, ldx);
    WMF_LAUNCH_CHECK("spmm_kernel");
    return WMF_OK;
}

}  // extern "C"
