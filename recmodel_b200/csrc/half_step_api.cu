// wmf_als_half_step: argument checks and algorithm dispatch (SIMT reference path /
// tcgen05 path).
#include "common.cuh"
#include "half_step.cuh"

using namespace wmf;

static int resolve_algo(int algo, int f, int bias) {
    if (algo == WMF_ALGO_AUTO) return tc_half_step_supported(f, bias) ? WMF_ALGO_TCGEN05 : WMF_ALGO_SIMT;
    return algo;
}

namespace wmf { int wmf_tc_split_length(); }

extern "C" {

int wmf_als_row_split_entries(void) { return wmf_tc_split_length(); }

int wmf_als_half_step_supports(int algo, int f, int bias) {
    if (f <= 0 || f > WMF_MAX_F || (bias && f < 2)) return 0;
    if (algo == WMF_ALGO_TCGEN05 || algo == WMF_ALGO_TCGEN05_DIRECT) return tc_half_step_supported(f, bias) ? 1 : 0;
    return (algo == WMF_ALGO_SIMT || algo == WMF_ALGO_AUTO) ? 1 : 0;
}

int wmf_als_dual_max_entries(void) { return tc_dual_max_entries(); }

static size_t half_step_workspace(int64_t rows, int64_t cols, int f, int algo, int64_t segments) {
    if (f <= 0 || f > WMF_MAX_F || rows < 0 || cols < 0) return 0;
    size_t a = simt_half_step_workspace_bytes(f);
    size_t b = 0;
    if (algo != WMF_ALGO_SIMT && tc_half_step_supported(f, 0)) b = tc_half_step_workspace_bytes(rows, cols, f, 0, segments);
    return a > b ? a : b;
}

size_t wmf_als_half_step_workspace_bytes(int64_t rows, int64_t cols, int f, int algo) {
    return half_step_workspace(rows, cols, f, algo, -1);
}

size_t wmf_als_half_step_workspace_bytes_split(int64_t rows, int64_t cols, int f, int algo, int64_t segments) {
    return half_step_workspace(rows, cols, f, algo, segments < 0 ? 0 : segments);
}

int wmf_als_half_step_status(const void* ws, int* flags_host, int* fixup_rows_host, void* stream) {
    WMF_REQUIRE(ws != nullptr, "wmf_als_half_step_status: null workspace");
    int hdr[9];
    WMF_CUDA(cudaMemcpyAsync(hdr, ws, sizeof(hdr), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    WMF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (flags_host) *flags_host = hdr[1];
    if (fixup_rows_host) *fixup_rows_host = hdr[8];
    return WMF_OK;
}

int wmf_als_half_step_used_fallback(const void* ws, int* any_row_host, void* stream) {
    WMF_REQUIRE(ws != nullptr && any_row_host != nullptr, "wmf_als_half_step_used_fallback: null pointer");
    WMF_CUDA(cudaMemcpyAsync(any_row_host, reinterpret_cast<const int*>(ws) + 12, sizeof(int), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    WMF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return WMF_OK;
}

int wmf_als_half_step(const int64_t* indptr, const int32_t* indices, const float* data, int64_t rows, int64_t cols,
                      const int32_t* row_order, int64_t order_len, const float* Y, int64_t ldy, int f,
                      const float* G, int bias,
                      float* X, int64_t ldx, int algo, void* ws, size_t ws_bytes, void* stream) {
    WMF_REQUIRE(f > 0 && f <= WMF_MAX_F && (!bias || f >= 2), "wmf_als_half_step: f=%d outside 1..%d", f, WMF_MAX_F);
    WMF_REQUIRE(rows >= 0 && rows < (1ll << 31), "wmf_als_half_step: rows=%lld out of range", (long long)rows);
    WMF_REQUIRE(cols >= 0 && cols < (1ll << 31), "wmf_als_half_step: cols=%lld out of range", (long long)cols);
    if (rows == 0) return WMF_OK;
    WMF_REQUIRE(indptr && Y && G && X && ldy >= f && ldx >= f, "wmf_als_half_step: null pointer or short leading dimension");
    WMF_REQUIRE(algo == WMF_ALGO_AUTO || algo == WMF_ALGO_SIMT || algo == WMF_ALGO_TCGEN05 || algo == WMF_ALGO_TCGEN05_DIRECT,
                "wmf_als_half_step: unknown algo %d", algo);
    const int chosen = resolve_algo(algo, f, bias);
    HalfStepParams p{};
    WMF_REQUIRE(row_order == nullptr || order_len >= rows, "wmf_als_half_step: schedule shorter than rows");
    p.indptr = indptr; p.indices = indices; p.data = data; p.rows = rows; p.row_order = row_order;
    p.sched_len = row_order ? order_len : rows;
    p.Y = Y; p.ldy = ldy; p.f = f; p.G = G; p.bias = bias; p.X = X; p.ldx = ldx; p.cols = cols;
    if (chosen == WMF_ALGO_TCGEN05 || chosen == WMF_ALGO_TCGEN05_DIRECT) {
        p.cg_maxit = chosen == WMF_ALGO_TCGEN05_DIRECT ? 0 : tc_cg_max_products();
        if (!tc_half_step_supported(f, bias)) {
            set_error("wmf_als_half_step: the tcgen05 path takes f <= 256 (f=%d, bias=%d)", f, bias);
            return WMF_ERR_UNSUPPORTED;
        }
        return tc_half_step(p, ws, ws_bytes, (cudaStream_t)stream);
    }
    return simt_half_step(p, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
