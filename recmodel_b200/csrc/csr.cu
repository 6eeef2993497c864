// N1: CSR transpose on the device, replacing `count_mat.T.tocsr()` (wmf_model.py:128).
//
// SciPy's conversion is a counting sort of the entries by column that keeps the CSR traversal order inside every
// output row, i.e. ascending original row ids (whatever the column order inside the input rows, duplicates kept).
// Here: a stable least-significant-digit radix sort of (column, row id, value) by column, 8 bits per pass
// (ceil(log2(cols) / 8) passes: two for the 26 744 items of ML-20M), each pass
//     tile histograms (4096 entries per tile)  ->  exclusive scan in digit-major order  ->  stable scatter
// (ranks inside a tile from per-warp digit histograms and match_any ballots), plus the output row pointer from a
// column histogram. HBM-bound integer work: 3 x 12 B read + 12 B written per entry and pass. Transposing twice
// canonicalises a matrix (sorted column indices inside every row).
#include "common.cuh"

namespace wmf {

namespace {

constexpr int TILE = 4096, TPB = 256, IPT = TILE / TPB, NWARP = TPB / 32;   // entries per tile / threads / entries per thread

// row id of every entry: one warp per row
__global__ void expand_rows_kernel(const int64_t* __restrict__ indptr, int64_t rows, int32_t* __restrict__ row_of) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    for (int64_t i = indptr[r] + lane; i < indptr[r + 1]; i += 32) row_of[i] = (int32_t)r;
}

__global__ void column_hist_kernel(const int32_t* __restrict__ cols, int64_t nnz, uint32_t* __restrict__ count) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) atomicAdd(count + cols[i], 1u);
}

// hist[d * ntiles + tile] = entries of the tile whose digit is d
__global__ void __launch_bounds__(TPB) tile_hist_kernel(const int32_t* __restrict__ keys, int64_t nnz, int shift,
                                                        uint32_t* __restrict__ hist, int64_t ntiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * TILE;
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        const int64_t i = base + r * TPB + threadIdx.x;
        if (i < nnz) atomicAdd(&h[((uint32_t)keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// ---- exclusive scan of a uint32 array (total < 2^32), three launches: block sums, scan of the sums, final pass
constexpr int SCAN_TPB = 256, SCAN_IPT = 16, SCAN_TILE = SCAN_TPB * SCAN_IPT;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < SCAN_TPB / 32 ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < SCAN_TPB / 32) warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    total = warp_sums[SCAN_TPB / 32 - 1];
    const uint32_t before = w > 0 ? warp_sums[w - 1] : 0u;
    __syncthreads();
    return before + x - v;
}

__global__ void __launch_bounds__(SCAN_TPB) scan_sums_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t ws[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) if (base + k < n) v += in[base + k];
    uint32_t total;
    block_exclusive_scan(v, ws, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_TPB) scan_top_kernel(uint32_t* __restrict__ sums, int64_t nb) {  // one block, in place
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < nb; b0 += SCAN_TPB) {
        const int64_t i = b0 + threadIdx.x;
        const uint32_t v = i < nb ? sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, ws, total);
        if (i < nb) sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
}

template <typename OUT>
__global__ void __launch_bounds__(SCAN_TPB) scan_final_kernel(const uint32_t* __restrict__ in, int64_t n, const uint32_t* __restrict__ sums,
                                                               OUT* __restrict__ out, int write_total) {
    __shared__ uint32_t ws[32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_IPT;
    uint32_t x[SCAN_IPT], v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) { x[k] = base + k < n ? in[base + k] : 0u; v += x[k]; }
    uint32_t total;
    uint32_t run = sums[blockIdx.x] + block_exclusive_scan(v, ws, total);
#pragma unroll
    for (int k = 0; k < SCAN_IPT; ++k) {
        if (base + k < n) out[base + k] = (OUT)run;
        run += x[k];
        if (write_total && base + k == n - 1) out[n] = (OUT)run;
    }
}

// stable scatter of one tile: entry i goes to offs[d * ntiles + tile] + (rank of i among the tile's entries with digit d)
__global__ void __launch_bounds__(TPB) tile_scatter_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ rows_in,
                                                           const float* __restrict__ vals_in, int64_t nnz, int shift,
                                                           const uint32_t* __restrict__ offs, int64_t ntiles,
                                                           int32_t* __restrict__ keys_out, int32_t* __restrict__ rows_out,
                                                           float* __restrict__ vals_out) {
    __shared__ uint32_t whist[NWARP][256];   // per-warp digit counts, then running offsets
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NWARP * 256; i += TPB) (&whist[0][0])[i] = 0;
    __syncthreads();
    // warp w owns entries [base + w * 512, +512) in rounds of 32 consecutive entries: lane order = entry order
    const int64_t wbase = (int64_t)blockIdx.x * TILE + (int64_t)w * (TILE / NWARP);
    int32_t key[IPT];
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        key[r] = i < nnz ? keys[i] : -1;
    }
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        const bool live = key[r] >= 0;
        const uint32_t d = ((uint32_t)key[r] >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, live ? d : 256u + lane);
        if (live && (peers & ((1u << lane) - 1)) == 0) whist[w][d] += __popc(peers);   // lowest lane of each digit group
        __syncwarp();
    }
    __syncthreads();
    // digit d: exclusive prefix over the warps + the tile's global offset
    {
        const int d = threadIdx.x;   // TPB == 256 digits
        uint32_t run = offs[(int64_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < NWARP; ++ww) { const uint32_t c = whist[ww][d]; whist[ww][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        const bool live = key[r] >= 0;
        const uint32_t d = ((uint32_t)key[r] >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, live ? d : 256u + lane);
        const unsigned lower = peers & ((1u << lane) - 1);
        if (live) {
            const uint32_t pos = whist[w][d] + __popc(lower);
            keys_out[pos] = key[r];
            rows_out[pos] = rows_in[i];
            vals_out[pos] = vals_in[i];
        }
        __syncwarp();
        if (live && lower == 0) whist[w][d] += __popc(peers);
        __syncwarp();
    }
}

int scan_u32(const uint32_t* in, int64_t n, uint32_t* sums, void* out, bool out64, bool write_total, cudaStream_t st) {
    const int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_sums_kernel<<<(unsigned)nb, SCAN_TPB, 0, st>>>(in, n, sums);
    WMF_LAUNCH_CHECK("scan_sums_kernel");
    scan_top_kernel<<<1, SCAN_TPB, 0, st>>>(sums, nb);
    WMF_LAUNCH_CHECK("scan_top_kernel");
    if (out64) scan_final_kernel<int64_t><<<(unsigned)nb, SCAN_TPB, 0, st>>>(in, n, sums, reinterpret_cast<int64_t*>(out), write_total ? 1 : 0);
    else scan_final_kernel<uint32_t><<<(unsigned)nb, SCAN_TPB, 0, st>>>(in, n, sums, reinterpret_cast<uint32_t*>(out), write_total ? 1 : 0);
    WMF_LAUNCH_CHECK("scan_final_kernel");
    return WMF_OK;
}

struct TrLayout { size_t off_rows[2], off_keys[2], off_vals, off_hist, off_sums, off_count, total; };
TrLayout tr_layout(int64_t cols, int64_t nnz) {
    TrLayout L;
    const int64_t ntiles = (nnz + TILE - 1) / TILE;
    const size_t e4 = align_up((size_t)(nnz > 0 ? nnz : 1) * 4, 256);
    size_t o = 0;
    L.off_rows[0] = o; o += e4;
    L.off_rows[1] = o; o += e4;
    L.off_keys[0] = o; o += e4;
    L.off_keys[1] = o; o += e4;
    L.off_vals = o; o += e4;
    L.off_hist = o; o += align_up((size_t)256 * (size_t)(ntiles > 0 ? ntiles : 1) * 4, 256);
    const int64_t hist_n = 256 * ntiles > cols + 1 ? 256 * ntiles : cols + 1;
    L.off_sums = o; o += align_up((size_t)((hist_n + SCAN_TILE - 1) / SCAN_TILE + 1) * 4, 256);
    L.off_count = o; o += align_up((size_t)(cols + 1) * 4, 256);
    L.total = o;
    return L;
}

}  // namespace

}  // namespace wmf

using namespace wmf;

extern "C" {

size_t wmf_csr_transpose_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz) {
    (void)rows;
    return tr_layout(cols, nnz).total;
}

int wmf_csr_transpose(const int64_t* indptr, const int32_t* indices, const float* data, int64_t rows, int64_t cols, int64_t nnz,
                      int64_t* out_indptr, int32_t* out_indices, float* out_data, void* ws, size_t ws_bytes, void* stream) {
    WMF_REQUIRE(rows >= 0 && cols >= 0 && nnz >= 0 && rows < (1ll << 31) && cols < (1ll << 31) && nnz < (1ll << 32) - TILE,
                "wmf_csr_transpose: shape out of range (%lld x %lld, %lld entries)", (long long)rows, (long long)cols, (long long)nnz);
    WMF_REQUIRE(indptr && out_indptr && (nnz == 0 || (indices && data && out_indices && out_data)), "wmf_csr_transpose: null pointer");
    const TrLayout L = tr_layout(cols, nnz);
    if (ws == nullptr || ws_bytes < L.total) {
        set_error("wmf_csr_transpose: workspace %zu < %zu", ws_bytes, L.total);
        return WMF_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* base = reinterpret_cast<char*>(ws);
    uint32_t* count = reinterpret_cast<uint32_t*>(base + L.off_count);
    uint32_t* sums = reinterpret_cast<uint32_t*>(base + L.off_sums);
    uint32_t* hist = reinterpret_cast<uint32_t*>(base + L.off_hist);
    // ---- output row pointer: column histogram + exclusive scan (int64)
    WMF_CUDA(cudaMemsetAsync(count, 0, (size_t)(cols + 1) * 4, st));
    if (nnz > 0) {
        int64_t blocks = (nnz + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        column_hist_kernel<<<(unsigned)blocks, 256, 0, st>>>(indices, nnz, count);
        WMF_LAUNCH_CHECK("column_hist_kernel");
    }
    int rc = scan_u32(count, cols + 1, sums, out_indptr, true, false, st);   // out_indptr[cols] = total: count[cols] = 0
    if (rc) return rc;
    if (nnz == 0) return WMF_OK;
    // ---- entries: row ids, then stable LSD radix sort by column
    int32_t* rows_buf[2] = {reinterpret_cast<int32_t*>(base + L.off_rows[0]), reinterpret_cast<int32_t*>(base + L.off_rows[1])};
    int32_t* keys_buf[2] = {reinterpret_cast<int32_t*>(base + L.off_keys[0]), reinterpret_cast<int32_t*>(base + L.off_keys[1])};
    float* vals_buf = reinterpret_cast<float*>(base + L.off_vals);
    expand_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(indptr, rows, rows_buf[0]);
    WMF_LAUNCH_CHECK("expand_rows_kernel");
    int bits = 1;
    while ((1ll << bits) < cols) ++bits;
    const int passes = (bits + 7) / 8;
    const int64_t ntiles = (nnz + TILE - 1) / TILE;
    const int32_t* kin = indices;
    const int32_t* rin = rows_buf[0];
    const float* vin = data;
    for (int ps = 0; ps < passes; ++ps) {
        const bool last = ps == passes - 1;
        int32_t* kout = keys_buf[ps & 1];
        int32_t* rout = last ? out_indices : rows_buf[(ps + 1) & 1];
        // values ping-pong between the scratch and the output so that the last pass lands in out_data
        float* vout = ((passes - 1 - ps) & 1) ? vals_buf : out_data;
        tile_hist_kernel<<<(unsigned)ntiles, TPB, 0, st>>>(kin, nnz, 8 * ps, hist, ntiles);
        WMF_LAUNCH_CHECK("tile_hist_kernel");
        rc = scan_u32(hist, 256 * ntiles, sums, hist, false, false, st);
        if (rc) return rc;
        tile_scatter_kernel<<<(unsigned)ntiles, TPB, 0, st>>>(kin, rin, vin, nnz, 8 * ps, hist, ntiles, kout, rout, vout);
        WMF_LAUNCH_CHECK("tile_scatter_kernel");
        kin = kout; rin = rout; vin = vout;
    }
    return WMF_OK;
}

}  // extern "C"
