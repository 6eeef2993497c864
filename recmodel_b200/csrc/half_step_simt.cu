// K2 (FP32 CUDA-core version): one ALS half-step, any f <= WMF_MAX_F, with or without biases.
// Replaces the per-row loop of recompute_factors / recompute_factors_bias
// (wmf_model.py:220-239, :337-350). One persistent CTA per resident slot pulls rows from an
// atomic counter (optionally in the caller's longest-first order), builds
//     A = G + sum_j d_j y_j y_j^T   (lower triangle, 4x4 register tiles over KC staged rows)
//     b = sum_j (d_j + 1) y_j
// in shared memory (or an L2-resident slab when f is too wide for 227 KB), and solves in place.
// SPD rows use Cholesky; the bias formula (d~ = d - beta can be negative) and any row whose
// Cholesky meets a non-positive pivot use LU with partial pivoting, as LAPACK sgesv does.
// This kernel is the accuracy reference for the tcgen05 path and the fallback for shapes it
// does not take.
#include "common.cuh"
#include "solve.cuh"
#include "half_step.cuh"

namespace wmf {

constexpr int HS_THREADS = 256;

__device__ __forceinline__ void tile_from_linear(int t, int& ti, int& tj) {
    int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    while (r * (r + 1) / 2 > t) --r;
    ti = r;
    tj = t - r * (r + 1) / 2;
}

template <bool A_GLOBAL>
__global__ __launch_bounds__(HS_THREADS) void als_half_step_simt_kernel(HalfStepParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = p.f, lda = p.lda, FP = p.FP, KC = p.KC;
    float* Ys = smem;                 // KC x FP
    float* Ws = Ys + KC * FP;         // KC x FP
    float* ds = Ws + KC * FP;         // KC (d+1)
    float* dinv = ds + KC;            // f
    int* misc = reinterpret_cast<int*>(dinv + ((f + 3) & ~3));  // [0] next row, [1] pivot
    float* A = A_GLOBAL ? p.slab + (size_t)blockIdx.x * (size_t)(f + 1) * lda
                        : reinterpret_cast<float*>(misc + 4);
    const int T = FP / 4;
    const int ntiles = T * (T + 1) / 2;
    if (p.run_if != nullptr && *p.run_if == 0) return;  // fix-up launch with nothing to fix
    const int64_t sched_len = p.sched_len_dev ? (int64_t)*p.sched_len_dev : p.sched_len;  // device-side fix-up list
    // fix-up of the tcgen05 pipeline: Y holds the whitened factors (ones column already folded in), G the identity,
    // the bias that shifts the weights comes from the original factors
    const bool whitened = p.Yraw != nullptr;

    while (true) {
        if (tid == 0) misc[0] = atomicAdd(p.counter, 1);
        __syncthreads();
        const int64_t r = misc[0];
        __syncthreads();
        if (r >= sched_len) break;
        const int64_t row = p.row_order ? p.row_order[r] : r;
        if (row < 0) continue;  // padding slot of a balanced schedule
        const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
        float* xout = p.X + row * p.ldx;
        if (lo == hi) {  // wmf_model.py:223-225
            for (int c = tid; c < f; c += HS_THREADS) xout[c] = 0.0f;
            continue;
        }
        bool use_lu = p.bias != 0;
        while (true) {
            // The weighted Gram is summed from zero and G is added ONCE at the end, as the reference
            // does (wmf_model.py:237-239): adding small chunks into the large entries of G would round
            // every partial sum at G's magnitude.
            for (int e = tid; e < f * f; e += HS_THREADS) A[(e / f) * lda + (e % f)] = 0.0f;
            // rhs: fp32 FMAs inside one staged chunk, chunks summed in double (a plain sequential
            // fp32 sum over a 2000-entry row is 3x noisier than the reference's sgemv)
            double bacc0 = 0.0, bacc1 = 0.0;
            __syncthreads();
            for (int64_t base = lo; base < hi; base += KC) {
                const int kc = (int)((hi - base) < KC ? (hi - base) : KC);
                for (int k = warp; k < KC; k += HS_THREADS / 32) {
                    float d = 0.f;
                    const float* yrow = p.Y;
                    if (k < kc) {
                        const int64_t col = p.indices[base + k];
                        yrow = p.Y + col * p.ldy;
                        d = p.data[base + k];
                        if (p.bias) d = __fsub_rn(d, whitened ? p.Yraw[col * p.ldraw] : yrow[0]);  // wmf_model.py:343
                    }
                    for (int c = lane; c < FP; c += 32) {
                        float y = 0.f;
                        if (k < kc && c < f) y = (p.bias && c == 0 && !whitened) ? 1.0f : yrow[c];
                        Ys[k * FP + c] = y;
                        Ws[k * FP + c] = d * y;
                    }
                    if (lane == 0) ds[k] = k < kc ? __fadd_rn(d, 1.0f) : 0.f;
                }
                __syncthreads();
                if (tid < f) {
                    float part = 0.f;
                    for (int k = 0; k < KC; ++k) part = fmaf(ds[k], Ys[k * FP + tid], part);
                    bacc0 += (double)part;
                }
                if (tid + HS_THREADS < f) {
                    float part = 0.f;
                    for (int k = 0; k < KC; ++k) part = fmaf(ds[k], Ys[k * FP + tid + HS_THREADS], part);
                    bacc1 += (double)part;
                }
                for (int t = tid; t < ntiles; t += HS_THREADS) {
                    int ti, tj;
                    tile_from_linear(t, ti, tj);
                    float acc[4][4];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
                    for (int k = 0; k < KC; ++k) {
                        float4 w4 = *reinterpret_cast<const float4*>(&Ws[k * FP + ti * 4]);
                        float4 y4 = *reinterpret_cast<const float4*>(&Ys[k * FP + tj * 4]);
                        float wv[4] = {w4.x, w4.y, w4.z, w4.w};
                        float yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(wv[a], yv[b], acc[a][b]);
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            int i = ti * 4 + a, j = tj * 4 + b;
                            if (i < f && j <= i) A[i * lda + j] += acc[a][b];
                        }
                }
                __syncthreads();
            }
            for (int e = tid; e < f * f; e += HS_THREADS) {
                int i = e / f, j = e % f;
                if (j <= i) A[i * lda + j] = __fadd_rn(A[i * lda + j], p.G[e]);
            }
            __syncthreads();
            if (use_lu) {
                if (tid < f) A[tid * lda + f] = (float)bacc0;
                if (tid + HS_THREADS < f) A[(tid + HS_THREADS) * lda + f] = (float)bacc1;
                for (int e = tid; e < f * f; e += HS_THREADS) {
                    int i = e / f, j = e % f;
                    if (j < i) A[j * lda + i] = A[i * lda + j];
                }
                __syncthreads();
                lu_solve_aug<HS_THREADS>(A, lda, f, misc + 1, xout, tid);
                break;
            }
            if (tid < f) A[f * lda + tid] = (float)bacc0;
            if (tid + HS_THREADS < f) A[f * lda + tid + HS_THREADS] = (float)bacc1;
            __syncthreads();
            if (chol_factor_aug<HS_THREADS>(A, lda, f, dinv, tid)) {
                chol_back_solve(A, lda, f, dinv, xout, tid);
                break;
            }
            __syncthreads();
            use_lu = true;  // not positive definite: redo this row the general way
        }
        __syncthreads();
    }
}

struct SimtPlan {
    int lda, FP, KC;
    size_t smem_bytes;
    bool a_global;
    int grid;
    size_t slab_bytes;  // per CTA
};

static SimtPlan simt_plan(int f) {
    SimtPlan pl;
    pl.FP = (f + 3) & ~3;
    pl.lda = (f + 1) | 1;
    pl.KC = 16;
    size_t a_bytes = (size_t)(f + 1) * pl.lda * sizeof(float);
    size_t fixed = ((size_t)2 * pl.KC * pl.FP + pl.KC + ((f + 3) & ~3) + 4) * sizeof(float);
    pl.a_global = fixed + a_bytes > 220 * 1024;
    pl.smem_bytes = fixed + (pl.a_global ? 0 : a_bytes);
    int per_sm = pl.a_global ? 2 : (int)((220 * 1024) / (pl.smem_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    pl.grid = sm_count() * per_sm;
    pl.slab_bytes = pl.a_global ? align_up(a_bytes, 256) : 0;
    return pl;
}

size_t simt_half_step_workspace_bytes(int f) {  // 256-byte header (row counter, flags) + slabs
    SimtPlan pl = simt_plan(f);
    return 256 + pl.slab_bytes * pl.grid;
}

int simt_half_step(const HalfStepParams& in, void* ws, size_t ws_bytes, cudaStream_t st) {
    SimtPlan pl = simt_plan(in.f);
    size_t need = simt_half_step_workspace_bytes(in.f);
    if (ws == nullptr || ws_bytes < need) {
        set_error("wmf_als_half_step(simt): workspace %zu < %zu", ws_bytes, need);
        return WMF_ERR_WORKSPACE;
    }
    HalfStepParams p = in;
    p.counter = reinterpret_cast<int*>(ws);
    const bool fixup = in.run_if != nullptr || in.sched_len_dev != nullptr;  // header already initialised by the caller
    p.slab = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 256);
    p.lda = pl.lda;
    p.FP = pl.FP;
    p.KC = pl.KC;
    if (!fixup) WMF_CUDA(cudaMemsetAsync(ws, 0, 256, st));
    int grid = pl.grid;
    if ((int64_t)grid > in.rows) grid = (int)in.rows;
    if (pl.a_global) {
        WMF_CUDA(cudaFuncSetAttribute(als_half_step_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
        als_half_step_simt_kernel<true><<<grid, HS_THREADS, pl.smem_bytes, st>>>(p);
    } else {
        WMF_CUDA(cudaFuncSetAttribute(als_half_step_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem_bytes));
        als_half_step_simt_kernel<false><<<grid, HS_THREADS, pl.smem_bytes, st>>>(p);
    }
    WMF_LAUNCH_CHECK("als_half_step_simt_kernel");
    return WMF_OK;
}

}  // namespace wmf
