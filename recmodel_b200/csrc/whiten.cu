// Whitening of the fixed-side factors for the tcgen05 half-step (half_step_tc.cu, half_step_dual.cu).
//
// The reference solves (G + sum_j d_j y_j y_j^T) x = sum_j (d_j+1) y_j per row (wmf_model.py:237-239) with
// G = Y^T Y + lambda I (:215). With G = L L^T and y~_j = L^-1 y_j the same system reads
//     (I + sum_j d_j y~_j y~_j^T) x' = sum_j (d_j+1) y~_j ,   x = L^-T x'.
// Every y_j is a row of the Y that formed G, so y~_j^T y~_j = y_j^T G^-1 y_j <= 1 and the matrix above has its
// spectrum in [1, 1 + sum d_j y~_j^T y~_j]: condition numbers of ~10 where G + ... has 10^2..10^5 on the
// all-positive factors of the first epochs. Measured (numpy emulation of the tensor-core arithmetic, fp64 as
// truth): row-relative error 2e-7..1.5e-6 for this form against 3e-5..7.5e-5 for the unwhitened one.
//
// Kernels here:
//   chol_whiten_kernel  one CTA: L = chol(G) and L^-1 in double; writes the two fp32 multiplier matrices
//                       Mw[k][n] = Linv[n][k] (whiten:   Y~ = Y  Mw) and Mu[k][n] = Linv[k][n] (unwhiten: X = X' Mu).
//   rmul_kernel         out = in * M for a tall `in` (rows x K) and a small square M: FP32 register-blocked GEMM,
//                       optional ones column (the bias formula's Y[:,0] = 1, wmf_model.py:331), zero padding of
//                       the output to FP columns, running maximum of out^2 (fixes the FP16 scale of the Gram).
#include "common.cuh"
#include "whiten.cuh"

namespace wmf {

namespace {

constexpr int CW_THREADS = 256;   // one thread per matrix row / column (f <= 256)

// A (double, ld) holds L in its lower triangle (diagonal included) and, after the second phase, column j of
// L^-1 below the diagonal in ROW j of the strict upper triangle; 1 / L_jj in dinv.
// The work is ~f^3/2 double FMAs (nothing); the cost is the chain of f dependent columns, so every phase is
// written for latency: thread i owns row i (Cholesky, two barriers per column, four accumulators per dot product),
// thread j owns column j of L^-1 (forward substitution, no barrier at all: it only reads L and its own column).
__global__ void __launch_bounds__(CW_THREADS, 1)
chol_whiten_kernel(const float* __restrict__ G, int f, int FP, double* __restrict__ gscratch, int use_smem,
                   float* __restrict__ Mw, float* __restrict__ Mu, float* __restrict__ eye, int* __restrict__ flags) {
    extern __shared__ double cw_smem[];
    __shared__ double s_rpiv;
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) s_bad = 0;
    const int ld = f | 1;  // odd leading dimension: row-strided accesses spread over the banks
    double* A = use_smem ? cw_smem : gscratch;
    double* dinv = A + (size_t)f * ld;
    for (int e = tid; e < f * f; e += CW_THREADS) {
        const int i = e / f, j = e % f;
        if (j <= i) A[i * ld + j] = (double)G[(size_t)i * f + j];
    }
    __syncthreads();
    // ---- left-looking Cholesky: column k from the k columns before it
    const int i = tid;
    const double* ai = A + (size_t)(i < f ? i : 0) * ld;
    for (int k = 0; k < f; ++k) {
        double s = 0.0;
        if (i >= k && i < f) {
            const double* ak = A + (size_t)k * ld;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int j = 0;
            for (; j + 4 <= k; j += 4) {
                a0 = fma(ai[j], ak[j], a0);
                a1 = fma(ai[j + 1], ak[j + 1], a1);
                a2 = fma(ai[j + 2], ak[j + 2], a2);
                a3 = fma(ai[j + 3], ak[j + 3], a3);
            }
            for (; j < k; ++j) a0 = fma(ai[j], ak[j], a0);
            s = ai[k] - ((a0 + a1) + (a2 + a3));
            if (i == k) {
                if (!(s > 0.0)) { s_bad = 1; s = 1.0; }  // G is not positive definite
                const double r = rsqrt(s);
                s_rpiv = r;           // L_kk = s * r = sqrt(s) falls out of the common store below
                dinv[k] = r;          // 1 / L_kk
            }
        }
        __syncthreads();
        if (i >= k && i < f) A[(size_t)i * ld + k] = s * s_rpiv;
        __syncthreads();
    }
    // ---- L^-1 column by column (forward substitution): z_j = 1/L_jj, z_i = -(sum_{j<=k<i} L_ik z_k) / L_ii
    if (tid < f) {
        const int j = tid;
        double* zrow = A + (size_t)j * ld;   // z_i (i > j) lives at A[j][i]
        const double zj = dinv[j];
        for (int r = j + 1; r < f; ++r) {
            const double* ar = A + (size_t)r * ld;
            double a0 = ar[j] * zj, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int k = j + 1;
            for (; k + 4 <= r; k += 4) {
                a0 = fma(ar[k], zrow[k], a0);
                a1 = fma(ar[k + 1], zrow[k + 1], a1);
                a2 = fma(ar[k + 2], zrow[k + 2], a2);
                a3 = fma(ar[k + 3], zrow[k + 3], a3);
            }
            for (; k < r; ++k) a0 = fma(ar[k], zrow[k], a0);
            zrow[r] = -((a0 + a1) + (a2 + a3)) * dinv[r];
        }
    }
    __syncthreads();
    // ---- multipliers, zero padded to FP x FP (zero matrices when G is not positive definite: every row
    // then goes to the LU fix-up list, see tc_prep_rows_kernel)
    const bool bad = s_bad != 0;
    if (bad && tid == 0) atomicOr(flags, 8);
    for (int e = tid; e < FP * FP; e += CW_THREADS) {
        const int k = e / FP, n = e % FP;
        float w = 0.0f, u = 0.0f;
        if (!bad && k < f && n < f) {
            if (n > k) w = (float)A[(size_t)k * ld + n];        // Linv[n][k], n > k
            else if (n == k) w = u = (float)dinv[k];
            else u = (float)A[(size_t)n * ld + k];              // Linv[k][n], k > n
        }
        Mw[e] = w;
        Mu[e] = u;
    }
    for (int e = tid; e < f * f; e += CW_THREADS) eye[e] = (e / f == e % f) ? 1.0f : 0.0f;  // the whitened G (ld = f)
}

// ---------------------------------------------------------------------------------------------------
// out[r][0..nout) = sum_k in[r][k] * M[k][.]   (k < kin; columns of `in` beyond kin read as 0)
// ---------------------------------------------------------------------------------------------------
constexpr int RM_BM = 128, RM_BN = 128, RM_BK = 16, RM_THREADS = 256;

__global__ void __launch_bounds__(RM_THREADS)
rmul_kernel(const float* __restrict__ in, int64_t rows, int64_t ldin, int kin, int ones_col0,
            const float* __restrict__ M, int FP, float* __restrict__ out, int64_t ldout, int nout,
            unsigned int* __restrict__ maxsq) {
    __shared__ __align__(16) float As[2][RM_BK][RM_BM + 4];
    __shared__ __align__(16) float Bs[2][RM_BK][RM_BN];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * RM_BM;
    const int n0 = blockIdx.y * RM_BN;
    const int ty = tid >> 4, tx = tid & 15;
    const int a_row = tid >> 1, a_k = (tid & 1) * 8;    // global -> smem roles
    const int b_k = tid >> 4, b_n = (tid & 15) * 8;
    const bool in_vec = (ldin & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    const int ksteps = (kin + RM_BK - 1) / RM_BK;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    float ra[8], rb[8];
    auto load_regs = [&](int ks) {
        const int k0 = ks * RM_BK;
        const int64_t r = r0 + a_row;
        const float* src = in + r * ldin + k0 + a_k;
        if (r < rows && in_vec && k0 + a_k + 8 <= kin) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
            ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w; ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) ra[i] = (r < rows && k0 + a_k + i < kin) ? __ldg(src + i) : 0.0f;
        }
        if (ones_col0 && k0 + a_k == 0 && r < rows) ra[0] = 1.0f;
        const float* bsrc = M + (size_t)(k0 + b_k) * FP + n0 + b_n;   // k0 + b_k < FP: M is FP x FP and kin <= FP
        const bool bok = k0 + b_k < FP;
        const float4 w0 = bok ? __ldg(reinterpret_cast<const float4*>(bsrc)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 w1 = bok ? __ldg(reinterpret_cast<const float4*>(bsrc) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        rb[0] = w0.x; rb[1] = w0.y; rb[2] = w0.z; rb[3] = w0.w; rb[4] = w1.x; rb[5] = w1.y; rb[6] = w1.z; rb[7] = w1.w;
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][a_k + i][a_row] = ra[i];
        *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
        *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n + 4]) = make_float4(rb[4], rb[5], rb[6], rb[7]);
    };
    if (ksteps > 0) { load_regs(0); store_smem(0); }
    __syncthreads();
    for (int ks = 0; ks < ksteps; ++ks) {
        const int buf = ks & 1;
        if (ks + 1 < ksteps) load_regs(ks + 1);
#pragma unroll
        for (int kk = 0; kk < RM_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 8 + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (ks + 1 < ksteps) store_smem(buf ^ 1);
        __syncthreads();
    }
    float mx = 0.0f;
    const bool out_vec = (ldout & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = r0 + ty * 8 + i;
        if (r >= rows) continue;
        const int c = n0 + tx * 8;
        float* dst = out + r * ldout + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) mx = fmaxf(mx, acc[i][j] * acc[i][j]);
        if (out_vec && c + 8 <= nout) {
            *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (c + j < nout) dst[j] = acc[i][j];
        }
    }
    if (maxsq != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((tid & 31) == 0 && mx > 0.0f) atomicMax(maxsq, __float_as_uint(mx));  // non-negative floats order like their bits
    }
}

}  // namespace

size_t whiten_scratch_bytes(int f) {  // double f x (f|1) + f, used when the matrix does not fit shared memory
    return align_up(((size_t)f * (f | 1) + f) * sizeof(double), 256);
}

int chol_whiten(const float* G, int f, int FP, void* scratch, float* Mw, float* Mu, float* eye, int* flags, cudaStream_t st) {
    const size_t need = ((size_t)f * (f | 1) + f) * sizeof(double);
    const int use_smem = need <= 200 * 1024 ? 1 : 0;
    if (use_smem)
        WMF_CUDA(cudaFuncSetAttribute(chol_whiten_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    chol_whiten_kernel<<<1, CW_THREADS, use_smem ? need : 0, st>>>(G, f, FP, reinterpret_cast<double*>(scratch), use_smem,
                                                                    Mw, Mu, eye, flags);
    WMF_LAUNCH_CHECK("chol_whiten_kernel");
    return WMF_OK;
}

int right_multiply(const float* in, int64_t rows, int64_t ldin, int kin, int ones_col0, const float* M, int FP,
                   float* out, int64_t ldout, int nout, unsigned int* maxsq, cudaStream_t st) {
    if (rows <= 0) return WMF_OK;
    dim3 grid((unsigned)((rows + RM_BM - 1) / RM_BM), (unsigned)((nout + RM_BN - 1) / RM_BN));
    rmul_kernel<<<grid, RM_THREADS, 0, st>>>(in, rows, ldin, kin, ones_col0, M, FP, out, ldout, nout, maxsq);
    WMF_LAUNCH_CHECK("rmul_kernel");
    return WMF_OK;
}

}  // namespace wmf
