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

constexpr int CW_THREADS = 512;    // 16 warps: warp ty owns the rows ty, ty + 16, ty + 32, ... of the matrix, lane tx the columns tx + 32 b
constexpr int CW_NB = 8;           // panel width

// Cholesky of an 8 x 8 block (lower triangle d, packed rows) and, in place, the inverse of its factor, in double,
// straight line: every lane of the calling warp computes all of it (the chain of dependent operations is what costs).
#define T8(i, j) ((i) * ((i) + 1) / 2 + (j))
__device__ __forceinline__ bool chol8_inv(double (&d)[36]) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double p = d[T8(k, k)];
        if (!(p > 0.0)) { ok = false; p = 1.0; }
        const double r = rsqrt(p);
        d[T8(k, k)] = r;   // 1 / L_kk (the inversion below wants the reciprocal)
#pragma unroll
        for (int i = k + 1; i < 8; ++i) d[T8(i, k)] *= r;
#pragma unroll
        for (int j = k + 1; j < 8; ++j)
#pragma unroll
            for (int i = j; i < 8; ++i) d[T8(i, j)] = fma(-d[T8(i, k)], d[T8(j, k)], d[T8(i, j)]);
    }
    // in-place inverse of the lower-triangular factor, last column first (LAPACK trti2): the trailing block is
    // already inverted when column j is formed: x <- -(1 / L_jj) T x
#pragma unroll
    for (int j = 6; j >= 0; --j) {
#pragma unroll
        for (int i = 7; i > j; --i) {   // descending i: x_k (k < i) is still the old column
            double acc = 0.0;
#pragma unroll
            for (int k = j + 1; k < i; ++k) acc = fma(d[T8(i, k)], d[T8(k, j)], acc);
            acc = fma(d[T8(i, i)], d[T8(i, j)], acc);   // T_ii = 1 / L_ii
            d[T8(i, j)] = -acc * d[T8(j, j)];
        }
    }
    return ok;
}

// L = chol(G) and Z = L^-1 in ONE right-looking sweep over 8-column panels, in double, in a single lower-triangular
// array M (columns left of the current panel already hold Z, columns from the panel on hold the trailing matrix):
//   N  = inverse of the Cholesky factor of the 8 x 8 pivot block
//   l_i = M[i][panel] N^T for the rows below (the panel of L, needed by this step only)
//   pivot rows:  Z[p][j] <- N Z[p][j] (j left of the panel),  Z[p][panel] = N
//   rows below:  M[i][j] -= l_i . R[:, j]  for j <= i, with R = [N Z[p] | N | l^T]  (Z part, new Z columns, trailing matrix)
// Every step is a rank-8 update of the whole triangle: no separate triangular inversion, f / 8 steps of three
// barriers each. One CTA; the work is ~f^3 / 2 double FMAs (nothing), the cost is the chain of f / 8 panel steps.
__global__ void __launch_bounds__(CW_THREADS, 1)
chol_whiten_kernel(const float* __restrict__ G, int f, int FP, double* __restrict__ gscratch, int use_smem,
                   float* __restrict__ Mw, float* __restrict__ Mu, float* __restrict__ eye, int* __restrict__ flags) {
    extern __shared__ double cw_smem[];
    __shared__ double Nsh[36];
    __shared__ int s_bad;
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
    const int FPAD = (f + CW_NB - 1) / CW_NB * CW_NB;
    const int ld = FPAD | 1;  // odd leading dimension: row-strided accesses spread over the banks
    double* M = use_smem ? cw_smem : gscratch;
    double* R = M + (size_t)FPAD * ld;          // [8][FPAD]
    if (tid == 0) s_bad = 0;
    for (int e = tid; e < FPAD * FPAD; e += CW_THREADS) {
        const int i = e / FPAD, j = e % FPAD;
        if (j <= i) M[(size_t)i * ld + j] = (i < f) ? (double)G[(size_t)i * f + j] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int k0 = 0; k0 < FPAD; k0 += CW_NB) {
        // ---- A: pivot block (warp 0, every lane the same straight-line code)
        if (ty == 0) {
            double d[36];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) d[T8(i, j)] = M[(size_t)(k0 + i) * ld + k0 + j];
            const bool ok = chol8_inv(d);
            if (tx == 0) {
#pragma unroll
                for (int q = 0; q < 36; ++q) Nsh[q] = d[q];
                if (!ok) s_bad = 1;
            }
        }
        __syncthreads();
        // ---- B: thread j builds column j of R and rewrites the pivot rows
        if (tid < FPAD) {
            const int j = tid;
            double nn[36];
#pragma unroll
            for (int q = 0; q < 36; ++q) nn[q] = Nsh[q];
            if (j < k0) {                     // R[:, j] = N Z[pivot rows][j]
                double z[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) z[c] = M[(size_t)(k0 + c) * ld + j];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double acc = 0.0;
#pragma unroll
                    for (int c2 = 0; c2 <= c; ++c2) acc = fma(nn[T8(c, c2)], z[c2], acc);
                    R[(size_t)c * FPAD + j] = acc;
                    M[(size_t)(k0 + c) * ld + j] = acc;
                }
            } else if (j < k0 + CW_NB) {      // R[:, panel] = N (lower), which is also the pivot block of Z
                const int c2 = j - k0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double v = 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) v = (q == c2 && q <= c) ? nn[T8(c, q)] : v;
                    R[(size_t)c * FPAD + j] = v;
                    if (c2 <= c) M[(size_t)(k0 + c) * ld + j] = v;
                }
            } else {                          // R[:, j] = l_j = N M[j][panel] (the row of the L panel)
                double m[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) m[c] = M[(size_t)j * ld + k0 + c];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double acc = 0.0;
#pragma unroll
                    for (int c2 = 0; c2 <= c; ++c2) acc = fma(nn[T8(c, c2)], m[c2], acc);
                    R[(size_t)c * FPAD + j] = acc;
                }
            }
        }
        __syncthreads();
        // ---- C: rows below the panel: M[i][j] -= l_i . R[:, j] for j <= i (panel columns start from zero)
        for (int i0 = ty; i0 < FPAD; i0 += 32) {     // two rows per pass share the loads of R
            const int i1 = i0 + 16;
            const bool on0 = i0 >= k0 + CW_NB, on1 = i1 >= k0 + CW_NB && i1 < FPAD;
            if (!on0 && !on1) continue;              // warp-uniform
            double l0[8], l1[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                l0[c] = on0 ? R[(size_t)c * FPAD + i0] : 0.0;
                l1[c] = on1 ? R[(size_t)c * FPAD + i1] : 0.0;
            }
            const int jmax = on1 ? i1 : i0;
            for (int j = tx; j <= jmax; j += 32) {
                double r[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) r[c] = R[(size_t)c * FPAD + j];
                const bool panel = j >= k0 && j < k0 + CW_NB;
                if (on0 && j <= i0) {
                    double m = panel ? 0.0 : M[(size_t)i0 * ld + j];
#pragma unroll
                    for (int c = 0; c < 8; ++c) m = fma(-l0[c], r[c], m);
                    M[(size_t)i0 * ld + j] = m;
                }
                if (on1 && j <= i1) {
                    double m = panel ? 0.0 : M[(size_t)i1 * ld + j];
#pragma unroll
                    for (int c = 0; c < 8; ++c) m = fma(-l1[c], r[c], m);
                    M[(size_t)i1 * ld + j] = m;
                }
            }
        }
        __syncthreads();
    }
    // ---- multipliers, zero padded to FP x FP (zero matrices when G is not positive definite: every row
    // then goes to the LU fix-up list, see tc_prep_rows_kernel). M's lower triangle is Z = L^-1.
    const bool bad = s_bad != 0;
    if (bad && tid == 0) atomicOr(flags, 8);
    for (int e = tid; e < FP * FP; e += CW_THREADS) {
        const int k = e / FP, n = e % FP;
        float w = 0.0f, u = 0.0f;
        if (!bad && k < f && n < f) {
            if (n >= k) w = (float)M[(size_t)n * ld + k];   // Linv[n][k]
            if (k >= n) u = (float)M[(size_t)k * ld + n];   // Linv[k][n]
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

static size_t chol_doubles(int f) {   // triangular array (square storage, odd ld) + the 8-row panel R
    const size_t FPAD = (size_t)(f + CW_NB - 1) / CW_NB * CW_NB;
    return FPAD * (FPAD | 1) + CW_NB * FPAD;
}

size_t whiten_scratch_bytes(int f) {  // only touched when the matrix does not fit shared memory (f > ~160)
    return align_up(chol_doubles(f) * sizeof(double), 256);
}

int chol_whiten(const float* G, int f, int FP, void* scratch, float* Mw, float* Mu, float* eye, int* flags, cudaStream_t st) {
    const size_t need = chol_doubles(f) * sizeof(double);
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
