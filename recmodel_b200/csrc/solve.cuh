// Dense f x f solves shared by the half-step kernels (replaces np.linalg.solve = LAPACK sgesv
// at wmf_model.py:239,:350). All NT threads of the CTA call these together; A may live in
// shared memory or in an L2-resident global slab (large f).
//
//  chol_factor_aug : SPD path. A is (f+1) x lda, lower triangle used; row f carries the
//                    right-hand side, so the forward substitution falls out of the
//                    factorisation (row f is just one more row of L). One barrier per column.
//  chol_back_solve : L^T x = z by warp 0.
//  lu_solve_aug    : general path (bias formula can be indefinite, wmf_model.py:343): LU with
//                    partial pivoting like sgesv, RHS carried as column f.
#pragma once
#include "common.cuh"

namespace wmf {

// Barrier over the threads that run the solve: the whole CTA (SIMT kernel) or a named
// barrier over the solver warps only (warp-specialised tcgen05 kernel).
struct BlockSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
template <int ID, int NTHREADS>
struct NamedSync {
    __device__ __forceinline__ void operator()() const {
        asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTHREADS) : "memory");
    }
};

template <int NT, class Sync = BlockSync>
__device__ __forceinline__ bool chol_factor_aug(float* A, int lda, int f, float* dinv, int tid, Sync sync = Sync()) {
    constexpr int NY = NT / 16;
    const int ty = tid / 16, tx = tid % 16;
    for (int k = 0; k < f; ++k) {
        const float p = A[k * lda + k];
        if (!(p > 0.0f)) return false;  // uniform: every thread reads the same word
        // multipliers by true division: a reciprocal-multiply costs an extra rounding that is
        // coherent along the whole row and shows up 3x in the solution error
        for (int i = k + 1 + ty; i <= f; i += NY) {
            const float lik = __fdiv_rn(A[i * lda + k], p);
            const int jmax = i < f ? i : f - 1;
            for (int j = k + 1 + tx; j <= jmax; j += 16) A[i * lda + j] = fmaf(-lik, A[j * lda + k], A[i * lda + j]);
        }
        sync();
        const float sq = __fsqrt_rn(p);
        for (int i = k + 1 + tid; i <= f; i += NT) A[i * lda + k] = __fdiv_rn(A[i * lda + k], sq);
        if (tid == 0) dinv[k] = sq;  // holds L[k][k]
    }
    sync();
    return true;
}

// x (length f) written to xout (global). z = row f of A. Warp 0 only; others fall through.
__device__ __forceinline__ void chol_back_solve(float* A, int lda, int f, const float* dinv, float* xout, int tid) {
    if (tid >= 32) return;
    float* z = A + (size_t)f * lda;
    for (int k = f - 1; k >= 0; --k) {
        const float xk = __fdiv_rn(z[k], dinv[k]);
        const float* Lk = A + (size_t)k * lda;
        for (int j = tid; j < k; j += 32) z[j] = fmaf(-Lk[j], xk, z[j]);
        if (tid == 0) xout[k] = xk;
        __syncwarp();
    }
}

template <int NT>
__device__ __forceinline__ void lu_solve_aug(float* A, int lda, int f, int* piv_sh, float* xout, int tid) {
    constexpr int NY = NT / 16;
    const int ty = tid / 16, tx = tid % 16;
    for (int k = 0; k < f; ++k) {
        if (tid < 32) {
            float best = -1.0f;
            int bi = k;
            for (int i = k + tid; i < f; i += 32) {
                float v = fabsf(A[i * lda + k]);
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) *piv_sh = bi;
        }
        __syncthreads();
        const int p = *piv_sh;
        if (p != k) {
            for (int j = k + tid; j <= f; j += NT) {
                float t = A[k * lda + j];
                A[k * lda + j] = A[p * lda + j];
                A[p * lda + j] = t;
            }
        }
        __syncthreads();
        const float pivot = A[k * lda + k];
        for (int i = k + 1 + ty; i < f; i += NY) {
            const float m = __fdiv_rn(A[i * lda + k], pivot);
            for (int j = k + 1 + tx; j <= f; j += 16) A[i * lda + j] = fmaf(-m, A[k * lda + j], A[i * lda + j]);
        }
        __syncthreads();
    }
    if (tid < 32) {
        for (int k = f - 1; k >= 0; --k) {
            const float xk = A[k * lda + f] / A[k * lda + k];
            for (int i = tid; i < k; i += 32) A[i * lda + f] = fmaf(-A[i * lda + k], xk, A[i * lda + f]);
            if (tid == 0) xout[k] = xk;
            __syncwarp();
        }
    }
}

}  // namespace wmf
