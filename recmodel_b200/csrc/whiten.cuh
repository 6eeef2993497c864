// Whitening helpers of the tcgen05 half-step (whiten.cu).
#pragma once
#include "common.cuh"

namespace wmf {

// bytes of double scratch chol_whiten needs for factor width f (only touched when f*f doubles exceed shared memory)
size_t whiten_scratch_bytes(int f);
// L = chol(G) in double; Mw[k][n] = Linv[n][k], Mu[k][n] = Linv[k][n], both fp32 FP x FP, zero padded.
// G not positive definite: bit 3 of *flags is set and both multipliers are zero matrices.
// eye: f x f identity (ld = f), the Gram of the whitened factors, for the CUDA-core fix-up kernel.
int chol_whiten(const float* G, int f, int FP, void* scratch, float* Mw, float* Mu, float* eye, int* flags, cudaStream_t st);
// out[r][0..nout) = sum_{k<kin} in[r][k] M[k][.]; M is FP x FP (FP a multiple of 128, kin <= FP, nout <= FP);
// ones_col0: column 0 of `in` reads as 1; maxsq (nullable): atomic maximum of out^2 (float bits).
int right_multiply(const float* in, int64_t rows, int64_t ldin, int kin, int ones_col0, const float* M, int FP,
                   float* out, int64_t ldout, int nout, unsigned int* maxsq, cudaStream_t st);

}  // namespace wmf
