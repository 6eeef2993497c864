// Blocked FP32 Cholesky solve of one 128x128 SPD system by a group of 128 threads
// (replaces np.linalg.solve / LAPACK sgesv at wmf_model.py:239 on the tcgen05 path).
//
// The per-row solve is a dependency chain of 128 pivots, so one matrix can never fill an SM;
// the kernel keeps several groups busy on different rows instead. Inside a group:
//   * panels of NB = 8 columns. Every thread loads the 8x8 diagonal block (broadcast) and
//     factors it redundantly in registers - no communication inside the 8-pivot chain - then
//     solves its own matrix row against it (thread t owns row t; the thread that owns row 0
//     also carries the right-hand side as row 128, which fuses the forward substitution).
//   * the panel goes to shared memory transposed (LpT[k][row]); the trailing update runs on
//     4x4 register tiles over the lower triangle with 128-bit shared-memory accesses.
//   * back substitution in blocks of 8 (block solve redundant per thread, one barrier per block).
// Two named-barrier syncs per panel, one per back-substitution block.
//
// Storage: A is ROWS x LDA floats (LDA = 132 keeps rows 16-byte aligned and rotates banks by 4
// per row); rows 0..127 lower triangle, row 128 = rhs, rows 129..131 scratch so that 4-row tiles
// never leave the buffer.
#pragma once
#include "common.cuh"

namespace wmf {
namespace s128 {

constexpr int F = 128;
constexpr int NB = 8;
constexpr int LDA = 132;
constexpr int ROWS = 132;
constexpr int A_FLOATS = ROWS * LDA;
constexpr int LPT_LD = 132;
constexpr int LPT_FLOATS = NB * LPT_LD;
constexpr int GROUP = 128;

template <int BAR_ID>
__device__ __forceinline__ void group_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(BAR_ID), "n"(GROUP) : "memory");
}

__device__ __forceinline__ void tile_from_linear(int t, int& a, int& b) {
    int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    while (r * (r + 1) / 2 > t) --r;
    a = r;
    b = t - r * (r + 1) / 2;
}

// rows handled by one thread in the panel phase
__device__ __forceinline__ void panel_row(const float* __restrict__ Arow, int c0, int rel, const float (&L)[NB][NB],
                                          const float (&rinv)[NB], float (&l)[NB]) {
    const float4 a0 = *reinterpret_cast<const float4*>(Arow + c0);
    const float4 a1 = *reinterpret_cast<const float4*>(Arow + c0 + 4);
    const float a[NB] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        float t = a[j];
#pragma unroll
        for (int k = 0; k < j; ++k) t = fmaf(-l[k], L[j][k], t);
        t *= rinv[j];
        l[j] = (j > rel) ? 0.0f : t;  // rows inside the diagonal block: nothing right of the diagonal
    }
}

// Factor + solve. `t` = this thread's matrix row (0..127, a permutation of the group's
// threads). x is left in xs[0..127] (shared). Returns false (uniformly) on a non-positive pivot.
template <int BAR_ID>
__device__ __forceinline__ bool chol_solve_128(float* __restrict__ A, float* __restrict__ LpT,
                                               float* __restrict__ dinv, float* __restrict__ xs, int t) {
    bool ok = true;
    for (int c0 = 0; c0 < F; c0 += NB) {
        // ---------------- panel ----------------
        float L[NB][NB], rinv[NB];
        {
            float D[NB][NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                const float4 d0 = *reinterpret_cast<const float4*>(A + (c0 + r) * LDA + c0);
                const float4 d1 = *reinterpret_cast<const float4*>(A + (c0 + r) * LDA + c0 + 4);
                D[r][0] = d0.x; D[r][1] = d0.y; D[r][2] = d0.z; D[r][3] = d0.w;
                D[r][4] = d1.x; D[r][5] = d1.y; D[r][6] = d1.z; D[r][7] = d1.w;
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                float s = D[j][j];
#pragma unroll
                for (int k = 0; k < j; ++k) s = fmaf(-L[j][k], L[j][k], s);
                ok = ok && (s > 0.0f);
                float r = rsqrtf(s);
                r = r * fmaf(-0.5f * s, r * r, 1.5f);  // one Newton step: ~1 ulp
                rinv[j] = r;
                L[j][j] = s * r;
#pragma unroll
                for (int i = j + 1; i < NB; ++i) {
                    float v = D[i][j];
#pragma unroll
                    for (int k = 0; k < j; ++k) v = fmaf(-L[i][k], L[j][k], v);
                    L[i][j] = v * r;
                }
            }
        }
        float l_own[NB], l_rhs[NB];
        const bool own = t >= c0;
        if (own) {
            panel_row(A + t * LDA, c0, t - c0, L, rinv, l_own);
#pragma unroll
            for (int j = 0; j < NB; ++j) LpT[j * LPT_LD + t] = l_own[j];
        }
        if (t == 0) {
            panel_row(A + F * LDA, c0, 1 << 20, L, rinv, l_rhs);
#pragma unroll
            for (int j = 0; j < NB; ++j) LpT[j * LPT_LD + F] = l_rhs[j];
        }
        if (t == c0) {
#pragma unroll
            for (int j = 0; j < NB; ++j) dinv[c0 + j] = rinv[j];
        }
        group_sync<BAR_ID>();
        // ---------------- write the finished panel back, trailing update ----------------
        if (own) {
            *reinterpret_cast<float4*>(A + t * LDA + c0) = make_float4(l_own[0], l_own[1], l_own[2], l_own[3]);
            *reinterpret_cast<float4*>(A + t * LDA + c0 + 4) = make_float4(l_own[4], l_own[5], l_own[6], l_own[7]);
        }
        if (t == 0) {
            *reinterpret_cast<float4*>(A + F * LDA + c0) = make_float4(l_rhs[0], l_rhs[1], l_rhs[2], l_rhs[3]);
            *reinterpret_cast<float4*>(A + F * LDA + c0 + 4) = make_float4(l_rhs[4], l_rhs[5], l_rhs[6], l_rhs[7]);
        }
        const int t0 = (c0 + NB) / 4;      // first tile row/column of the trailing matrix
        const int n = 33 - t0;             // tile rows t0..32 (tile row 32 = the rhs row)
        const int ntile = n * (n + 1) / 2 - 1;  // lower triangle without the (32,32) tile
        for (int tl = t; tl < ntile; tl += GROUP) {
            int ta, tb;
            tile_from_linear(tl, ta, tb);
            const int ti = t0 + ta, tj = t0 + tb;
            float acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const float4 li = *reinterpret_cast<const float4*>(LpT + k * LPT_LD + 4 * ti);
                const float4 lj = *reinterpret_cast<const float4*>(LpT + k * LPT_LD + 4 * tj);
                const float iv[4] = {li.x, li.y, li.z, li.w};
                const float jv[4] = {lj.x, lj.y, lj.z, lj.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(iv[a], jv[b], acc[a][b]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float4* dst = reinterpret_cast<float4*>(A + (4 * ti + a) * LDA + 4 * tj);
                float4 v = *dst;
                v.x -= acc[a][0]; v.y -= acc[a][1]; v.z -= acc[a][2]; v.w -= acc[a][3];
                *dst = v;
            }
        }
        group_sync<BAR_ID>();
    }
    if (!ok) return false;
    // ---------------- back substitution  L^T x = z  (z = row 128) ----------------
    float* z = A + F * LDA;
    for (int c0 = F - NB; c0 >= 0; c0 -= NB) {
        float x[NB];
        {
            const float4 z0 = *reinterpret_cast<const float4*>(z + c0);
            const float4 z1 = *reinterpret_cast<const float4*>(z + c0 + 4);
            const float4 r0 = *reinterpret_cast<const float4*>(dinv + c0);
            const float4 r1 = *reinterpret_cast<const float4*>(dinv + c0 + 4);
            const float zb[NB] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            const float ri[NB] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            float Lb[NB][NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                const float4 d0 = *reinterpret_cast<const float4*>(A + (c0 + r) * LDA + c0);
                const float4 d1 = *reinterpret_cast<const float4*>(A + (c0 + r) * LDA + c0 + 4);
                Lb[r][0] = d0.x; Lb[r][1] = d0.y; Lb[r][2] = d0.z; Lb[r][3] = d0.w;
                Lb[r][4] = d1.x; Lb[r][5] = d1.y; Lb[r][6] = d1.z; Lb[r][7] = d1.w;
            }
#pragma unroll
            for (int j = NB - 1; j >= 0; --j) {
                float s = zb[j];
#pragma unroll
                for (int r = j + 1; r < NB; ++r) s = fmaf(-Lb[r][j], x[r], s);
                x[j] = s * ri[j];
            }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j)
            if (t == c0 + j) xs[t] = x[j];
        if (t < c0) {
            float s = z[t];
#pragma unroll
            for (int k = 0; k < NB; ++k) s = fmaf(-A[(c0 + k) * LDA + t], x[k], s);
            z[t] = s;
        }
        group_sync<BAR_ID>();
    }
    return true;
}

}  // namespace s128
}  // namespace wmf
