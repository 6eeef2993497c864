// Conjugate-gradient solve of a whitened ALS system whose matrix is resident in tensor memory.
// Shared by the three tcgen05 half-step kernels (half_step_tc.cu, half_step_dual.cu, half_step_tc256.cu);
// replaces np.linalg.solve of wmf_model.py:239 / :350 for the rows these kernels take.
//
// Why an iteration and not a factorisation: the whitened system A = I + sum_j d_j y~_j y~_j^T (or I + W W^T in the
// dual form) has its spectrum in [1, 1 + max_j d_j] (whiten.cu): condition numbers of 4 ... 20 on
// the reference's weightings, so CG reaches fp32 round-off in 9 ... 20 matrix-vector products, each of which is
// f FMAs per thread against the thread's own matrix row (TMEM lane = matrix row) and ONE block-wide reduction.
// The block Gauss-Jordan it replaces as the default needs f/8 steps of ~4000 cycles each (a serial 8 x 8 pivot
// factor, three barriers and a tensor-core round trip per step) with only four systems in flight per SM.
// The factorisation stays in the kernels as the fallback: a system that has not converged after a bounded number of products
// (weights in the thousands) is solved by it from the untouched matrix.
//
// Variant: Chronopoulos/Gear (one reduction per iteration: gamma = r.r and delta = r.Ar together):
//     w = A r;  beta = gamma / gamma_old;  alpha = gamma / (delta - beta gamma / alpha_old)
//     p = r + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s
// Every scalar is computed by every thread from the same shared-memory partials in the same order, so control
// flow is uniform and a row's arithmetic depends on nothing but the row (row-sharded runs stay bitwise equal).
// Cost per product: the matrix row is re-read from tensor memory (64 B/cycle per SM sub-partition, measured by
// scripts/probe/tmem_ld_probe.cu: a 128 x 128 product occupies the four sub-partitions for 256 cycles).
#pragma once
#include "tc_common.cuh"

namespace wmf {
namespace tc {

// relative residual (2-norm) at which the iteration stops; the error of x is then <= cond(A) * CG_TOL
constexpr float CG_TOL = 1.0e-6f;
constexpr float CG_TOL2 = CG_TOL * CG_TOL;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float2 lds2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts2(uint32_t a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}

// sum_c T[t][c] v[c] over the W live columns (W a multiple of 16); four accumulation chains in a fixed order
__device__ __forceinline__ float cg_row_dot(uint32_t t_row, uint32_t vec, int W) {
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    int c = 0;
#pragma unroll 1
    for (; c + 32 <= W; c += 32) {
        const uint32_t va = vec + (uint32_t)c * 4u;
        const float4 v0 = lds4(va), v1 = lds4(va + 16), v2 = lds4(va + 32), v3 = lds4(va + 48);
        float m[32];
        tmem_ld32(t_row + (uint32_t)c, m);
        const float4 v4 = lds4(va + 64), v5 = lds4(va + 80), v6 = lds4(va + 96), v7 = lds4(va + 112);
        a0 = fmaf(m[0], v0.x, a0); a1 = fmaf(m[1], v0.y, a1); a2 = fmaf(m[2], v0.z, a2); a3 = fmaf(m[3], v0.w, a3);
        a0 = fmaf(m[4], v1.x, a0); a1 = fmaf(m[5], v1.y, a1); a2 = fmaf(m[6], v1.z, a2); a3 = fmaf(m[7], v1.w, a3);
        a0 = fmaf(m[8], v2.x, a0); a1 = fmaf(m[9], v2.y, a1); a2 = fmaf(m[10], v2.z, a2); a3 = fmaf(m[11], v2.w, a3);
        a0 = fmaf(m[12], v3.x, a0); a1 = fmaf(m[13], v3.y, a1); a2 = fmaf(m[14], v3.z, a2); a3 = fmaf(m[15], v3.w, a3);
        a0 = fmaf(m[16], v4.x, a0); a1 = fmaf(m[17], v4.y, a1); a2 = fmaf(m[18], v4.z, a2); a3 = fmaf(m[19], v4.w, a3);
        a0 = fmaf(m[20], v5.x, a0); a1 = fmaf(m[21], v5.y, a1); a2 = fmaf(m[22], v5.z, a2); a3 = fmaf(m[23], v5.w, a3);
        a0 = fmaf(m[24], v6.x, a0); a1 = fmaf(m[25], v6.y, a1); a2 = fmaf(m[26], v6.z, a2); a3 = fmaf(m[27], v6.w, a3);
        a0 = fmaf(m[28], v7.x, a0); a1 = fmaf(m[29], v7.y, a1); a2 = fmaf(m[30], v7.z, a2); a3 = fmaf(m[31], v7.w, a3);
    }
    if (c < W) {
        const uint32_t va = vec + (uint32_t)c * 4u;
        const float4 v0 = lds4(va), v1 = lds4(va + 16), v2 = lds4(va + 32), v3 = lds4(va + 48);
        float m[16];
        tmem_ld16(t_row + (uint32_t)c, m);
        a0 = fmaf(m[0], v0.x, a0); a1 = fmaf(m[1], v0.y, a1); a2 = fmaf(m[2], v0.z, a2); a3 = fmaf(m[3], v0.w, a3);
        a0 = fmaf(m[4], v1.x, a0); a1 = fmaf(m[5], v1.y, a1); a2 = fmaf(m[6], v1.z, a2); a3 = fmaf(m[7], v1.w, a3);
        a0 = fmaf(m[8], v2.x, a0); a1 = fmaf(m[9], v2.y, a1); a2 = fmaf(m[10], v2.z, a2); a3 = fmaf(m[11], v2.w, a3);
        a0 = fmaf(m[12], v3.x, a0); a1 = fmaf(m[13], v3.y, a1); a2 = fmaf(m[14], v3.z, a2); a3 = fmaf(m[15], v3.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// Solves (I + inv_s2 T) x = b for the W x W matrix T in tensor memory. Called by the `nw` warps whose lanes hold
// matrix rows (warp index `wi` among them, `nthr` = 32 nw threads on named barrier `bar`); thread `t` owns row t
// (rows >= W of a partly filled warp are ignored). Returns the number of products spent when the residual has dropped
// below CG_TOL |b| within `maxit` of them (x is then the solution), -1 otherwise (the matrix is untouched: factorise it).
//   vec   shared, W floats, 16-byte aligned: the vector being multiplied
//   red   shared, 2 x nw x 2 floats, 8-byte aligned: warp partials of (r.r, r.Ar), double buffered by iteration
__device__ __forceinline__ int cg_solve(uint32_t t_row, int t, int W, float b, float inv_s2, uint32_t vec, uint32_t red,
                                        int wi, int nw, int bar, int nthr, int maxit, float& x_out,
                                        long long* pf = nullptr) {   // pf: cycle counters of a profiling build
    const bool live = t < W;
    const int lane = threadIdx.x & 31;
    const bool solo = nw == 1;   // a single warp holds every row: no block barrier, no exchange through shared memory
    float x = 0.0f, r = live ? b : 0.0f, pv = 0.0f, sv = 0.0f;
    float gamma_old = 1.0f, alpha_old = 1.0f, gamma0 = 0.0f;
    int products = -1;
#pragma unroll 1
    for (int it = 0;; ++it) {
#ifdef WMF_TC_PROFILE_BUILD
        long long c0 = pf ? clock64() : 0, c1 = 0, c2 = 0, c3 = 0;
#endif
        if (live) sts1(vec + (uint32_t)t * 4u, r);
        if (solo) __syncwarp(); else named_bar(bar, nthr);
#ifdef WMF_TC_PROFILE_BUILD
        if (pf) c1 = clock64();
#endif
        float w = fmaf(cg_row_dot(t_row, vec, W), inv_s2, r);   // (A r)_t
        w = live ? w : 0.0f;
#ifdef WMF_TC_PROFILE_BUILD
        if (pf) { pf[3] += (long long)(__float_as_int(w) & 0); c2 = clock64(); }   // (the dependency keeps the clock behind the product)
#endif
        float g = r * r, d = r * w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // (every lane ends up with the warp's sums; the shuffles also order this
            g += __shfl_xor_sync(0xffffffffu, g, o);   //  step's reads of `vec` before the next step's writes)
            d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        float gamma = g, delta = d;
        if (!solo) {
            const uint32_t rb = red + (uint32_t)((it & 1) * nw) * 8u;
            if (lane == 0) sts2(rb + (uint32_t)wi * 8u, g, d);
            named_bar(bar, nthr);
            gamma = 0.0f;
            delta = 0.0f;
            for (int k = 0; k < nw; ++k) {
                const float2 q = lds2(rb + (uint32_t)k * 8u);
                gamma += q.x;
                delta += q.y;
            }
        }
#ifdef WMF_TC_PROFILE_BUILD
        if (pf) { c3 = clock64() + (long long)(__float_as_int(gamma) & 0); pf[0] += c1 - c0; pf[1] += c2 - c1; pf[2] += c3 - c2; }
#endif
        if (it == 0) gamma0 = gamma;
        if (gamma <= CG_TOL2 * gamma0) { products = it + 1; break; }   // also a zero right-hand side
        if (it >= maxit) break;
        const float beta = it == 0 ? 0.0f : __fdividef(gamma, gamma_old);
        const float den = it == 0 ? delta : delta - beta * __fdividef(gamma, alpha_old);
        if (!(den > 0.0f)) break;   // break-down (cannot happen for a positive definite matrix in exact arithmetic)
        const float alpha = __fdividef(gamma, den);
        pv = fmaf(beta, pv, r);
        sv = fmaf(beta, sv, w);
        x = fmaf(alpha, pv, x);
        r = fmaf(-alpha, sv, r);
        gamma_old = gamma;
        alpha_old = alpha;
    }
    x_out = x;
    return products;
}

}  // namespace tc
}  // namespace wmf
