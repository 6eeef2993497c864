// K2 (tcgen05 version): one ALS half-step for dim = 128, weighted, no bias.
// Replaces the per-row loop of recompute_factors (wmf_model.py:220-239).
//
// Per CSR row r:  A_r = G + sum_j d_j y_j y_j^T  (128x128, K = n_r),  b_r = sum_j (d_j+1) y_j,
// x_r = A_r^-1 b_r. The K-loop runs on the 5th-gen tensor cores with FP32-equivalent accuracy:
//
//   z_j = sqrt(d_j) y_j,  z = zh + zl  (zh, zl exactly representable in TF32, zl = z - zh)
//   sum_j z_j z_j^T ~= P + Q + Q^T,   P = sum zh zh^T,  Q = sum zh zl^T         (3xTF32 with
//   the symmetric pair folded: the dropped term zl zl^T is 2^-22 relative)
//
// so ONE tcgen05.mma (kind::tf32, M=128, N=256, K=8) per 8 stored entries produces [P | Q] in
// a 256-column fp32 TMEM accumulator: A-operand = the zh tile, B-operand = [zh ; zl] tiles.
// Plain single-pass TF32 fails the 1e-4 parity bar; this split passes it.
//
// One persistent CTA per SM, warp-specialised, all hand-offs through mbarriers:
//   warps 0-3  gather  thread m owns feature m: per 32-entry chunk it loads Y[idx_j][m] (one
//              coalesced 128-B line per warp and entry), forms zh/zl and stores them K-major
//              into a 128B-swizzled shared-memory tile (what a TMA load would have produced;
//              the operand is a gather + scale + split, which TMA cannot do), and keeps the
//              rhs b[m] in registers.
//   warp 4     MMA     one thread issues tcgen05.mma over the staged tiles, commits to the
//              stage's "empty" barrier and, after the last chunk, to the accumulator's "full".
//   warps 5-12 solve   tcgen05.ld the accumulator, form the lower triangle of
//              A = (P + Q + Q^T) + G in shared memory, release the accumulator, then Cholesky
//              (forward substitution fused as an extra row) + back substitution, write x_r.
// The accumulator is double-buffered (2 x 256 = all 512 TMEM columns) so the tensor core
// works on row r+1 while row r is being solved.
//
// Rows whose weights are negative (sqrt undefined) or whose Cholesky meets a non-positive
// pivot raise a flag; the caller then re-runs the half-step with the SIMT kernel (LU).
#include <stdlib.h>
#include "common.cuh"
#include "half_step.cuh"
#include "solve.cuh"
#include "solve128.cuh"

namespace wmf {

namespace tc {

constexpr int F = 128;               // factor width handled by this kernel
constexpr int CHUNK = 32;            // stored entries per staged tile (= one 128-byte swizzle row)
constexpr int NSTAGE = 2;
constexpr int TILE_BYTES = F * 128;  // 128 rows (features) x 32 fp32 (K) = 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;  // [zh ; zl]
constexpr int NGROUP = 2;            // solver groups, each 128 threads on its own row
constexpr int GATHER_THREADS = 128, MMA_WARP = 4, SOLVER_THREADS = NGROUP * s128::GROUP;
constexpr int THREADS = GATHER_THREADS + 32 + SOLVER_THREADS;  // 416
constexpr uint32_t TMEM_COLS = 512;

// shared memory carve-up (bytes from a 1024-aligned base)
constexpr int OFF_STAGES = 0;
constexpr int OFF_A = OFF_STAGES + NSTAGE * STAGE_BYTES;            // NGROUP x (132 x 132 floats)
constexpr int A_BYTES = s128::A_FLOATS * 4;
constexpr int OFF_LPT = OFF_A + NGROUP * A_BYTES;                   // NGROUP x (8 x 132 floats)
constexpr int LPT_BYTES = s128::LPT_FLOATS * 4;
constexpr int OFF_DINV = OFF_LPT + NGROUP * LPT_BYTES;              // NGROUP x F floats
constexpr int OFF_XS = OFF_DINV + NGROUP * F * 4;                   // NGROUP x F floats
constexpr int OFF_BVEC = OFF_XS + NGROUP * F * 4;                   // 2 x F floats
constexpr int OFF_BARS = OFF_BVEC + 2 * F * 4;                      // mbarriers (8 B each)
constexpr int NBARS = 2 * NSTAGE + 8;
constexpr int OFF_TMEM_PTR = OFF_BARS + NBARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;                // + slack for alignment
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(OFF_A % 16 == 0 && OFF_LPT % 16 == 0 && OFF_DINV % 16 == 0 && OFF_BARS % 8 == 0, "alignment");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, 128-byte swizzle, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256 (cute::UMMA::InstrDescriptor)
constexpr uint32_t IDESC_TF32_M128_N256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// round-to-nearest (ties away) to the 10-bit TF32 mantissa, done with integer ops
__device__ __forceinline__ float tf32_round(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

}  // namespace tc

using namespace tc;

__global__ void __launch_bounds__(THREADS, 1) als_half_step_tc_kernel(HalfStepParams p, int* __restrict__ flags) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    float* bvec = reinterpret_cast<float*>(smem + OFF_BVEC);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_TMEM_PTR);
    const uint32_t bars = smem_base + OFF_BARS;
    // barrier ids
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 8u * (NSTAGE + s); };
    auto bar_acc_full = [&](int b) { return bars + 8u * (2 * NSTAGE + b); };
    auto bar_acc_empty = [&](int b) { return bars + 8u * (2 * NSTAGE + 2 + b); };
    auto bar_b_full = [&](int b) { return bars + 8u * (2 * NSTAGE + 4 + b); };
    auto bar_b_empty = [&](int b) { return bars + 8u * (2 * NSTAGE + 6 + b); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full(s), GATHER_THREADS); mbar_init(bar_empty(s), 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full(b), 1);
            mbar_init(bar_acc_empty(b), s128::GROUP);
            mbar_init(bar_b_full(b), GATHER_THREADS);
            mbar_init(bar_b_empty(b), s128::GROUP);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int64_t rows = p.rows;
    const int64_t first = blockIdx.x, step = gridDim.x;

    if (warp < 4) {
        // =============================== GATHER ===============================
        // Software pipeline over the CTA's flattened (row, chunk) sequence: while chunk c is
        // scaled/split/stored, the 32 factor loads of chunk c+1 and the index/weight loads of
        // chunk c+2 are already in flight.
        const int m = tid;  // feature index
        struct Cursor {
            int64_t r, base, hi;  // position in the CTA's row list, first entry of the chunk, row end
        };
        auto seek = [&](Cursor& c) {  // move c.r forward to the next non-empty row (or past the end)
            while (c.r < rows) {
                const int64_t row = p.row_order ? p.row_order[c.r] : c.r;
                const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
                if (lo != hi) { c.base = lo; c.hi = hi; return; }
                c.r += step;
            }
        };
        auto advance = [&](Cursor& c) {
            c.base += CHUNK;
            if (c.base >= c.hi) { c.r += step; seek(c); }
        };
        // raw index/weight of this lane's entry: loaded two chunks ahead, first touched one chunk
        // later (the in-order pipe would otherwise stall on the load right here)
        struct Raw { int idx; float d; };
        struct Meta { int64_t off; float sq, dp1; };
        bool saw_negative = false;
        auto load_raw = [&](const Cursor& c) {
            Raw rw{-1, 0.f};
            if (c.r < rows) {
                const int64_t e = c.base + lane;
                if (e < c.hi) { rw.d = __ldg(p.data + e); rw.idx = __ldg(p.indices + e); }
            }
            return rw;
        };
        auto to_meta = [&](const Raw& rw) {
            Meta mt{0, 0.f, 0.f};
            if (rw.idx >= 0) {
                mt.off = (int64_t)rw.idx * p.ldy;
                if (rw.d < 0.f) saw_negative = true;
                mt.sq = sqrtf(fabsf(rw.d));
                mt.dp1 = __fadd_rn(rw.d, 1.0f);
            }
            return mt;
        };
        Cursor c0{first, 0, 0};
        seek(c0);
        Cursor c1 = c0;
        if (c1.r < rows) advance(c1);
        Cursor c2 = c1;
        if (c2.r < rows) advance(c2);
        Meta m0 = to_meta(load_raw(c0));
        Raw r1 = load_raw(c1);
        float v0[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) v0[j] = __ldg(p.Y + __shfl_sync(0xffffffffu, m0.off, j) + m);
        uint32_t chunk_n = 0, row_n = 0;
        double bacc = 0.0;
        const bool prof = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
        long long t_empty = 0, t_bempty = 0, t_start = prof ? clock64() : 0, tt = 0;
        while (c0.r < rows) {
            // issue the next chunk's factor loads and the chunk-after-next's index loads
            const Meta m1 = to_meta(r1);  // loaded one iteration ago
            float v1[CHUNK];
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) v1[j] = __ldg(p.Y + __shfl_sync(0xffffffffu, m1.off, j) + m);
            const Raw r2 = load_raw(c2);  // first use: next iteration
            const int s = chunk_n % NSTAGE;
            const uint32_t ph = (chunk_n / NSTAGE) & 1u;
            if (prof) tt = clock64();
            mbar_wait(bar_empty(s), ph ^ 1u);
            if (prof) t_empty += clock64() - tt;
            uint8_t* tile_h = smem + OFF_STAGES + s * STAGE_BYTES + m * 128;
            uint8_t* tile_l = tile_h + TILE_BYTES;
            float part = 0.f;
#pragma unroll
            for (int g = 0; g < CHUNK / 4; ++g) {
                float zh[4], zl[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = g * 4 + jj;
                    const float sj = __shfl_sync(0xffffffffu, m0.sq, j);
                    const float cj = __shfl_sync(0xffffffffu, m0.dp1, j);
                    const float z = sj * v0[j];
                    zh[jj] = tf32_round(z);
                    zl[jj] = tf32_round(z - zh[jj]);
                    part = fmaf(cj, v0[j], part);
                }
                const int sw = (g ^ (m & 7)) << 4;  // 128B swizzle: 16-byte chunk index XOR (row mod 8)
                *reinterpret_cast<float4*>(tile_h + sw) = make_float4(zh[0], zh[1], zh[2], zh[3]);
                *reinterpret_cast<float4*>(tile_l + sw) = make_float4(zl[0], zl[1], zl[2], zl[3]);
            }
            bacc += (double)part;
            fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(bar_full(s));
            ++chunk_n;
            if (c0.base + CHUNK >= c0.hi) {  // last chunk of its row: hand the rhs to the solver group
                const int b = row_n & 1;
                const uint32_t bph = (row_n >> 1) & 1u;
                if (prof) tt = clock64();
                mbar_wait(bar_b_empty(b), bph ^ 1u);
                if (prof) t_bempty += clock64() - tt;
                bvec[b * F + m] = (float)bacc;
                mbar_arrive(bar_b_full(b));
                bacc = 0.0;
                ++row_n;
            }
            c0 = c1; c1 = c2;
            if (c2.r < rows) advance(c2);
            m0 = m1; r1 = r2;
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) v0[j] = v1[j];
        }
        if (saw_negative) atomicOr(flags, 1);
        if (prof) { p.prof[0] = clock64() - t_start; p.prof[1] = t_empty; p.prof[2] = t_bempty; p.prof[3] = chunk_n; p.prof[4] = row_n; }
    } else if (warp == MMA_WARP) {
        // =============================== MMA ISSUE ===============================
        if (lane == 0) {
            uint32_t chunk_n = 0, row_n = 0;
            const bool prof = p.prof != nullptr && blockIdx.x == 0;
            long long t_full = 0, t_accempty = 0, t_start = prof ? clock64() : 0, tt = 0;
            for (int64_t r = first; r < rows; r += step) {
                const int64_t row = p.row_order ? p.row_order[r] : r;
                const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
                if (lo == hi) continue;
                const int b = row_n & 1;
                const uint32_t aph = (row_n >> 1) & 1u;
                if (prof) tt = clock64();
                mbar_wait(bar_acc_empty(b), aph ^ 1u);
                if (prof) t_accempty += clock64() - tt;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * 256);
                uint32_t accumulate = 0;
                for (int64_t base = lo; base < hi; base += CHUNK, ++chunk_n) {
                    const int s = chunk_n % NSTAGE;
                    const uint32_t ph = (chunk_n / NSTAGE) & 1u;
                    if (prof) tt = clock64();
                    mbar_wait(bar_full(s), ph);
                    if (prof) t_full += clock64() - tt;
                    tc_fence_after();
                    const int kc = (int)((hi - base) < CHUNK ? (hi - base) : CHUNK);
                    const int nk = (kc + 7) >> 3;
                    const uint32_t tile = smem_base + OFF_STAGES + s * STAGE_BYTES;
                    const uint64_t desc = umma_desc(tile);
                    for (int k = 0; k < nk; ++k) {
                        // advance 8 fp32 = 32 B along K inside the 128-B swizzle row
                        const uint64_t dk = desc + (uint64_t)(k * 2);
                        umma_tf32(d_tmem, dk, dk, IDESC_TF32_M128_N256, accumulate);
                        accumulate = 1;
                    }
                    tc_commit(bar_empty(s));
                }
                tc_commit(bar_acc_full(b));
                ++row_n;
            }
            if (prof) { p.prof[8] = clock64() - t_start; p.prof[9] = t_full; p.prof[10] = t_accempty; }
        }
    } else {
        // =============================== SOLVE ===============================
        // Two groups of 128 threads; group g owns the CTA's non-empty rows with (row_n & 1) == g,
        // which is also the accumulator buffer it drains.
        const int g = (warp - 5) >> 2;
        const int q = warp & 3;        // TMEM lane quarter this warp may read
        const int i = q * 32 + lane;   // matrix row owned by this thread (0..127)
        float* A = reinterpret_cast<float*>(smem + OFF_A + g * A_BYTES);
        float* LpT = reinterpret_cast<float*>(smem + OFF_LPT + g * LPT_BYTES);
        float* dinv = reinterpret_cast<float*>(smem + OFF_DINV + g * F * 4);
        float* xs = reinterpret_cast<float*>(smem + OFF_XS + g * F * 4);
        constexpr int LDA = s128::LDA;
        auto gsync = [&]() {
            if (g == 0) s128::group_sync<1>(); else s128::group_sync<2>();
        };
        uint32_t row_n = 0;
        const bool prof = p.prof != nullptr && blockIdx.x == 0 && g == 0 && i == 0;
        long long t_accfull = 0, t_drain = 0, t_g = 0, t_bfull = 0, t_solve = 0, t_start = prof ? clock64() : 0, tt = 0;
        for (int64_t r = first; r < rows; r += step) {
            const int64_t row = p.row_order ? p.row_order[r] : r;
            const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
            float* xout = p.X + row * p.ldx;
            if (lo == hi) {  // wmf_model.py:223-225
                if (g == 0) xout[i] = 0.0f;
                continue;
            }
            const int b = row_n & 1;
            const uint32_t ph = (row_n >> 1) & 1u;
            ++row_n;
            if (b != g) continue;
            if (prof) tt = clock64();
            mbar_wait(bar_acc_full(b), ph);
            if (prof) { t_accfull += clock64() - tt; tt = clock64(); }
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256);
            // phase 1: own row: A[i][j] = P[i][j] + Q[i][j] (j < i), P + 2Q on the diagonal
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                const int col0 = cb * 32;
                if (col0 > i) break;  // nothing at or left of the diagonal in this block (warp-uniform: i>>5)
                float pv[32], qv[32];
                tmem_ld32(t_row + col0, pv);
                tmem_ld32(t_row + 128 + col0, qv);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const int j0 = col0 + c4 * 4;
                    if (j0 <= i) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float qq = qv[c4 * 4 + e];
                            o[e] = pv[c4 * 4 + e] + qq + ((j0 + e == i) ? qq : 0.0f);
                        }
                        *reinterpret_cast<float4*>(A + i * LDA + j0) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
            gsync();
            // phase 2: transposed part: A[j][i] += Q[i][j] (j > i); exactly one writer per element
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                const int col0 = cb * 32;
                if (col0 + 31 <= (i & ~31)) continue;  // whole block at or left of this warp's rows (warp-uniform)
                float qv[32];
                tmem_ld32(t_row + 128 + col0, qv);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int j = col0 + c;
                    if (j > i) A[j * LDA + i] += qv[c];
                }
            }
            tc_fence_before();
            mbar_arrive(bar_acc_empty(b));  // the tensor core may start the row after next
            gsync();
            if (prof) { t_drain += clock64() - tt; tt = clock64(); }
            // + G once (wmf_model.py:239 adds YTY_I to the finished weighted Gram), rhs into row 128
            // (batches of 8 independent 16-byte loads: G lives in L2, one exposed latency per batch)
#pragma unroll 1
            for (int e0 = i; e0 < F * (F / 4); e0 += s128::GROUP * 8) {
                float4 gv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e4 = e0 + u * s128::GROUP;
                    const int rr = e4 >> 5, c4 = (e4 & 31) * 4;
                    gv[u] = (c4 <= rr) ? __ldg(reinterpret_cast<const float4*>(p.G + rr * F + c4)) : make_float4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e4 = e0 + u * s128::GROUP;
                    const int rr = e4 >> 5, c4 = (e4 & 31) * 4;
                    if (c4 <= rr) {
                        float4* dst = reinterpret_cast<float4*>(A + rr * LDA + c4);
                        float4 v = *dst;
                        v.x = __fadd_rn(v.x, gv[u].x); v.y = __fadd_rn(v.y, gv[u].y);
                        v.z = __fadd_rn(v.z, gv[u].z); v.w = __fadd_rn(v.w, gv[u].w);
                        *dst = v;
                    }
                }
            }
            if (prof) { t_g += clock64() - tt; tt = clock64(); }
            mbar_wait(bar_b_full(b), ph);
            if (prof) { t_bfull += clock64() - tt; tt = clock64(); }
            A[F * LDA + i] = bvec[b * F + i];
            mbar_arrive(bar_b_empty(b));
            gsync();
            bool ok;
            if (g == 0) ok = s128::chol_solve_128<1>(A, LpT, dinv, xs, i);
            else ok = s128::chol_solve_128<2>(A, LpT, dinv, xs, i);
            if (ok) xout[i] = xs[i];
            else if (i == 0) atomicOr(flags, 2);  // not positive definite: the SIMT/LU kernel redoes the half-step
            gsync();
            if (prof) t_solve += clock64() - tt;
        }
        if (prof) { p.prof[16] = clock64() - t_start; p.prof[17] = t_accfull; p.prof[18] = t_drain; p.prof[19] = t_g; p.prof[20] = t_bfull; p.prof[21] = t_solve; p.prof[22] = row_n; }
    }
    // =============================== TEARDOWN ===============================
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

bool tc_half_step_supported(int f, int bias) { return f == tc::F && !bias; }

size_t tc_half_step_workspace_bytes(int64_t, int f, int) { return simt_half_step_workspace_bytes(f); }

int tc_half_step(const HalfStepParams& in, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t need = tc_half_step_workspace_bytes(in.rows, in.f, in.bias);
    if (ws == nullptr || ws_bytes < need) {
        set_error("wmf_als_half_step(tcgen05): workspace %zu < %zu", ws_bytes, need);
        return WMF_ERR_WORKSPACE;
    }
    WMF_CUDA(cudaMemsetAsync(ws, 0, 256, st));
    int* flags = reinterpret_cast<int*>(ws) + 1;  // [0] = SIMT row counter, [1] = redo flags
    WMF_CUDA(cudaFuncSetAttribute(als_half_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    HalfStepParams p = in;
    p.prof = getenv("WMF_TC_PROFILE") ? reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + 64) : nullptr;
    int grid = sm_count();
    if ((int64_t)grid > in.rows) grid = (int)in.rows;
    als_half_step_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(p, flags);
    WMF_LAUNCH_CHECK("als_half_step_tc_kernel");
    // fix-up: runs the FP32/LU kernel over the whole half-step only if a flag was raised
    HalfStepParams fix = in;
    fix.run_if = flags;
    return simt_half_step(fix, ws, ws_bytes, st);
}

}  // namespace wmf
