// K2 (tcgen05 version): one ALS half-step for dim = 128, weighted, no bias.
// Replaces the per-row loop of recompute_factors (wmf_model.py:220-239).
//
// Per CSR row r:  A_r = G + sum_j d_j y_j y_j^T  (128x128, K = n_r),  b_r = sum_j (d_j+1) y_j,
// x_r = A_r^-1 b_r.  Everything O(f^2) per entry and O(f^3) per row runs on the 5th-gen tensor
// cores with FP32-equivalent accuracy (3xTF32: plain TF32 fails the 1e-4 bar):
//
//   Gram      z_j = sqrt(d_j) y_j = zh + zl (both TF32-exact);  W = sum zh zh^T + zh zl^T + zl zh^T
//             -> three tcgen05.mma (kind::tf32, M=N=128, K=8) per 8 stored entries into a
//             128-column fp32 TMEM accumulator; operand tiles K-major, 128-byte swizzled.
//   Cholesky  the matrix STAYS in TMEM. Panels of 8 columns: each thread (= TMEM lane = matrix
//             row) loads its 8 panel entries (tcgen05.ld), adds G lazily, the 8x8 diagonal block
//             is factored in registers, the row is solved against it, the panel L (hi/lo split)
//             goes to shared memory as an MMA operand and the rank-8 trailing update
//             S -= L L^T is again three tcgen05.mma (negated A) on the whole 128x128 accumulator.
//             The right-hand side rides along in registers (forward substitution fused);
//             the finished L is written back over the dead columns (tcgen05.st) and read once
//             more, 8 rows at a time, for the back substitution.
//
// TMEM holds four such matrices (4 x 128 = 512 columns): four solver groups of 128 threads each
// own one accumulator, so while one row is in its pivot chain three others fill the issue slots.
// One persistent CTA per SM, warp-specialised, all hand-offs through mbarriers:
//   warps 0-3   gather   thread m owns feature m: loads Y[idx_j][m] for 32 entries (one coalesced
//               128-B line per warp and entry), forms zh/zl, stores them with the XOR swizzle a TMA
//               load would have produced (TMA cannot: the operand is gathered, scaled and split),
//               keeps the rhs b[m] in registers.
//   warps 4-19  solve    group g = (warp-4)/4 owns accumulator g and the rows n with n % 4 == g.
//   warp 20     MMA      one thread issues the Gram MMAs and commits stage/accumulator barriers.
//
// Rows whose weights are negative (sqrt undefined) or whose Cholesky meets a non-positive
// pivot raise a flag; the caller then re-runs the half-step with the SIMT kernel (LU).
#include <stdlib.h>
#include "common.cuh"
#include "half_step.cuh"
#include "factor8.cuh"

namespace wmf {

namespace tc {

constexpr int F = 128;               // factor width handled by this kernel
constexpr int CHUNK = 32;            // stored entries per staged tile (= one 128-byte swizzle row)
constexpr int NSTAGE = 3;            // operand-tile stages
constexpr int NSTG = 3;              // raw gather staging buffers (cp.async depth)
constexpr int TILE_BYTES = F * 128;  // 128 rows (features) x 32 fp32 (K) = 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;  // [zh ; zl]
constexpr int STG_BYTES = CHUNK * F * 4;     // 32 gathered factor rows, row-major, 16 KB
constexpr int NB = 8;                // Cholesky panel width
constexpr int NGROUP = 4;            // solver groups = TMEM accumulators
constexpr int GROUP = 128;
constexpr int GATHER_THREADS = 128;
constexpr int SOLVER_WARP0 = 0, GATHER_WARP0 = NGROUP * 4, MMA_WARP = GATHER_WARP0 + 4;  // high warp ids issue first
constexpr int THREADS = (MMA_WARP + 1) * 32;               // 672
constexpr uint32_t TMEM_COLS = 512;
static_assert(NGROUP * F <= 512, "TMEM columns");

// per-group shared memory
constexpr int PANEL_TILE_BYTES = F * NB * 4;          // 4 KB: 128 rows x 8 fp32, K-major, no swizzle
constexpr int G_OFF_TILEH = 0;
constexpr int G_OFF_TILEL = G_OFF_TILEH + PANEL_TILE_BYTES;
constexpr int G_OFF_NINV = G_OFF_TILEL + PANEL_TILE_BYTES;          // 16 blocks x (8 x 8) floats: N = L^-1 per pivot block
constexpr int G_OFF_ZB = G_OFF_NINV + (F / NB) * NB * NB * 4;       // 16 x 8 floats: N b_blk
constexpr int G_OFF_DBLK = G_OFF_ZB + (F / NB) * NB * 4;            // 8 x 8 pivot block + 8 rhs
constexpr int G_OFF_BFIN = G_OFF_DBLK + (NB * NB + 2 * NB) * 4;     // 128 floats: final rhs
constexpr int GROUP_BYTES = ((G_OFF_BFIN + F * 4 + 127) / 128) * 128;

// shared memory carve-up (bytes from a 1024-aligned base)
constexpr int OFF_STAGES = 0;
constexpr int OFF_STG = OFF_STAGES + NSTAGE * STAGE_BYTES;
constexpr int OFF_META = OFF_STG + NSTG * STG_BYTES;                // NSTG x 32 x float2
constexpr int OFF_GROUPS = OFF_META + NSTG * CHUNK * 8;
constexpr int OFF_BVEC = OFF_GROUPS + NGROUP * GROUP_BYTES;         // NGROUP x F floats
constexpr int OFF_BARS = OFF_BVEC + NGROUP * F * 4;                 // mbarriers (8 B each)
constexpr int NBARS = 2 * NSTAGE + NSTG + 5 * NGROUP;
constexpr int OFF_TMEM_PTR = OFF_BARS + NBARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;                // + slack for alignment
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(OFF_GROUPS % 128 == 0 && GROUP_BYTES % 128 == 0 && OFF_BARS % 8 == 0 && OFF_STG % 1024 == 0, "alignment");

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


// K-major, no swizzle: 8-row x 16-byte core matrices; the two K-chunks of a row group are
// LBO = 128 B apart, consecutive 8-row groups SBO = 256 B apart (panel operand, K = 8 fp32).
__device__ __forceinline__ uint64_t umma_desc_panel(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
           (1ull << 46);
}
// kind::tf32, fp32 accumulate, K-major A and B, M = N = 128; NEG: A negated (trailing update)
constexpr uint32_t IDESC_TF32_M128_N128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_TF32_NEG_M128 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((128u >> 4) << 24);  // N filled in at issue
#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {  // arrive when this thread's copies have landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// explicit shared-space accesses (generic ld/st on pointers derived from the aligned base cost an
// address-space check per access)
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds1(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts2(uint32_t a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts1(uint32_t a, float x) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory");
}
__device__ __forceinline__ void group_bar(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(GROUP) : "memory");
}

}  // namespace tc

using namespace tc;

template <bool PROF>
__global__ void __launch_bounds__(THREADS, 1) als_half_step_tc_kernel(HalfStepParams p, int* __restrict__ flags) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t smem_base = smem_u32(smem);
    float* bvec = reinterpret_cast<float*>(smem + OFF_BVEC);
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_TMEM_PTR);
    const uint32_t bars = smem_base + OFF_BARS;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 8u * (NSTAGE + s); };
    auto bar_stg = [&](int s) { return bars + 8u * (2 * NSTAGE + s); };
    auto bar_acc_full = [&](int g) { return bars + 8u * (2 * NSTAGE + NSTG + g); };
    auto bar_acc_empty = [&](int g) { return bars + 8u * (2 * NSTAGE + NSTG + NGROUP + g); };
    auto bar_b_full = [&](int g) { return bars + 8u * (2 * NSTAGE + NSTG + 2 * NGROUP + g); };
    auto bar_b_empty = [&](int g) { return bars + 8u * (2 * NSTAGE + NSTG + 3 * NGROUP + g); };
    auto bar_panel = [&](int g) { return bars + 8u * (2 * NSTAGE + NSTG + 4 * NGROUP + g); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_full(s), GATHER_THREADS); mbar_init(bar_empty(s), 1); }
        for (int s = 0; s < NSTG; ++s) mbar_init(bar_stg(s), GATHER_THREADS);
        for (int g = 0; g < NGROUP; ++g) {
            mbar_init(bar_acc_full(g), 1);
            mbar_init(bar_acc_empty(g), GROUP);
            mbar_init(bar_b_full(g), GATHER_THREADS);
            mbar_init(bar_b_empty(g), GROUP);
            mbar_init(bar_panel(g), 1);
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

    const int64_t rows = p.sched_len;  // schedule slots; slot s belongs to CTA s % gridDim
    const int64_t first = blockIdx.x, step = gridDim.x;

    if (warp >= GATHER_WARP0 && warp < MMA_WARP) {
        // =============================== GATHER + GRAM MMA ISSUE ===============================
        // Flattened (row, chunk) sequence of this CTA, three chunks deep:
        //   chunk i+3: index/weight of the entries -> registers (lanes 0-7 of each warp, 8 entries per warp)
        //   chunk i+2: 32 x 512-B factor rows -> raw staging buffer with cp.async (16 B per lane,
        //              one coalesced row per warp instruction, no registers held), completion on an mbarrier
        //   chunk i  : thread m reads column m of the staged rows, scales, splits to TF32 hi/lo, stores the
        //              K-major swizzled operand tiles; after the gather barrier thread 0 issues the MMAs.
        const int m = tid - GATHER_WARP0 * 32;  // feature index
        const int gw = warp - GATHER_WARP0;     // gather warp 0..3
        struct Cursor {
            int64_t r, base, hi;  // position in the CTA's row list, first entry of the chunk, row end
        };
        auto seek = [&](Cursor& c) {  // move c.r forward to the next non-empty row (or past the end)
            while (c.r < rows) {
                const int64_t row = p.row_order ? p.row_order[c.r] : c.r;
                if (row >= 0) {  // -1 = padding slot of the balanced schedule
                    const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
                    if (lo != hi) { c.base = lo; c.hi = hi; return; }
                }
                c.r += step;
            }
        };
        auto advance = [&](Cursor& c) {
            if (c.r >= rows) return;
            c.base += CHUNK;
            if (c.base >= c.hi) { c.r += step; seek(c); }
        };
        struct Raw { int idx; float d; };
        bool saw_negative = false;
        auto load_raw = [&](const Cursor& c) {  // lane l < 8 of warp w: entry 8w + l of the chunk
            Raw rw{-1, 0.f};
            if (c.r < rows && lane < 8) {
                const int64_t e = c.base + gw * 8 + lane;
                if (e < c.hi) { rw.d = __ldg(p.data + e); rw.idx = __ldg(p.indices + e); }
            }
            return rw;
        };
        float2* metaS = reinterpret_cast<float2*>(smem + OFF_META);
        auto issue = [&](const Raw& rw, int buf) {
            if (lane < 8) {
                float sq = 0.f, dp1 = 0.f;
                if (rw.idx >= 0) {
                    if (rw.d < 0.f) saw_negative = true;
                    sq = sqrtf(fabsf(rw.d));
                    dp1 = __fadd_rn(rw.d, 1.0f);
                }
                metaS[buf * CHUNK + gw * 8 + lane] = make_float2(sq, dp1);
            }
            const uint32_t dst0 = smem_base + OFF_STG + buf * STG_BYTES + (gw * 8) * (F * 4) + lane * 16;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int idx = __shfl_sync(0xffffffffu, rw.idx, jj);
                const float* src = p.Y + (int64_t)(idx >= 0 ? idx : 0) * p.ldy + lane * 4;
                cp_async16(dst0 + jj * (F * 4), src, idx >= 0 ? 16u : 0u);  // size 0 -> zero fill
            }
            cp_async_arrive(bar_stg(buf));
        };
        Cursor c0{first, 0, 0};
        seek(c0);
        Cursor c1 = c0; advance(c1);
        Cursor c2 = c1; advance(c2);
        Cursor c3 = c2; advance(c3);
        issue(load_raw(c0), 0);
        issue(load_raw(c1), 1);
        Raw r2 = load_raw(c2);
        uint32_t chunk_n = 0, row_n = 0;
        double bacc = 0.0;
        const bool prof = PROF && blockIdx.x == 0 && m == 0;
        long long t_empty = 0, t_bempty = 0, t_stg = 0, t_acc = 0, t_start = prof ? clock64() : 0, tt = 0;
        long long t_issue = 0, t_xform = 0, t_bar = 0, t_mma = 0, t2 = 0;
        while (c0.r < rows) {
            if (prof) t2 = clock64();
            // buffer (i+2)%3 was read by every gather thread during iteration i-1: barrier before refilling it
            asm volatile("bar.sync %0, %1;" ::"n"(1 + NGROUP), "n"(GATHER_THREADS) : "memory");
            issue(r2, (chunk_n + 2) % NSTG);   // chunk i+2
            r2 = load_raw(c3);                 // chunk i+3, first touched next iteration
            const int s = chunk_n % NSTAGE, sb = chunk_n % NSTG;
            if (prof) { tt = clock64(); t_issue += tt - t2; }
            mbar_wait(bar_stg(sb), (chunk_n / NSTG) & 1u);
            if (prof) { t_stg += clock64() - tt; tt = clock64(); }
            mbar_wait(bar_empty(s), ((chunk_n / NSTAGE) & 1u) ^ 1u);
            if (prof) { t2 = clock64(); t_empty += t2 - tt; }
            const float* stg = reinterpret_cast<const float*>(smem + OFF_STG + sb * STG_BYTES) + m;
            const float2* mt = metaS + sb * CHUNK;
            uint8_t* tile_h = smem + OFF_STAGES + s * STAGE_BYTES + m * 128;
            uint8_t* tile_l = tile_h + TILE_BYTES;
            float part = 0.f;
#pragma unroll
            for (int g4 = 0; g4 < CHUNK / 4; ++g4) {
                float zh[4], zl[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = g4 * 4 + jj;
                    const float v = stg[j * F];
                    const float2 sc = mt[j];
                    const float z = sc.x * v;
                    zh[jj] = tf32_round(z);
                    zl[jj] = tf32_round(z - zh[jj]);
                    part = fmaf(sc.y, v, part);
                }
                const int sw = (g4 ^ (m & 7)) << 4;  // 128B swizzle: 16-byte chunk index XOR (row mod 8)
                *reinterpret_cast<float4*>(tile_h + sw) = make_float4(zh[0], zh[1], zh[2], zh[3]);
                *reinterpret_cast<float4*>(tile_l + sw) = make_float4(zl[0], zl[1], zl[2], zl[3]);
            }
            bacc += (double)part;
            if (prof) { tt = clock64(); t_xform += tt - t2; }
            fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(bar_full(s));
            if (prof) { t2 = clock64(); t_bar += t2 - tt; }
            const bool last_chunk = c0.base + CHUNK >= c0.hi;
            const int g = row_n % NGROUP;
            ++chunk_n;
            if (last_chunk) {  // hand the rhs to the solver group that owns this row
                const uint32_t bph = (row_n / NGROUP) & 1u;
                if (prof) tt = clock64();
                mbar_wait(bar_b_empty(g), bph ^ 1u);
                if (prof) t_bempty += clock64() - tt;
                bvec[g * F + m] = (float)bacc;
                mbar_arrive(bar_b_full(g));
                bacc = 0.0;
                ++row_n;
            }
            c0 = c1; c1 = c2; c2 = c3;
            advance(c3);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (saw_negative) atomicOr(flags, 1);
        if (prof) { p.prof[0] = clock64() - t_start; p.prof[1] = t_empty; p.prof[2] = t_bempty; p.prof[3] = chunk_n; p.prof[4] = row_n; p.prof[5] = t_stg; p.prof[6] = t_acc; p.prof[11] = t_issue; p.prof[12] = t_xform; p.prof[13] = t_bar; p.prof[14] = t_mma; }
    } else if (warp == MMA_WARP) {
        // =============================== GRAM MMA ISSUE ===============================
        if (lane == 0) {
            uint32_t chunk_n = 0, row_n = 0;
            const bool prof = PROF && blockIdx.x == 0;
            long long t_full = 0, t_accempty = 0, t_start = prof ? clock64() : 0, tt = 0;
            for (int64_t r = first; r < rows; r += step) {
                const int64_t row = p.row_order ? p.row_order[r] : r;
                if (row < 0) continue;
                const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
                if (lo == hi) continue;
                const int g = row_n % NGROUP;
                const uint32_t aph = (row_n / NGROUP) & 1u;
                if (prof) tt = clock64();
                mbar_wait(bar_acc_empty(g), aph ^ 1u);
                if (prof) t_accempty += clock64() - tt;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(g * F);
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
                    const uint64_t dh = umma_desc(tile), dl = umma_desc(tile + TILE_BYTES);
                    for (int k = 0; k < nk; ++k) {
                        // advance 8 fp32 = 32 B along K inside the 128-B swizzle row
                        const uint64_t hk = dh + (uint64_t)(k * 2), lk = dl + (uint64_t)(k * 2);
                        umma_tf32(d_tmem, hk, hk, IDESC_TF32_M128_N128, accumulate);  // zh zh^T
                        umma_tf32(d_tmem, hk, lk, IDESC_TF32_M128_N128, 1u);          // zh zl^T
                        umma_tf32(d_tmem, lk, hk, IDESC_TF32_M128_N128, 1u);          // zl zh^T
                        accumulate = 1;
                    }
                    tc_commit(bar_empty(s));
                }
                tc_commit(bar_acc_full(g));
                ++row_n;
            }
            if (prof) { p.prof[8] = clock64() - t_start; p.prof[9] = t_full; p.prof[10] = t_accempty; }
        }
    } else {
        // =============================== SOLVE (matrix resident in TMEM) ===============================
        // Block Gauss-Jordan on the 128x128 system, 8 columns per step (see the header): no back
        // substitution and nothing is written back to TMEM.
        const int g = (warp - SOLVER_WARP0) >> 2;
        const int q = warp & 3;        // TMEM lane quarter this warp may access
        const int t = q * 32 + lane;   // matrix row owned by this thread = TMEM lane
        const int bar_id = 1 + g;
        const uint32_t gs = smem_base + OFF_GROUPS + g * GROUP_BYTES;
        const uint32_t tileH = gs + G_OFF_TILEH, tileL = gs + G_OFF_TILEL;
        const uint32_t Nst = gs + G_OFF_NINV, zst = gs + G_OFF_ZB, Dblk = gs + G_OFF_DBLK, bfin = gs + G_OFF_BFIN;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * F);
        const uint32_t d_tmem = tmem_base + (uint32_t)(g * F);
        const uint64_t descH = umma_desc_panel(tileH), descL = umma_desc_panel(tileL);
        const float* Grow = p.G + t * F;
        uint32_t row_n = 0, my_rows = 0, panel_n = 0;
        const bool prof = PROF && blockIdx.x == 0 && g == 0 && t == 0;
        long long t_accfull = 0, t_fact = 0, t_back = 0, t_start = prof ? clock64() : 0, tt = 0;
        long long ph_wait = 0, ph_ld = 0, ph_own = 0, ph_p = 0, ph_issue = 0, t3 = 0, t4 = 0;
        for (int64_t r = first; r < rows; r += step) {
            const int64_t row = p.row_order ? p.row_order[r] : r;
            if (row < 0) continue;
            const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
            float* xout = p.X + row * p.ldx;
            if (lo == hi) {  // wmf_model.py:223-225
                if (g == 0) xout[t] = 0.0f;
                continue;
            }
            const bool mine = (int)(row_n % NGROUP) == g;
            ++row_n;
            if (!mine) continue;
            const uint32_t ph = my_rows & 1u;
            ++my_rows;
            if (prof) tt = clock64();
            mbar_wait(bar_b_full(g), ph);
            float bt = lds1(smem_base + OFF_BVEC + (g * F + t) * 4);
            mbar_arrive(bar_b_empty(g));
            mbar_wait(bar_acc_full(g), ph);
            tc_fence_after();
            if (prof) { t_accfull += clock64() - tt; tt = clock64(); }
#pragma unroll 1
            for (int c0 = 0; c0 < F; c0 += NB) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(Grow + c0));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(Grow + c0 + 4));
                if (prof) t3 = clock64();
                if (c0 > 0) {  // the previous step's rank-8 update has landed in TMEM
                    mbar_wait(bar_panel(g), panel_n & 1u);
                    ++panel_n;
                    tc_fence_after();
                }
                if (prof) { t4 = clock64(); ph_wait += t4 - t3; }
                float a[NB];
                tmem_ld8(t_row + c0, a);
                if (c0 + NB == F) {  // last read of the accumulator: the Gram of this group's next row may start
                    tc_fence_before();
                    mbar_arrive(bar_acc_empty(g));
                }
                a[0] = __fadd_rn(a[0], g0.x); a[1] = __fadd_rn(a[1], g0.y); a[2] = __fadd_rn(a[2], g0.z);
                a[3] = __fadd_rn(a[3], g0.w); a[4] = __fadd_rn(a[4], g1.x); a[5] = __fadd_rn(a[5], g1.y);
                a[6] = __fadd_rn(a[6], g1.z); a[7] = __fadd_rn(a[7], g1.w);
                if (prof) { t3 = clock64(); ph_ld += t3 - t4; }
                const int rel = t - c0;
                const uint32_t nd = Nst + (c0 >> 3) * 256, zd = zst + (c0 >> 3) * 32;
                if (q == (c0 >> 5)) {
                    // ---- owner warp: Cholesky of the 8x8 pivot block, its inverse N = L^-1, zb = N b_blk ----
                    if (rel >= 0 && rel < NB) {
                        sts4(Dblk + rel * 32, a[0], a[1], a[2], a[3]);
                        sts4(Dblk + rel * 32 + 16, a[4], a[5], a[6], a[7]);
                        sts1(Dblk + 256 + rel * 4, bt);
                    }
                    __syncwarp();
                    float d[36], bb[NB];
#pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        const float4 d0 = lds4(Dblk + i * 32);
                        d[TRI(i, 0)] = d0.x;
                        if (i >= 1) d[TRI(i, 1)] = d0.y;
                        if (i >= 2) d[TRI(i, 2)] = d0.z;
                        if (i >= 3) d[TRI(i, 3)] = d0.w;
                        if (i >= 4) {
                            const float4 d1 = lds4(Dblk + i * 32 + 16);
                            d[TRI(i, 4)] = d1.x;
                            if (i >= 5) d[TRI(i, 5)] = d1.y;
                            if (i >= 6) d[TRI(i, 6)] = d1.z;
                            if (i >= 7) d[TRI(i, 7)] = d1.w;
                        }
                    }
                    {
                        const float4 b0 = lds4(Dblk + 256), b1 = lds4(Dblk + 272);
                        bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
                        bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
                    }
                    Factor8 fo;
                    const bool ok = factor8(d, bb, fo);
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < NB; ++i) {
                            sts4(nd + i * 32, fo.n[TRI(i, 0)], i >= 1 ? fo.n[TRI(i, 1)] : 0.f, i >= 2 ? fo.n[TRI(i, 2)] : 0.f,
                                 i >= 3 ? fo.n[TRI(i, 3)] : 0.f);
                            if (i >= 4)
                                sts4(nd + i * 32 + 16, fo.n[TRI(i, 4)], i >= 5 ? fo.n[TRI(i, 5)] : 0.f,
                                     i >= 6 ? fo.n[TRI(i, 6)] : 0.f, i >= 7 ? fo.n[TRI(i, 7)] : 0.f);
                        }
                        sts4(zd, fo.zb[0], fo.zb[1], fo.zb[2], fo.zb[3]);
                        sts4(zd + 16, fo.zb[4], fo.zb[5], fo.zb[6], fo.zb[7]);
                        if (!ok) atomicOr(flags, 2);  // not positive definite: the SIMT/LU kernel redoes the half-step
                    }
                    if (prof) { t4 = clock64(); ph_own += t4 - t3; }
                }
                group_bar(bar_id);
                // ---- every row outside the block: P = a N^T, rhs -= P zb; the block's own rows are pivots (P = 0) ----
                float P[NB];
                {
                    const bool pivot = rel >= 0 && rel < NB;
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const float4 n0 = lds4(nd + j * 32);
                        float v = a[0] * n0.x;
                        if (j >= 1) v = fmaf(a[1], n0.y, v);
                        if (j >= 2) v = fmaf(a[2], n0.z, v);
                        if (j >= 3) v = fmaf(a[3], n0.w, v);
                        if (j >= 4) {
                            const float4 n1 = lds4(nd + j * 32 + 16);
                            v = fmaf(a[4], n1.x, v);
                            if (j >= 5) v = fmaf(a[5], n1.y, v);
                            if (j >= 6) v = fmaf(a[6], n1.z, v);
                            if (j >= 7) v = fmaf(a[7], n1.w, v);
                        }
                        P[j] = pivot ? 0.0f : v;
                    }
                    const float4 z0 = lds4(zd), z1 = lds4(zd + 16);
                    float u0 = P[0] * z0.x, u1 = P[1] * z0.y;  // two chains, fixed order
                    u0 = fmaf(P[2], z0.z, u0); u1 = fmaf(P[3], z0.w, u1);
                    u0 = fmaf(P[4], z1.x, u0); u1 = fmaf(P[5], z1.y, u1);
                    u0 = fmaf(P[6], z1.z, u0); u1 = fmaf(P[7], z1.w, u1);
                    bt -= u0 + u1;
                }
                if (c0 + NB < F) {
                    float lh[NB], ll[NB];
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        lh[j] = tf32_round(P[j]);
                        ll[j] = tf32_round(P[j] - lh[j]);
                    }
                    const uint32_t o = (uint32_t)((t >> 3) * 256 + (t & 7) * 16);
                    sts4(tileH + o, lh[0], lh[1], lh[2], lh[3]);
                    sts4(tileH + o + 128, lh[4], lh[5], lh[6], lh[7]);
                    sts4(tileL + o, ll[0], ll[1], ll[2], ll[3]);
                    sts4(tileL + o + 128, ll[4], ll[5], ll[6], ll[7]);
                    fence_async_smem();
                    tc_fence_before();
                    if (prof) t3 = clock64();
                    group_bar(bar_id);
                    if (t == 0) {
                        // S[:, j] -= P P[j]^T for the live columns j >= c0 + 8 (all 128 rows: Gauss-Jordan)
                        tc_fence_after();
                        if (prof) { t4 = clock64(); ph_p += t4 - t3; }
                        const uint32_t start = (uint32_t)((c0 + NB) >> 4) << 4;
                        const uint32_t idesc = IDESC_TF32_NEG_M128 | (((F - start) >> 3) << 17);
                        const uint64_t bH = descH + (uint64_t)(start * 2), bL = descL + (uint64_t)(start * 2);
                        umma_tf32(d_tmem + start, descH, bH, idesc, 1u);
                        umma_tf32(d_tmem + start, descH, bL, idesc, 1u);
                        umma_tf32(d_tmem + start, descL, bH, idesc, 1u);
                        tc_commit(bar_panel(g));
                        if (prof) ph_issue += clock64() - t4;
                    }
                }
            }
            if (prof) { t_fact += clock64() - tt; tt = clock64(); }
            // ---- the system is block diagonal now: x_blk = (L L^T)^-1 b_blk = N^T (N b_blk) ----
            sts1(bfin + t * 4, bt);
            group_bar(bar_id);
            {
                const int r8 = t & 7;
                const uint32_t nb = Nst + (t >> 3) * 256, bq = bfin + (t >> 3) * 32;
                const float4 b0 = lds4(bq), b1 = lds4(bq + 16);
                const float bb[NB] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float xt = 0.0f;
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const float4 n0 = lds4(nb + j * 32), n1 = lds4(nb + j * 32 + 16);
                    const float nr[NB] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
                    float y = 0.0f;
#pragma unroll
                    for (int k = 0; k <= j; ++k) y = fmaf(nr[k], bb[k], y);
                    float nsel = 0.0f;  // N[j][r8] (zero above the diagonal)
#pragma unroll
                    for (int k = 0; k <= j; ++k) nsel = (k == r8) ? nr[k] : nsel;
                    xt = fmaf(nsel, y, xt);
                }
                xout[t] = xt;
            }
            group_bar(bar_id);  // Nst / bfin are rewritten by the next row
            if (prof) t_back += clock64() - tt;
        }
        if (prof) {
            p.prof[16] = clock64() - t_start; p.prof[17] = t_accfull; p.prof[18] = t_fact; p.prof[19] = t_back;
            p.prof[22] = my_rows; p.prof[24] = ph_wait; p.prof[25] = ph_ld; p.prof[26] = ph_own; p.prof[28] = ph_p;
            p.prof[31] = ph_issue;
        }
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

size_t tc_half_step_workspace_bytes(int64_t, int f, int) {  // header + profile slots (f = 128 needs no SIMT slab)
    const size_t a = simt_half_step_workspace_bytes(f);
    return a > 1024 ? a : 1024;
}

int tc_half_step(const HalfStepParams& in, void* ws, size_t ws_bytes, cudaStream_t st) {
    const size_t need = tc_half_step_workspace_bytes(in.rows, in.f, in.bias);
    if (ws == nullptr || ws_bytes < need) {
        set_error("wmf_als_half_step(tcgen05): workspace %zu < %zu", ws_bytes, need);
        return WMF_ERR_WORKSPACE;
    }
    WMF_CUDA(cudaMemsetAsync(ws, 0, 1024, st));
    int* flags = reinterpret_cast<int*>(ws) + 1;  // [0] = SIMT row counter, [1] = redo flags
    WMF_CUDA(cudaFuncSetAttribute(als_half_step_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    WMF_CUDA(cudaFuncSetAttribute(als_half_step_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    HalfStepParams p = in;
    p.prof = getenv("WMF_TC_PROFILE") ? reinterpret_cast<long long*>(reinterpret_cast<char*>(ws) + 256) : nullptr;
    int grid = sm_count();
    if ((int64_t)grid > in.rows) grid = (int)in.rows;
    if (p.prof) als_half_step_tc_kernel<true><<<grid, THREADS, SMEM_BYTES, st>>>(p, flags);
    else als_half_step_tc_kernel<false><<<grid, THREADS, SMEM_BYTES, st>>>(p, flags);
    WMF_LAUNCH_CHECK("als_half_step_tc_kernel");
    // fix-up: runs the FP32/LU kernel over the whole half-step only if a flag was raised
    HalfStepParams fix = in;
    fix.run_if = flags;
    return simt_half_step(fix, ws, ws_bytes, st);
}

}  // namespace wmf
