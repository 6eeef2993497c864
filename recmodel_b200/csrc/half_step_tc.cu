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
#include "common.cuh"
#include "half_step.cuh"
#include "solve.cuh"

namespace wmf {

namespace tc {

constexpr int F = 128;               // factor width handled by this kernel
constexpr int CHUNK = 32;            // stored entries per staged tile (= one 128-byte swizzle row)
constexpr int NSTAGE = 4;
constexpr int TILE_BYTES = F * 128;  // 128 rows (features) x 32 fp32 (K) = 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;  // [zh ; zl]
constexpr int LDA = F + 1;           // 129: odd leading dimension, conflict-free both ways
constexpr int GATHER_THREADS = 128, MMA_WARP = 4, SOLVER_THREADS = 256;
constexpr int THREADS = GATHER_THREADS + 32 + SOLVER_THREADS;  // 416
constexpr int SOLVER_BAR_ID = 1;
constexpr uint32_t TMEM_COLS = 512;

// shared memory carve-up (bytes from a 1024-aligned base)
constexpr int OFF_STAGES = 0;
constexpr int OFF_A = OFF_STAGES + NSTAGE * STAGE_BYTES;            // (F+1) x LDA floats
constexpr int A_BYTES = ((F + 1) * LDA * 4 + 15) / 16 * 16;
constexpr int OFF_BVEC = OFF_A + A_BYTES;                           // 2 x F floats
constexpr int OFF_DIAG = OFF_BVEC + 2 * F * 4;                      // F floats
constexpr int OFF_BARS = OFF_DIAG + F * 4;                          // mbarriers (8 B each)
constexpr int NBARS = 2 * NSTAGE + 8;
constexpr int OFF_TMEM_PTR = OFF_BARS + NBARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;                // + slack for alignment

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
    float* A = reinterpret_cast<float*>(smem + OFF_A);
    float* bvec = reinterpret_cast<float*>(smem + OFF_BVEC);
    float* diag = reinterpret_cast<float*>(smem + OFF_DIAG);
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
            mbar_init(bar_acc_empty(b), SOLVER_THREADS);
            mbar_init(bar_b_full(b), GATHER_THREADS);
            mbar_init(bar_b_empty(b), SOLVER_THREADS);
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
        const int m = tid;  // feature index
        uint32_t chunk_n = 0, row_n = 0;
        bool saw_negative = false;
        for (int64_t r = first; r < rows; r += step) {
            const int64_t row = p.row_order ? p.row_order[r] : r;
            const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
            if (lo == hi) continue;
            double bacc = 0.0;
            for (int64_t base = lo; base < hi; base += CHUNK, ++chunk_n) {
                const int s = chunk_n % NSTAGE;
                const uint32_t ph = (chunk_n / NSTAGE) & 1u;
                // this lane's entry of the chunk (every warp keeps its own copy: no cross-warp sync)
                const int64_t e = base + lane;
                int64_t off = 0;
                float sq = 0.f, dp1 = 0.f;
                if (e < hi) {
                    const float d = p.data[e];
                    off = (int64_t)p.indices[e] * p.ldy;
                    if (d < 0.f) saw_negative = true;
                    sq = sqrtf(fabsf(d));
                    dp1 = __fadd_rn(d, 1.0f);
                }
                // issue all 32 factor loads first (32 independent 128-B lines per warp in flight)
                float v[CHUNK];
#pragma unroll
                for (int j = 0; j < CHUNK; ++j) {
                    const int64_t oj = __shfl_sync(0xffffffffu, off, j);
                    v[j] = __ldg(p.Y + oj + m);
                }
                mbar_wait(bar_empty(s), ph ^ 1u);
                uint8_t* tile_h = smem + OFF_STAGES + s * STAGE_BYTES + m * 128;
                uint8_t* tile_l = tile_h + TILE_BYTES;
                float part = 0.f;
#pragma unroll
                for (int g = 0; g < CHUNK / 4; ++g) {
                    float zh[4], zl[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = g * 4 + jj;
                        const float sj = __shfl_sync(0xffffffffu, sq, j);
                        const float cj = __shfl_sync(0xffffffffu, dp1, j);
                        const float z = sj * v[j];
                        zh[jj] = tf32_round(z);
                        zl[jj] = tf32_round(z - zh[jj]);
                        part = fmaf(cj, v[j], part);
                    }
                    const int sw = (g ^ (m & 7)) << 4;  // 128B swizzle: 16-byte chunk index XOR (row mod 8)
                    *reinterpret_cast<float4*>(tile_h + sw) = make_float4(zh[0], zh[1], zh[2], zh[3]);
                    *reinterpret_cast<float4*>(tile_l + sw) = make_float4(zl[0], zl[1], zl[2], zl[3]);
                }
                bacc += (double)part;
                fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
                mbar_arrive(bar_full(s));
            }
            const int b = row_n & 1;
            const uint32_t bph = (row_n >> 1) & 1u;
            mbar_wait(bar_b_empty(b), bph ^ 1u);
            bvec[b * F + m] = (float)bacc;
            mbar_arrive(bar_b_full(b));
            ++row_n;
        }
        if (saw_negative) atomicOr(flags, 1);
    } else if (warp == MMA_WARP) {
        // =============================== MMA ISSUE ===============================
        if (lane == 0) {
            uint32_t chunk_n = 0, row_n = 0;
            for (int64_t r = first; r < rows; r += step) {
                const int64_t row = p.row_order ? p.row_order[r] : r;
                const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
                if (lo == hi) continue;
                const int b = row_n & 1;
                const uint32_t aph = (row_n >> 1) & 1u;
                mbar_wait(bar_acc_empty(b), aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * 256);
                uint32_t accumulate = 0;
                for (int64_t base = lo; base < hi; base += CHUNK, ++chunk_n) {
                    const int s = chunk_n % NSTAGE;
                    const uint32_t ph = (chunk_n / NSTAGE) & 1u;
                    mbar_wait(bar_full(s), ph);
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
        }
    } else {
        // =============================== SOLVE ===============================
        const int st = tid - (GATHER_THREADS + 32);  // 0..255
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int half = (warp - 5) >> 2;            // which 64-column half it converts
        const int i = q * 32 + lane;                 // matrix row held by this thread
        NamedSync<SOLVER_BAR_ID, SOLVER_THREADS> sync;
        uint32_t row_n = 0;
        for (int64_t r = first; r < rows; r += step) {
            const int64_t row = p.row_order ? p.row_order[r] : r;
            const int64_t lo = p.indptr[row], hi = p.indptr[row + 1];
            float* xout = p.X + row * p.ldx;
            if (lo == hi) {  // wmf_model.py:223-225
                for (int c = st; c < F; c += SOLVER_THREADS) xout[c] = 0.0f;
                continue;
            }
            const int b = row_n & 1;
            const uint32_t ph = (row_n >> 1) & 1u;
            mbar_wait(bar_acc_full(b), ph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256);
            // phase 1: own row, lower part: A[i][j] = P[i][j] + Q[i][j]   (j <= i)
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
                const int col0 = half * 64 + cb * 32;
                float pv[32], qv[32];
                tmem_ld32(t_row + col0, pv);
                tmem_ld32(t_row + 128 + col0, qv);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int j = col0 + c;
                    if (j < i) A[i * LDA + j] = pv[c] + qv[c];
                    else if (j == i) A[i * LDA + j] = pv[c] + 2.0f * qv[c];  // diagonal: Q + Q^T
                }
            }
            sync();
            // phase 2: transposed part: A[j][i] += Q[i][j]   (j > i); one writer per element
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
                const int col0 = half * 64 + cb * 32;
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
            sync();
            // + G once (wmf_model.py:239 adds YTY_I to the finished weighted Gram), rhs into row F
            for (int e = st; e < F * F; e += SOLVER_THREADS) {
                const int ii = e >> 7, jj = e & (F - 1);
                if (jj <= ii) A[ii * LDA + jj] = __fadd_rn(A[ii * LDA + jj], __ldg(p.G + e));
            }
            mbar_wait(bar_b_full(b), ph);
            if (st < F) A[F * LDA + st] = bvec[b * F + st];
            mbar_arrive(bar_b_empty(b));
            sync();
            const bool ok = chol_factor_aug<SOLVER_THREADS>(A, LDA, F, diag, st, sync);
            if (ok) {
                chol_back_solve(A, LDA, F, diag, xout, st);
            } else if (st == 0) {
                atomicOr(flags, 2);  // not positive definite: the SIMT/LU kernel redoes the half-step
            }
            sync();
            ++row_n;
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
