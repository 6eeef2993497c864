// Probe: cycle cost of each phase of one Gauss-Jordan step of half_step_tc.cu in isolation
// (1 CTA, NG solver groups of 128 threads, no gather / Gram traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I recmodel_b200/csrc -o scripts/probe/step_latency scripts/probe/step_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "factor8.cuh"
using namespace wmf::tc;
constexpr int F = 128, NB = 8;
#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t umma_desc_panel(uint32_t a) { return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46); }
constexpr uint32_t IDESC_TF32_NEG_M128 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t a) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts4(uint32_t a, float x, float y, float z, float w) { asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory"); }
__device__ __forceinline__ void sts1(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

constexpr int GROUP_BYTES = 16384;
__global__ void __launch_bounds__(512, 1) probe(long long* out, int iters, int ngroups, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = warp >> 2, q = warp & 3, t = q * 32 + lane;
    const uint32_t gs = base + g * GROUP_BYTES;
    const uint32_t tileH = gs, tileL = gs + 4096, Nst = gs + 8192, zst = gs + 12288, Dblk = gs + 12800, bar = gs + 13312;
    const uint32_t tmem_slot = base + 4 * GROUP_BYTES;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(base + i * GROUP_BYTES + 13312, mode == 1 ? 3 : 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem_base; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * F);
    const uint32_t d_tmem = tmem_base + (uint32_t)(g * F);
    const uint64_t descH = umma_desc_panel(tileH), descL = umma_desc_panel(tileL);
    // a diagonally dominant matrix in TMEM
    for (int c0 = 0; c0 < F; c0 += NB) { float a[NB]; for (int i = 0; i < NB; ++i) a[i] = (c0 + i == t) ? 300.0f : 0.01f * ((t * 7 + c0 + i) % 13); tmem_st8(t_row + c0, a); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t panel_n = 0;
    float bt = 1.0f;
    const bool rec = (tid == 0);
    if (g < ngroups)
    for (int it = 0; it < iters; ++it) {
        for (int c0 = 0; c0 < F; c0 += NB) {
            long long t0 = clock64(), t1;
            if (c0 > 0) { mbar_wait(bar, panel_n & 1u); ++panel_n; tc_fence_after(); }
            t1 = clock64(); if (rec) acc[0] += t1 - t0; t0 = t1;                      // 0: MMA round trip (after issue)
            float a[NB];
            tmem_ld8(t_row + c0, a);
            t1 = clock64(); if (rec) acc[1] += t1 - t0; t0 = t1;                      // 1: tcgen05.ld + wait
            const int rel = t - c0;
            const uint32_t nd = Nst + (c0 >> 3) * 256, zd = zst + (c0 >> 3) * 32;
            if (q == (c0 >> 5)) {
                if (rel >= 0 && rel < NB) { sts4(Dblk + rel * 32, a[0], a[1], a[2], a[3]); sts4(Dblk + rel * 32 + 16, a[4], a[5], a[6], a[7]); sts1(Dblk + 256 + rel * 4, bt); }
                __syncwarp();
                float d[36], bb[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const float4 d0 = lds4(Dblk + i * 32);
                    d[TRI(i, 0)] = d0.x; if (i >= 1) d[TRI(i, 1)] = d0.y; if (i >= 2) d[TRI(i, 2)] = d0.z; if (i >= 3) d[TRI(i, 3)] = d0.w;
                    if (i >= 4) { const float4 d1 = lds4(Dblk + i * 32 + 16); d[TRI(i, 4)] = d1.x; if (i >= 5) d[TRI(i, 5)] = d1.y; if (i >= 6) d[TRI(i, 6)] = d1.z; if (i >= 7) d[TRI(i, 7)] = d1.w; }
                }
                { const float4 b0 = lds4(Dblk + 256), b1 = lds4(Dblk + 272); bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w; }
                float ncol[NB], z[NB];
                factor8(d, bb, lane & 7, ncol, z);
                if (lane < NB) { for (int i = 0; i < NB; ++i) sts1(nd + i * 32 + lane * 4, ncol[i]); }
                if (lane == 0) { sts4(zd, z[0], z[1], z[2], z[3]); sts4(zd + 16, z[4], z[5], z[6], z[7]); }
            }
            t1 = clock64(); if (rec) acc[2] += t1 - t0; t0 = t1;                      // 2: owner factor (warp 0: 4 of 16 steps)
            named_bar(1 + g, 128);
            t1 = clock64(); if (rec) acc[3] += t1 - t0; t0 = t1;                      // 3: barrier A (incl. waiting for other owners)
            float P[NB];
#pragma unroll
            for (int jj = 0; jj < NB; ++jj) {
                const float4 n0 = lds4(nd + jj * 32);
                float v = a[0] * n0.x;
                if (jj >= 1) v = fmaf(a[1], n0.y, v); if (jj >= 2) v = fmaf(a[2], n0.z, v); if (jj >= 3) v = fmaf(a[3], n0.w, v);
                if (jj >= 4) { const float4 n1 = lds4(nd + jj * 32 + 16); v = fmaf(a[4], n1.x, v); if (jj >= 5) v = fmaf(a[5], n1.y, v); if (jj >= 6) v = fmaf(a[6], n1.z, v); if (jj >= 7) v = fmaf(a[7], n1.w, v); }
                P[jj] = (rel >= 0 && rel < NB) ? 0.0f : v;
            }
            { const float4 z0 = lds4(zd), z1 = lds4(zd + 16); bt -= P[0] * z0.x + P[1] * z0.y + P[2] * z0.z + P[3] * z0.w + P[4] * z1.x + P[5] * z1.y + P[6] * z1.z + P[7] * z1.w; }
            t1 = clock64(); if (rec) acc[4] += t1 - t0; t0 = t1;                      // 4: P product + rhs
            if (c0 + NB < F) {
                float lh[NB], ll[NB];
                for (int jj = 0; jj < NB; ++jj) { lh[jj] = tf32_round(P[jj]); ll[jj] = P[jj] - lh[jj]; }
                const uint32_t o = (uint32_t)((t >> 3) * 256 + (t & 7) * 16);
                sts4(tileH + o, lh[0], lh[1], lh[2], lh[3]); sts4(tileH + o + 128, lh[4], lh[5], lh[6], lh[7]);
                sts4(tileL + o, ll[0], ll[1], ll[2], ll[3]); sts4(tileL + o + 128, ll[4], ll[5], ll[6], ll[7]);
                fence_async_smem();
                tc_fence_before();
                t1 = clock64(); if (rec) acc[5] += t1 - t0; t0 = t1;                  // 5: split + stores + fences
                named_bar(1 + g, 128);
                t1 = clock64(); if (rec) acc[6] += t1 - t0; t0 = t1;                  // 6: barrier B
                if (mode == 1) {
                    if (lane == 0 && q < 3) {   // three warps issue one MMA each and commit
                        tc_fence_after();
                        const uint32_t start = (uint32_t)((c0 + NB) >> 4) << 4;
                        const uint32_t idesc = IDESC_TF32_NEG_M128 | (((F - start) >> 3) << 17);
                        const uint64_t bH = descH + (uint64_t)(start * 2), bL = descL + (uint64_t)(start * 2);
                        umma_tf32(d_tmem + start, q == 2 ? descL : descH, q == 1 ? bL : bH, idesc, 1u);
                        tc_commit(bar);
                    }
                } else if (t == 0) {
                    tc_fence_after();
                    const uint32_t start = (uint32_t)((c0 + NB) >> 4) << 4;
                    const uint32_t idesc = IDESC_TF32_NEG_M128 | (((F - start) >> 3) << 17);
                    const uint64_t bH = descH + (uint64_t)(start * 2), bL = descL + (uint64_t)(start * 2);
                    umma_tf32(d_tmem + start, descH, bH, idesc, 1u);
                    if (mode != 2) umma_tf32(d_tmem + start, descH, bL, idesc, 1u);
                    if (mode != 2) umma_tf32(d_tmem + start, descL, bH, idesc, 1u);
                    if (mode != 3) tc_commit(bar);
                    else { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
                }
                t1 = clock64(); if (rec) acc[7] += t1 - t0; t0 = t1;                  // 7: MMA issue + commit
            }
        }
        // restore the matrix so that the factor stays well defined
        for (int c0 = 0; c0 < F; c0 += NB) { float a[NB]; for (int i = 0; i < NB; ++i) a[i] = (c0 + i == t) ? 300.0f : 0.01f * ((t * 7 + c0 + i) % 13); tmem_st8(t_row + c0, a); }
        tc_fence_before(); named_bar(1 + g, 128); tc_fence_after();
    }
    if (rec) for (int i = 0; i < 8; ++i) out[i] = acc[i];
    if (tid == 0) out[8] = (long long)__float_as_int(bt);
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory"); }
}
int main() {
    long long* d; cudaMalloc(&d, 128);
    const int smem = 4 * GROUP_BYTES + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[8] = {"mma_wait", "tmem_ld", "factor(own 4/16)", "barA", "P+rhs", "split+sts+fence", "barB", "mma_issue"};
    for (int mode = 0; mode < 4; ++mode)
    for (int ng = 1; ng <= 4; ng *= 4) {
        const int iters = 50;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            probe<<<1, 512, smem>>>(d, iters, ng, mode);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
            if (rep == 1) {
                printf("mode %d (0: 3 MMAs by one thread, 1: by three warps, 2: one MMA, 3: three MMAs without commit) groups %d: %.1f us per row\n  cycles per step:", mode, ng, ms * 1000 / iters);
                const int steps = iters * 16;
                for (int i = 0; i < 8; ++i) printf(" %s %.0f;", names[i], (double)h[i] / (i == 2 ? iters * 4 : steps));
                printf("\n");
            }
        }
    }
    return 0;
}
