// Probe: tcgen05.ld (TMEM -> registers) throughput per SM, as a function of how many warps read at once and of the
// load width (decides whether a solver may re-read its system matrix from tensor memory every iteration:
// half_step_cg.cuh). One CTA; warp w reads lane quarter w % 4. Also: the same loop with 32 FMAs per 32 loaded
// values against a vector in shared memory (the matrix-vector product of the conjugate-gradient solver).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ float ld_sum(uint32_t addr) {
    float s = 0.f;
    if constexpr (X == 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) s += __uint_as_float(r[i]);
    } else {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) s += __uint_as_float(r[i]);
    }
    return s;
}

// mode 0: loads + adds only; mode 1: loads + FMA against a shared vector (matvec)
template <int X, int MODE>
__global__ void __launch_bounds__(1024, 1) probe(float* out, long long* cyc, int iters, int nwarps) {
    __shared__ uint32_t tmem_ptr;
    __shared__ __align__(16) float vec[512];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) vec[i] = 1.0f + i * 1e-3f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_ptr;
    // zero the columns this warp will read (so the sums are finite)
    if (warp < 4) {
        for (int c = 0; c < 512; c += 8) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(base + ((uint32_t)(warp * 32) << 16) + c), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float s = 0.f;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        const uint32_t row = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128 % 512);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 128; c += X) {
                if (MODE == 0) s += ld_sum<X>(row + c);
                else {
                    // matvec piece: X loaded values times X shared values
                    uint32_t r[X];
                    if constexpr (X == 32) {
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                            : "r"(row + c));
                    } else {
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                            : "r"(row + c));
                    }
                    float4 v[X / 4];
#pragma unroll
                    for (int i = 0; i < X / 4; ++i) v[i] = *reinterpret_cast<const float4*>(vec + c + 4 * i);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                    for (int i = 0; i < X / 4; ++i) {
                        a0 = fmaf(__uint_as_float(r[4 * i]), v[i].x, a0);
                        a1 = fmaf(__uint_as_float(r[4 * i + 1]), v[i].y, a1);
                        a2 = fmaf(__uint_as_float(r[4 * i + 2]), v[i].z, a2);
                        a3 = fmaf(__uint_as_float(r[4 * i + 3]), v[i].w, a3);
                    }
                    s += (a0 + a1) + (a2 + a3);
                }
            }
        }
        t1 = clock64();
    }
    if (lane == 0 && warp < nwarps) cyc[warp] = t1 - t0;
    out[threadIdx.x] = s;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
    }
}

template <int X, int MODE>
void run(float* out, long long* cyc, int nwarps) {
    const int iters = 2000;
    probe<X, MODE><<<1, 1024>>>(out, cyc, iters, nwarps);
    long long h[32];
    cudaError_t e = cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long mx = 0;
    for (int w = 0; w < nwarps; ++w) mx = h[w] > mx ? h[w] : mx;
    const double per_pass = (double)mx / iters;           // cycles for every warp to read 32 lanes x 128 columns
    const double bytes = (double)nwarps * 32 * 128 * 4;
    printf("x%-2d %s warps %2d: %.0f cycles per 128-column pass per warp, %.0f B/cycle/SM\n", X, MODE ? "matvec" : "ld+add",
           nwarps, per_pass, bytes / per_pass);
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 256);
    for (int nw : {1, 4, 8, 16}) {
        run<32, 0>(out, cyc, nw);
        run<16, 0>(out, cyc, nw);
        run<32, 1>(out, cyc, nw);
        run<16, 1>(out, cyc, nw);
    }
    return 0;
}
