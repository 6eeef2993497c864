// Probe: do warp shuffles share the shared-memory data pipe with LDS? (Decides whether the conjugate-gradient
// solver and the operand transform of the half-step kernels should broadcast per-step values with LDS or SHFL:
// both kernels are bound by shared-memory wavefronts.) One CTA of 16 warps on one SM:
//   A: every warp issues N broadcast LDS.128        B: every warp issues N SHFL.IDX
//   C: warps 0-7 LDS.128, warps 8-15 SHFL (N each)  D: 4 N broadcast LDS.32 per warp (same bytes as A)
// If C takes ~max(A/2, B/2) the two use different pipes; if ~(A + B)/2 they share one.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) probe(int mode, int n, float* out, long long* cyc) {
    __shared__ __align__(16) float vec[1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) vec[i] = 1.0f + i;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(vec);
    float acc = 0.f, mine = lane * 0.5f;
    const bool do_lds = mode == 0 || mode == 3 || (mode == 2 && warp < 8);
    const long long t0 = clock64();
    if (do_lds && mode != 3) {
        for (int i = 0; i < n; ++i) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + (uint32_t)((i & 63) * 16)));
            acc += v.x + v.y + v.z + v.w;
        }
    } else if (mode == 3) {
        for (int i = 0; i < 4 * n; ++i) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (uint32_t)((i & 255) * 4)));
            acc += v;
        }
    } else {
        for (int i = 0; i < n; ++i) {
            float v;
            asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=f"(v) : "f"(mine), "r"(i & 31));
            acc += v;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    const long long t2 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    if (threadIdx.x == 0) cyc[16] = t2 - t0;
    out[threadIdx.x] = acc;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 256);
    const int n = 20000;
    const char* names[4] = {"A: 16 warps x N LDS.128 broadcast", "B: 16 warps x N SHFL.IDX", "C: 8 warps LDS.128 + 8 warps SHFL",
                            "D: 16 warps x 4N LDS.32 broadcast"};
    for (int mode = 0; mode < 4; ++mode) {
        probe<<<1, 512>>>(mode, n, out, cyc);
        long long h[17];
        if (cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
        printf("%-40s CTA %.0f cycles = %.2f cycles per warp-instruction of the busiest kind\n", names[mode], (double)h[16],
               (double)h[16] / (mode == 2 ? 8.0 * n : (mode == 3 ? 64.0 * n : 16.0 * n)));
        if (mode == 2) printf("    LDS warps finished after %lld cycles, SHFL warps after %lld\n", h[0], h[8]);
    }
    return 0;
}
