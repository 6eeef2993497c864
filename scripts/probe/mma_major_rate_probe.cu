// Probe: issue rate of tcgen05.mma kind::f16 M = N = 128, K = 16 with shared-memory operands in the K-major and in
// the MN-major 128-byte-swizzled layout (same tile for A and B, as the Gram kernels use it). One CTA, one issuing
// thread, NREP MMAs into one accumulator, cycles from clock64 around issue + commit + wait.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void probe(long long* out, int mode, int nrep, int n) {
    extern __shared__ uint8_t raw[];
    uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0u) : "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tm = tptr;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24) | (mode == 1 ? ((1u << 15) | (1u << 16)) : 0u);
        long long t0 = clock64();
        for (int r = 0; r < nrep; ++r) {
            // mode 0: K-major tile (128 rows x 64 fp16), K-step (r & 3); mode 1: MN-major tile, K-step (r & 1)
            uint64_t d = mode == 0 ? desc_k(base + (r & 3) * 32) : desc_mn(base + (r & 1) * 2048, 4096, 1024);
            uint64_t d2 = mode == 0 ? desc_k(base + 16384 + (r & 3) * 32) : desc_mn(base + 8192 + (r & 1) * 2048, 4096, 1024);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm + (r & 1) * 128), "l"(d), "l"(d2), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE;\n\tbra W;\n\tDONE:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        out[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tm) : "memory");
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    for (int n : {128, 96, 64})
        for (int mode = 0; mode < 2; ++mode)
            for (int nrep : {256, 2048}) {
                probe<<<1, 128, 70000>>>(d, mode, nrep, n);
                cudaError_t e = cudaDeviceSynchronize();
                long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
                printf("N=%d %s nrep %d: %s %lld cycles = %.1f per MMA\n", n, mode ? "MN-major" : "K-major ", nrep, cudaGetErrorString(e), c, (double)c / nrep);
            }
    return 0;
}
