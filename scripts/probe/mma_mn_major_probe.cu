// Probe: D[128x128] = Z^T Z for Z a (K entries) x (128 features) fp16 matrix stored ENTRY-major, i.e. each entry's
// features contiguous: the MN-major canonical layout of tcgen05 (cute: Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units). Entry j, feature block fb (64 features), 16-byte chunk c (8 features) lives at
//     fb * LBO + (j / 8) * SBO + (j % 8) * 128 + ((c ^ (j % 8)) * 16)
// A and B both use this tile with a_major = b_major = MN (instruction descriptor bits 15 / 16). One MMA consumes
// K = 16 entries = two 8-row groups, so K-step k starts at byte offset 2 * k * SBO.
// Variants: (LBO, SBO) as expected, swapped, and the 32-entry K extent in one or two MMAs.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
constexpr uint32_t IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr int KE = 32;  // entries
__global__ void probe(const float* Z, float* D, uint32_t lbo, uint32_t sbo, uint32_t desc_lbo, uint32_t desc_sbo, int nk) {
    extern __shared__ uint8_t raw[];
    uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 32768 / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0u) : "memory");
    __syncthreads();
    // warp w writes entries w, w+4, ...: lane l holds features 4l..4l+3 (what the gather warps of the kernels do)
    for (int j = warp; j < KE; j += 4) {
        const float* z = Z + j * 128 + lane * 4;
        __half2 h01 = __floats2half2_rn(z[0], z[1]), h23 = __floats2half2_rn(z[2], z[3]);
        uint32_t fb = lane >> 4, c = (lane & 15) >> 1, half8 = (lane & 1) * 8;
        uint32_t a = base + fb * lbo + (j >> 3) * sbo + (j & 7) * 128 + ((c ^ (j & 7)) << 4) + half8;
        asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(a), "r"(*reinterpret_cast<uint32_t*>(&h01)), "r"(*reinterpret_cast<uint32_t*>(&h23)) : "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tm = tptr;
    if (tid == 0) {
        for (int k = 0; k < nk; ++k) {
            uint64_t dk = umma_desc_mn(base + 2 * k * sbo, desc_lbo, desc_sbo);
            uint32_t acc = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(dk), "l"(dk), "r"(IDESC), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE;\n\tbra W;\n\tDONE:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        int q = warp & 3, t = q * 32 + lane;
        for (int c0 = 0; c0 < 128; c0 += 8) {
            uint32_t r[8];
            uint32_t ta = tm + ((uint32_t)(q * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 8; ++i) D[t * 128 + c0 + i] = __uint_as_float(r[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tm) : "memory");
}
int main() {
    std::vector<float> Z(KE * 128), D(128 * 128);
    for (auto& v : Z) v = (float)(rand() % 17 - 8) / 4.0f;  // exactly representable in fp16
    float *dZ, *dD;
    cudaMalloc(&dZ, Z.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    struct V { uint32_t lbo, sbo, dl, ds; const char* name; } vs[] = {
        {8192, 1024, 8192, 1024, "layout fb*8192 + kgroup*1024, desc LBO=8192 SBO=1024 (expected)"},
        {8192, 1024, 1024, 8192, "same layout, desc fields swapped"},
        {1024, 2048, 1024, 2048, "layout fb*1024 + kgroup*2048 (cute tile_to_shape order), desc LBO=1024 SBO=2048"},
    };
    for (auto& v : vs)
        for (int nk = 1; nk <= 2; ++nk) {
            cudaMemset(dD, 0, D.size() * 4);
            probe<<<1, 128, 40000>>>(dZ, dD, v.lbo, v.sbo, v.dl, v.ds, nk);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 128; ++j) {
                    double ref = 0;
                    for (int k = 0; k < nk * 16; ++k) ref += (double)Z[k * 128 + i] * Z[k * 128 + j];
                    maxerr = fmax(maxerr, fabs(ref - D[i * 128 + j]));
                }
            printf("%s, nk %d: %s max abs err %.3g  D[0][0]=%g D[5][70]=%g\n", v.name, nk, cudaGetErrorString(e), maxerr, D[0], D[5 * 128 + 70]);
        }
    return 0;
}
