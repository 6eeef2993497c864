// Probe: cp.async.bulk.tensor.2d ... tile::gather4 (TMA row gather) on sm_100a. Decides how the half-step kernel
// stages gathered factor rows (half_step_tc.cu): which box the tensor map needs, where the four rows land in
// shared memory, how many bytes complete_tx counts, and what an out-of-range row index does.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_gather4_probe tma_gather4_probe.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, const int* idx, float* out, int width, int nrows_smem,
                      unsigned expect_bytes, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x;
    float* stg = reinterpret_cast<float*>(smem);
    for (int i = lane; i < nrows_smem * width; i += 32) stg[i] = -7.0f;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const long long t0 = clock64();
    if (lane == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(expect_bytes) : "memory");
    __syncwarp();
    if (lane < 8) {
        const int i0 = idx[4 * lane], i1 = idx[4 * lane + 1], i2 = idx[4 * lane + 2], i3 = idx[4 * lane + 3];
        const uint32_t dst = smem_u32(stg) + (uint32_t)lane * 4u * (uint32_t)width * 4u;
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bar)),
            "r"(0), "r"(i0), "r"(i1), "r"(i2), "r"(i3) : "memory");
    }
    // wait (bounded)
    uint32_t done = 0;
    for (int spin = 0; spin < 2000000 && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const long long t1 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = done; }
    __syncwarp();
    for (int i = lane; i < nrows_smem * width; i += 32) out[i] = stg[i];
}

int main() {
    const int R = 1000, W = 128;
    std::vector<float> h((size_t)R * W);
    for (int r = 0; r < R; ++r) for (int c = 0; c < W; ++c) h[(size_t)r * W + c] = r * 1000.0f + c;
    float *Y, *out; int* idx; long long* cyc;
    cudaMalloc(&Y, h.size() * 4); cudaMemcpy(Y, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 40 * W * 4); cudaMalloc(&idx, 32 * 4); cudaMalloc(&cyc, 16);
    int hidx[32];
    for (int i = 0; i < 32; ++i) hidx[i] = (i * 37 + 11) % R;
    hidx[5] = R + 3;   // out of range
    hidx[30] = 0;
    cudaMemcpy(idx, hidx, sizeof(hidx), cudaMemcpyHostToDevice);
    for (int boxrows : {1, 4}) {
        for (unsigned expect : {32u * W * 4u, 31u * W * 4u}) {
            CUtensorMap tmap;
            cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)R};
            cuuint64_t gstride[1] = {(cuuint64_t)W * 4};
            cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)boxrows};
            cuuint32_t estr[2] = {1, 1};
            CUresult rc = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Y, gdim, gstride, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            printf("box rows %d expect_tx %u: encode rc=%d\n", boxrows, expect, (int)rc);
            if (rc != CUDA_SUCCESS) continue;
            cudaMemset(out, 0, 40 * W * 4);
            probe<<<1, 32, 40 * W * 4>>>(tmap, idx, out, W, 36, expect, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<float> o(36 * W); long long hc[2];
            cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
            int good = 0;
            for (int e2 = 0; e2 < 32; ++e2) {
                bool ok = true;
                for (int c = 0; c < W; ++c) {
                    const float want = hidx[e2] < R ? hidx[e2] * 1000.0f + c : 0.0f;
                    if (o[(size_t)e2 * W + c] != want) ok = false;
                }
                good += ok;
            }
            printf("  barrier completed=%lld after %lld cycles; rows as expected (row e at e*%d floats, OOB row = zeros): %d of 32\n",
                   hc[1], hc[0], W, good);
            printf("  row 0: %.0f %.0f ... row 5 (OOB): %.0f %.0f  row 31: %.0f  rows 32..35 (untouched = -7): %.0f %.0f\n", o[0], o[1],
                   o[5 * W], o[5 * W + 1], o[31 * W], o[32 * W], o[35 * W]);
        }
    }
    return 0;
}
