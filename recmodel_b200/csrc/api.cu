// Error state, device check and the small element-wise entry points of libwmf_b200.
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace wmf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return WMF_OK;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
        set_error("no CUDA device: %s (%s)", cudaGetErrorString(e), what);
        return WMF_ERR_NO_DEVICE;
    }
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return WMF_ERR_CUDA;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long& g_launches_ref() { return g_launches; }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

__global__ void preprocess_kernel(float* __restrict__ d, int64_t n, int mode, float alpha, float beta) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float x = d[i];
        // alpha*log(1+beta*x): same operation order as wmf_model.py:120 (fp32 throughout)
        d[i] = mode == WMF_PREPROCESS_LOG ? __fmul_rn(alpha, logf(__fadd_rn(1.0f, __fmul_rn(beta, x))))
                                          : __fmul_rn(alpha, x);
    }
}

}  // namespace wmf

using namespace wmf;

extern "C" {

const char* wmf_last_error(void) { return g_err; }

int wmf_version(void) { return 200; }

long long wmf_launch_count(void) { return (long long)__atomic_load_n(&wmf::g_launches_ref(), __ATOMIC_RELAXED); }

int wmf_device_check(int* sms) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible (%s)", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        return WMF_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0;
    WMF_CUDA(cudaGetDevice(&dev));
    WMF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("device %d has compute capability %d.x; libwmf_b200 is built for sm_100a only", dev, major);
        return WMF_ERR_NO_DEVICE;
    }
    if (sms) *sms = sm_count();
    return WMF_OK;
}

int wmf_preprocess(float* data, int64_t nnz, int mode, float alpha, float beta, void* stream) {
    WMF_REQUIRE(mode == WMF_PREPROCESS_LOG || mode == WMF_PREPROCESS_LINEAR, "wmf_preprocess: unknown mode %d", mode);
    WMF_REQUIRE(nnz >= 0 && (data != nullptr || nnz == 0), "wmf_preprocess: null data");
    if (nnz == 0) return WMF_OK;
    int grid = (int)((nnz + 1023) / 1024);
    int cap = sm_count() * 16;
    if (grid > cap) grid = cap;
    preprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(data, nnz, mode, alpha, beta);
    WMF_LAUNCH_CHECK("preprocess_kernel");
    return WMF_OK;
}

}  // extern "C"
