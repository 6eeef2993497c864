// Shared helpers for libwmf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/wmf_b200.h"

namespace wmf {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();
void count_launch();   // process-wide count of kernel launches issued by this library (wmf_launch_count)
unsigned long long& g_launches_ref();

#define WMF_CUDA(call)                                        \
    do {                                                      \
        int _rc = ::wmf::check_cuda((call), #call);           \
        if (_rc) return _rc;                                  \
    } while (0)

#define WMF_LAUNCH_CHECK(name)                                \
    do {                                                      \
        ::wmf::count_launch();                                \
        int _rc = ::wmf::check_cuda(cudaGetLastError(), name);\
        if (_rc) return _rc;                                  \
    } while (0)

#define WMF_REQUIRE(cond, ...)                                \
    do {                                                      \
        if (!(cond)) {                                        \
            ::wmf::set_error(__VA_ARGS__);                    \
            return WMF_ERR_INVALID;                           \
        }                                                     \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------
// NumPy pairwise add.reduce order for one dot product, computed by a group of 8 lanes.
// Restates wmf_model.py:206 `(u * v).sum(axis=1)`: p_i = fl(u_i * v_i) (no FMA), then
// n < 8: sequential; n <= 128: accumulator k (= lane k of the group) sums p_{8b+k} over the
// full blocks, combine ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the scalar tail;
// n > 128: split at n/2 rounded down to a multiple of 8 and add the halves.
// All 8 lanes of the group return the same value. `gl` = lane index within the group (0..7),
// `gmask` = the 8-lane mask of this group inside the warp.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float np_block_sum(const float* __restrict__ u, const float* __restrict__ v,
                                              int n, int gl, unsigned gmask) {
    if (n < 8) {
        float acc = n > 0 ? __fmul_rn(u[0], v[0]) : 0.0f;  // np.sum of empty = 0
        for (int i = 1; i < n; ++i) acc = __fadd_rn(acc, __fmul_rn(u[i], v[i]));
        return acc;
    }
    float r = __fmul_rn(u[gl], v[gl]);
    int i = 8;
    for (; i + 8 <= n; i += 8) r = __fadd_rn(r, __fmul_rn(u[i + gl], v[i + gl]));
    r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 4));
    for (; i < n; ++i) r = __fadd_rn(r, __fmul_rn(u[i], v[i]));
    return r;
}

__device__ inline float np_pairwise_dot(const float* __restrict__ u, const float* __restrict__ v, int n,
                                        int gl, unsigned gmask) {
    if (n <= 128) return np_block_sum(u, v, n, gl, gmask);
    int half = n / 2;
    half -= half % 8;
    // WMF_MAX_F = 320 < 512: at most two levels, so unroll the recursion by hand.
    float a, b;
    if (half <= 128) a = np_block_sum(u, v, half, gl, gmask);
    else {
        int h2 = half / 2; h2 -= h2 % 8;
        a = __fadd_rn(np_block_sum(u, v, h2, gl, gmask), np_block_sum(u + h2, v + h2, half - h2, gl, gmask));
    }
    int rest = n - half;
    if (rest <= 128) b = np_block_sum(u + half, v + half, rest, gl, gmask);
    else {
        int h2 = rest / 2; h2 -= h2 % 8;
        b = __fadd_rn(np_block_sum(u + half, v + half, h2, gl, gmask),
                      np_block_sum(u + half + h2, v + half + h2, rest - h2, gl, gmask));
    }
    return __fadd_rn(a, b);
}

// score of (user row u, item row v) with optional biases in column 0 (wmf_model.py:209-211):
// latent sum, then + user bias, then + item bias.
__device__ __forceinline__ float np_score(const float* __restrict__ u, const float* __restrict__ v, int f,
                                          int bias, int gl, unsigned gmask) {
    if (!bias) return np_pairwise_dot(u, v, f, gl, gmask);
    float s = np_pairwise_dot(u + 1, v + 1, f - 1, gl, gmask);
    return __fadd_rn(__fadd_rn(s, u[0]), v[0]);
}

}  // namespace wmf
