// Parameter block and internal entry points shared by the half-step implementations.
#pragma once
#include "common.cuh"

namespace wmf {

struct HalfStepParams {
    const int64_t* indptr;
    const int32_t* indices;
    const float* data;
    int64_t rows;
    const int32_t* row_order;
    int64_t sched_len;  // schedule slots to walk: order_len, or rows without a schedule
    const float* Y;
    int64_t ldy;
    int f;
    const float* G;
    int bias;
    float* X;
    int64_t ldx;
    // filled by the implementation
    int* counter;
    long long* prof;    // nullable: per-role cycle counters of CTA 0 (tcgen05 kernel, debugging)
    const int* run_if;  // nullable: the SIMT kernel returns at once when *run_if == 0
    float* slab;
    int lda;
    int FP;
    int KC;
};

size_t simt_half_step_workspace_bytes(int f);
int simt_half_step(const HalfStepParams& in, void* ws, size_t ws_bytes, cudaStream_t st);
bool tc_half_step_supported(int f, int bias);
size_t tc_half_step_workspace_bytes(int64_t rows, int f, int bias, int64_t segments);  // segments < 0: default scratch
int tc_half_step(const HalfStepParams& in, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace wmf
