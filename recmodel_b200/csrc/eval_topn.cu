// N2: the counting step of the sampled Recall@N protocol, RecModel.compute_hit
// (base_model.py:51-98), for all held-out interactions at once.
//
// The reference draws, per user, rand_sampled+1 candidate ids and one slot, and for every held-out item
// of that user overwrites the slot, ranks the list and tests `item in top[:k]`. Only the slot changes
// between a user's rankings, so the scores of the list are computed once per user (wmf_predict_pairs,
// bit-exact) and a held-out item's place in the ranking is a count: the candidates that WMF.rank would
// put ahead of it (higher score, or equal score at a lower position; rank breaks ties by position).
// `item in top[:k]` also matches a second copy of the item id elsewhere in the list, whose score is the
// same, so the position that counts is the lowest one holding the item id.
#include "common.cuh"

namespace wmf {

// one warp per held-out interaction p: ahead[p] = candidates ranked ahead of the item
__global__ void rank_ahead_kernel(const float* __restrict__ S, const int32_t* __restrict__ cand,
                                  const int32_t* __restrict__ slot, int64_t L, const int32_t* __restrict__ pair_user,
                                  const int32_t* __restrict__ pair_item, const float* __restrict__ pair_score,
                                  int64_t np, int32_t* __restrict__ ahead) {
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= np) return;
    const int64_t u = pair_user[p];
    const int32_t item = pair_item[p];
    const float s = pair_score[p];
    const int32_t sl = slot[u];
    const float* Su = S + u * L;
    const int32_t* cu = cand + u * L;
    // lowest position that holds the item id (the slot does, by construction)
    int32_t pmin = sl;
    for (int64_t q = lane; q < L; q += 32)
        if (q != sl && cu[q] == item && (int32_t)q < pmin) pmin = (int32_t)q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pmin = min(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
    int cnt = 0;
    for (int64_t q = lane; q < L; q += 32) {
        if (q == sl || cu[q] == item) continue;  // the item itself (any copy) is never ahead of itself
        const float sq = Su[q];
        cnt += (sq > s || (sq == s && (int32_t)q < pmin)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) ahead[p] = cnt;
}

}  // namespace wmf

using namespace wmf;

extern "C" int wmf_rank_ahead(const float* S, const int32_t* cand, const int32_t* slot, int64_t nu, int64_t L,
                              const int32_t* pair_user, const int32_t* pair_item, const float* pair_score, int64_t np,
                              int32_t* ahead, void* stream) {
    WMF_REQUIRE(nu >= 0 && L > 0 && np >= 0, "wmf_rank_ahead: bad sizes");
    if (np == 0) return WMF_OK;
    WMF_REQUIRE(S && cand && slot && pair_user && pair_item && pair_score && ahead, "wmf_rank_ahead: null pointer");
    const int threads = 256;
    const int64_t blocks = (np * 32 + threads - 1) / threads;
    WMF_REQUIRE(blocks < (1ll << 31), "wmf_rank_ahead: too many interactions in one call (%lld)", (long long)np);
    rank_ahead_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(S, cand, slot, L, pair_user, pair_item,
                                                                             pair_score, np, ahead);
    WMF_LAUNCH_CHECK("rank_ahead_kernel");
    return WMF_OK;
}
