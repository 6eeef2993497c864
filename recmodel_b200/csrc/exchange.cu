// Factor / Gram-block exchange of the row-sharded half-step over peer memory (NVLink / NVSwitch).
//
// After a half-step rank g holds the new factor rows of its shard and the double-precision Gram partials of the
// blocks inside it. Every peer needs both before its next half-step. Instead of an NCCL all-gather
// (factors) and an all-reduce of zero-padded block partials, the owner WRITES its rows and blocks straight into
// every peer's copy of the full buffers: the buffers are symmetric allocations whose peer-mapped base addresses
// the caller passes in a device table, the stores travel over NVLink as posted writes, and one cross-rank barrier
// per half-step (caller's, e.g. the signal-pad barrier of torch's symmetric memory) orders them before the readers.
// Each rank then adds ALL blocks in block order (wmf_gram_reduce): the bits of the single-GPU Gram.
#include "common.cuh"

namespace wmf {

template <typename V>
__global__ void __launch_bounds__(256)
peer_broadcast_kernel(const V* __restrict__ src, size_t nvec, void* const* __restrict__ peers, int world, int self,
                      size_t dst_offset_bytes) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const V v = src[i];
        for (int r = 0; r < world; ++r) {
            if (r == self) continue;   // the source already lives in this rank's copy
            V* dst = reinterpret_cast<V*>(reinterpret_cast<char*>(peers[r]) + dst_offset_bytes);
            dst[i] = v;
        }
    }
}

}  // namespace wmf

using namespace wmf;

extern "C" int wmf_peer_broadcast(const void* src, size_t bytes, const void* const* peer_bases_dev, int world, int self,
                                  size_t dst_offset_bytes, void* stream) {
    WMF_REQUIRE(src && peer_bases_dev && world >= 1 && self >= -1 && self < world, "wmf_peer_broadcast: bad arguments");
    if (bytes == 0 || world == 1) return WMF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    void* const* peers = const_cast<void* const*>(reinterpret_cast<const void* const*>(peer_bases_dev));
    const bool v16 = ((reinterpret_cast<uintptr_t>(src) | dst_offset_bytes | bytes) & 15) == 0;
    const size_t nvec = v16 ? bytes / 16 : bytes / 4;
    WMF_REQUIRE(v16 || (((reinterpret_cast<uintptr_t>(src) | dst_offset_bytes | bytes) & 3) == 0),
                "wmf_peer_broadcast: buffers must be 4-byte aligned");
    size_t blocks = (nvec + 255) / 256;
    const size_t cap = (size_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (v16) peer_broadcast_kernel<uint4><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), nvec, peers, world, self, dst_offset_bytes);
    else peer_broadcast_kernel<uint32_t><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(src), nvec, peers, world, self, dst_offset_bytes);
    WMF_LAUNCH_CHECK("peer_broadcast_kernel");
    return WMF_OK;
}
