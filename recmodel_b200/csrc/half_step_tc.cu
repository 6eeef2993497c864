// placeholder until the tcgen05 kernel lands (next commit): reports "unsupported" so AUTO
// resolves to the SIMT path.
#include "half_step.cuh"
namespace wmf {
bool tc_half_step_supported(int, int) { return false; }
size_t tc_half_step_workspace_bytes(int64_t, int, int) { return 0; }
int tc_half_step(const HalfStepParams&, void*, size_t, cudaStream_t) {
    set_error("tcgen05 half-step not built");
    return WMF_ERR_UNSUPPORTED;
}
}  // namespace wmf
