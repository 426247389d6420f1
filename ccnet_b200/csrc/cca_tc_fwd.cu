// tcgen05 forward kernels -- placeholder until the tensor-core path lands.
#include "cca_common.cuh"
namespace cca {
bool tc_forward_supported(Dims, int) { return false; }
cudaError_t tc_forward(const void *, const void *, const void *, void *, float *, void *, Dims, int, cudaStream_t,
                       const char **why)
{
    if (why) *why = "tcgen05 forward not built";
    return cudaErrorNotSupported;
}
}  // namespace cca
