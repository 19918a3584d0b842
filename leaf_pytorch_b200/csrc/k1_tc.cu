// placeholder: tensor-core kernel lands in the next commit
#include "leafk_common.cuh"
namespace leafk {
bool k1_tc_supported(const Geom&, const char** why) { if (why) *why = "not built"; return false; }
cudaError_t launch_k1_tc(const Geom&, const float*, const uint8_t*, const float*, float*, int, int, cudaStream_t) {
  return cudaErrorNotSupported;
}
}
