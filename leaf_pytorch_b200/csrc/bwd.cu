// placeholder: backward kernels land in a later commit
#include "../../include/leafk.h"
#include "leafk_common.cuh"
namespace leafk {
int fail(int code, const char* fmt, ...);
size_t bwd_workspace_bytes(const leafk_config*, int, int) { return 0; }
int bwd_run(const leafk_config*, const leafk_params*, const float*, int, int, const float*, const float*,
            const leafk_grads*, float*, void*, size_t, cudaStream_t) {
  return fail(LEAFK_EINVAL, "backward not implemented yet");
}
}
