// joint_at.cu -- fused joint + log-softmax + gradient, ADD_TANH mode (placeholder launchers;
// replaced by the SIMT fp32 and tcgen05 bf16 kernels).
#include "common.cuh"

namespace rnntb200 {

int launch_at_lse(const float*, const float*, const float*, const float*, int, const int32_t*,
                  const int32_t*, const int32_t*, int, int, int, int, int, int, float2*, float*,
                  cudaStream_t) {
    return RNNTB200_STATUS_EXECUTION_FAILED;
}
int launch_at_grad(const float*, const float*, const float*, const float*, int, const int32_t*,
                   const int32_t*, const int32_t*, int, int, int, int, int, int, const float*,
                   const float*, const float*, const float*, const float*, float*, float*, float*,
                   float*, cudaStream_t) {
    return RNNTB200_STATUS_EXECUTION_FAILED;
}

}  // namespace rnntb200
