// joint_at_tc.cu -- tcgen05 (5th-gen tensor core) forward of the ADD_TANH joint.  Placeholder:
// reports "unsupported" so RNNTB200_GEMM_BF16 runs the bf16-emulating CUDA-core kernels.
#include "common.cuh"

namespace rnntb200 {

bool at_tc_supported(int, int) { return false; }

int launch_at_lse_tc(const float*, const float*, const float*, const float*, const int32_t*,
                     const int32_t*, const int32_t*, int, int, int, int, int, int, float2*, float*,
                     cudaStream_t) {
    return RNNTB200_STATUS_EXECUTION_FAILED;
}

}  // namespace rnntb200
