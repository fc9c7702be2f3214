// field_tc.cu -- tcgen05 / TMEM implementation of the LSEField MLPs (3xTF32 split for fp32 parity).
// Placeholder until the tensor-core kernels land: requesting use_tensor_cores fails loudly (no silent fallback).
#include "common.cuh"
#include "../../include/lsenerf_b200.h"

namespace lse {

int tc_density_fwd(const LseFieldParams &, const float *, const float *, const int32_t *, const float *, const float *,
                   int64_t, const int32_t *, float *, float *, cudaStream_t)
{
    set_error("lse_field_density_fwd: use_tensor_cores=1 requested but the tcgen05 kernels are not built into this library");
    return -38;
}

int tc_color_fwd(const LseFieldParams &, const float *, const int32_t *, const int32_t *, int64_t, const int32_t *,
                 const float *, float *, cudaStream_t)
{
    set_error("lse_field_color_fwd: use_tensor_cores=1 requested but the tcgen05 kernels are not built into this library");
    return -38;
}

}  // namespace lse
