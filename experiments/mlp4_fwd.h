// Entry points of the single-CTA forward kernel (mlp3_fwd.cu) used by the dispatcher in mlp4_fwd.cu.
#pragma once
#include <cuda_runtime.h>

namespace nerfq {
int mlp3_forward_launch(const void* packed, const float* rays, const float* z, long long n_rays, int samples_per_ray, float* raw, void* save,
                        int max_ctas, cudaStream_t stream);
bool mlp3_forward_tracing();      // nerfq_mlp_set_trace installed a trace buffer
// CTA-pair, points-on-lanes forward for rendering (mlp5_fwd.cu); no saved activations
int mlp5_forward_launch(const void* packed, const float* rays, const float* z, long long n_rays, int samples_per_ray, float* raw,
                        int max_ctas, cudaStream_t stream);
}  // namespace nerfq
