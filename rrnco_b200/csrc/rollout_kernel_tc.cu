// Second translation unit of the fused rollout kernel: instantiates the tcgen05-FFN variants
// (rollout_kernel<..., kTc = true>) so that they compile in parallel with the mma.sync variants.
#define RRNCO_BUILD_TC 1
#include "rollout_kernel.cu"
