// G1 (Fq) instantiation of the batch-affine pre-reduction kernels (msm_ba_impl.cuh).
#include "common.h"
#include "msm_ba_impl.cuh"
namespace ozk {
OZK_DEFINE_MSM_BA_LAUNCH(Fq, kMsmBaG1)
}  // namespace ozk
