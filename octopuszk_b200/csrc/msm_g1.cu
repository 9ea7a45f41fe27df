// G1 (Fq) instantiation of the MSM kernels.
#include "common.h"
#include "msm_impl.cuh"
namespace ozk {
OZK_DEFINE_MSM_LAUNCH(Fq, kMsmG1)
}  // namespace ozk
