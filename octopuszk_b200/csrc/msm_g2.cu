// G2 (Fq2) instantiation of the MSM kernels.
#include "common.h"
#include "msm_impl.cuh"
namespace ozk {
OZK_DEFINE_MSM_LAUNCH(Fq2, kMsmG2)
}  // namespace ozk
