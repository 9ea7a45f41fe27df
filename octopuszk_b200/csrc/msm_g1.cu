// G1 (Fq) instantiation of the MSM and fixed-base kernels.
#include "common.h"
#include "fixed_impl.cuh"
namespace ozk {
OZK_DEFINE_MSM_LAUNCH(Fq, kMsmG1)
OZK_DEFINE_FIXED_LAUNCH(Fq, kFixedG1)
}  // namespace ozk
