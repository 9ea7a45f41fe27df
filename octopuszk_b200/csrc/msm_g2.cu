// G2 (Fq2) instantiation of the MSM and fixed-base kernels.
#include "common.h"
#include "fixed_impl.cuh"
namespace ozk {
OZK_DEFINE_MSM_LAUNCH(Fq2, kMsmG2)
OZK_DEFINE_FIXED_LAUNCH(Fq2, kFixedG2)
}  // namespace ozk
