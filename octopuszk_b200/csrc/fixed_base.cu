// placeholder until the fixed-base kernels land (see msm.cu for the shared pipeline)
#include "common.h"
namespace ozk {
void fixed_free_tables(ozk_ctx*) {}
}  // namespace ozk
