// deposit_tiled.cu -- shared-memory tile accumulation path of psc_deposit (placeholder: falls back).
#include "common.cuh"
namespace psc {
int deposit_tiled(const float *, int64_t, int, int, float, float, float, float *, cudaStream_t) { return 0; }
}  // namespace psc
