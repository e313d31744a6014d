// interp_tiled.cu -- shared-memory staged path of psc_interp_kick (placeholder: falls back).
#include "common.cuh"
namespace psc {
int interp_kick_tiled(const float *, const float *, float *, float *, int64_t, int, int, float, float *,
                      cudaStream_t) { return 0; }
}  // namespace psc
