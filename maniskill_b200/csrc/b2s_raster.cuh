// b200sim batched tiled rasteriser (device).  Placeholder interface until the rasteriser lands (next milestone).
#pragma once
#include <cuda_runtime.h>

#include "../../include/b200sim.h"
#include "b2s_step.cuh"

namespace b2s {
struct RasterGroup {
  int dummy;
};
inline const char* raster_create(const DevModel&, const DevState&, const B2SModel&, const B2SCameraDesc*, int, const B2SVisualTable*,
                                 RasterGroup**, B2SRenderTargets*) {
  return "rasteriser not built into this library yet";
}
inline const char* raster_run(const DevModel&, const DevState&, RasterGroup*, cudaStream_t) { return "rasteriser not built"; }
inline void raster_destroy(RasterGroup*) {}
}  // namespace b2s
