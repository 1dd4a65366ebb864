// pp_kernels_msc3d.cuh — specialised multispin kernel for 3-D hypercubic lattices (placeholder until
// the generic kernel's parity is confirmed on hardware).
#pragma once
#include "pp_device.cuh"
#include "pp_plan.h"
#include "../../include/peapods_b200.h"

namespace pp {
inline bool msc3d_supported(const LatticePlan &) { return false; }
inline pp_status launch_msc3d(const ModelView &, cudaStream_t, uint32_t, int, bool, bool, int64_t, int64_t *) {
    return PP_ERR_UNSUPPORTED;
}
}  // namespace pp
