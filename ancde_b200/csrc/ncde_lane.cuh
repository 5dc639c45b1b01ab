// The lane engine (ANCDE_ENGINE_LANE): one WARP integrates one sample.  Host-side interface of ncde_lane.cu.
#pragma once
#include "ncde_common.cuh"

namespace ancde {

// Pure shape predicate (no layout needed): can the lane engine run this descriptor at all?
bool lane_shape_ok(const ancde_ncde_desc* d);
// Forward / recurrent backward for a TS = 1 layout (a.lo filled by make_layout, weights packed by pack_weights).
int launch_lane_fwd(SolveArgs a, cudaStream_t stream);
int launch_lane_bwd(SolveArgs a, cudaStream_t stream);

}  // namespace ancde
