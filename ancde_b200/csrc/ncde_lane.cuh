// The lane engine (ANCDE_ENGINE_LANE): one small CTA integrates one sample.  Host-side interface of ncde_lane.cu.
#pragma once
#include "ncde_common.cuh"

namespace ancde {

// Pure shape predicate (no layout needed): can the lane engine run this descriptor at all?
bool lane_shape_ok(const ancde_ncde_desc* d);
// Samples per tile of the saved layout when the lane engine runs this descriptor (1, or 8 for large batches), 0 when it does not
// (the automatic choice for engine = 0, the forced one for engine = ANCDE_ENGINE_LANE).
int lane_choice(const ancde_ncde_desc* d);
// Forward / recurrent backward for a TS = 1 or TS = 8 layout (a.lo filled by make_layout, weights packed by pack_weights).
int launch_lane_fwd(SolveArgs a, cudaStream_t stream);
int launch_lane_bwd(SolveArgs a, cudaStream_t stream);

}  // namespace ancde
