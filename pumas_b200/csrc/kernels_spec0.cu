/*
 * kernels_spec0.cu -- resident kernel specialised at compile time for
 * (STRAGGLED, MIXED, FORWARD, WEIGHTED) in a single infinite medium without
 * locals (examples/pumas/straight.c geometry). Same source as the general
 * kernel (resident.cuh); only unreachable code is removed.
 */
/* Hard inelastic events served by dedicated CTAs (resident.cuh). */
#define PB_RS_OFFLOAD 1
#include "resident.cuh"

namespace pb {

PB_DEFINE_SPECIALISATION(launch_spec_straggled_mixed_forward, MODE_STRAGGLED, MODE_MIXED, MODE_FORWARD, MODE_WEIGHTED)

} /* namespace pb */
