/*
 * kernels_spec3.cu -- resident kernel specialised at compile time for
 * (STRAGGLED, DISABLED, BACKWARD, WEIGHTED) in a single infinite medium without
 * locals (examples/pumas/straight.c geometry). Same source as the general
 * kernel (resident.cuh); only unreachable code is removed.
 */
#include "resident.cuh"

namespace pb {

PB_DEFINE_SPECIALISATION(launch_spec_straggled_noscat_backward, MODE_STRAGGLED, MODE_DISABLED, MODE_BACKWARD, MODE_WEIGHTED)

} /* namespace pb */
