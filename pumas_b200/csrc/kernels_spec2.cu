/*
 * kernels_spec2.cu -- resident kernel specialised at compile time for
 * (STRAGGLED, DISABLED, FORWARD, WEIGHTED) in a single infinite medium without
 * locals (examples/pumas/straight.c geometry). Same source as the general
 * kernel (resident.cuh); only unreachable code is removed.
 */
#include "resident.cuh"

namespace pb {

PB_DEFINE_SPECIALISATION(launch_spec_straggled_noscat_forward, MODE_STRAGGLED, MODE_DISABLED, MODE_FORWARD, MODE_WEIGHTED)

} /* namespace pb */
