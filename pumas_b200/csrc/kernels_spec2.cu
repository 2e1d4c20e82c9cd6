/*
 * kernels_spec2.cu -- resident kernel specialised at compile time for
 * (STRAGGLED, DISABLED, FORWARD, WEIGHTED) in a single infinite medium without
 * locals (examples/pumas/straight.c geometry). Same source as the general
 * kernel (resident.cuh); only unreachable code is removed.
 */
/* Without scattering the step is half as long and needs fewer registers: 16
 * warps of 128 registers beat the 12 of 168 the other kernels use (B200,
 * straight workload: 7.6e6 against 6.9e6 tracks/s). */
#define PB_RS_WARPS 16
#include "resident.cuh"

namespace pb {

PB_DEFINE_SPECIALISATION(launch_spec_straggled_noscat_forward, MODE_STRAGGLED, MODE_DISABLED, MODE_FORWARD, MODE_WEIGHTED)

} /* namespace pb */
