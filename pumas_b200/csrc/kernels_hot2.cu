/*
 * kernels_hot2.cu -- shared memory resident kernel (hot_kernel.cuh) for
 * (STRAGGLED, MODE_DISABLED, MODE_FORWARD, WEIGHTED) in a single infinite medium without locals
 * (examples/pumas/straight.c geometry).
 */
#include "hot_kernel.cuh"

namespace pb {

PB_DEFINE_HOT_SPECIALISATION(launch_hot_straggled_noscat_forward, MODE_STRAGGLED, MODE_DISABLED, MODE_FORWARD, MODE_WEIGHTED)

} /* namespace pb */
