/*
 * kernels_hot3.cu -- shared memory resident kernel (hot_kernel.cuh) for
 * (STRAGGLED, MODE_DISABLED, MODE_BACKWARD, WEIGHTED) in a single infinite medium without locals
 * (examples/pumas/straight.c geometry).
 */
#include "hot_kernel.cuh"

namespace pb {

PB_DEFINE_HOT_SPECIALISATION(launch_hot_straggled_noscat_backward, MODE_STRAGGLED, MODE_DISABLED, MODE_BACKWARD, MODE_WEIGHTED)

} /* namespace pb */
