/*
 * kernels_spec4.cu -- resident kernel specialised at compile time for
 * (STRAGGLED, MIXED, BACKWARD, WEIGHTED) in any descriptor geometry: the
 * detailed phase of examples/pumas/geometry.c (BASELINE.json configs[3]: rock
 * + exponential air, geomagnetic field, backward). Same source as the general
 * kernel (resident.cuh); only the code the settings make unreachable is
 * removed (CSDA and mixed loss branches, randomised decays, the forward
 * branches of every sampler).
 */
/* Hard inelastic events served by dedicated CTAs (resident.cuh); 16 warps of
 * 128 registers (B200, geometry_detailed, 1 M tracks: general kernel 1.89e6
 * tracks/s, this one with 12 warps of 168 registers 2.04e6 - 2.10e6, with 16
 * warps 2.17e6; same bits). */
#define PB_RS_OFFLOAD 1
#define PB_RS_WARPS 16
#include "resident.cuh"

namespace pb {

PB_DEFINE_SPECIALISATION_GEO(launch_spec_geo_straggled_mixed_backward, MODE_STRAGGLED, MODE_MIXED, MODE_BACKWARD, MODE_WEIGHTED)

} /* namespace pb */
