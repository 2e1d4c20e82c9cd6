/*
 * kernels_csda.cu -- the closed form CSDA kernel (BASELINE.json configs[1]:
 * "forward CSDA-only transport ... pure table-lookup path").
 *
 * With energy loss = CSDA, no scattering and a single uniform medium without
 * locals, every track ends in transport_with_csda (pumas.c:4777-4949): half a
 * dozen PCHIP look-ups, no loop, no random number. Nothing needs scheduling, so
 * the kernel is one thread per track; instantiated on (StaticSettings,
 * UniformGeometry) the stepping engine is unreachable and compiled out, which
 * leaves a small kernel (64 registers instead of 255, 32 warps per SM; the
 * long scoreboard is its first stall, so occupancy is worth 24-48 B of spills:
 * +2 to +11 % against 80 registers and 24 warps, profiles/r1_experiments.md). Per
 * track it moves 104 B of state in, 104 B out, 4 B of event (+ 16 B of optional
 * media and counters): 2.6 TB/s at 1.1e10 tracks/s, 40 % of the measured HBM
 * peak, and about as much time goes to issuing the ~6 table searches of a
 * track. Each thread reads and writes its own 104 B record; staging the
 * records of a CTA through shared memory with coalesced 8 B words was measured
 * 15-25 % slower (two barriers, no overlap of the phases at 3-4 CTAs per SM;
 * profiles/r1_experiments.md), so the direct form stays.
 */
#include "resident.cuh"

namespace pb {

constexpr int CSDA_BLOCK = 256;
#ifndef PB_CSDA_MIN_CTAS
#define PB_CSDA_MIN_CTAS 4
#endif

template <class Cfg>
__global__ void __launch_bounds__(CSDA_BLOCK, PB_CSDA_MIN_CTAS)
    transport_csda_kernel(const Physics p, const Cfg cfg,
        const UniformGeometry geo, const BatchArgs a)
{
        const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= a.n) return;

        State state = a.states[i];
        DeviceRng rng;
        rng.init(a.seed, a.offset + (uint64_t)i);
        unsigned int steps = 0;
        const Outcome out = transport_track(p, cfg, geo, rng, state, &steps);
        a.states[i] = state;
        if (a.events != nullptr) a.events[i] = out.event;
        if (a.media != nullptr) {
                a.media[2 * i] = out.medium[0];
                a.media[2 * i + 1] = out.medium[1];
        }
        if (out.rc != RETURN_SUCCESS) atomicCAS(a.status, 0, out.rc);
        if (a.n_random != nullptr) a.n_random[i] = rng.count;
        if (a.n_steps != nullptr) a.n_steps[i] = steps;
}

template <class Cfg>
static cudaError_t launch_csda(const Physics & p, const Cfg & cfg,
    const UniformGeometry & ug, const BatchArgs & a, cudaStream_t stream)
{
        const long blocks = (a.n + CSDA_BLOCK - 1) / CSDA_BLOCK;
        transport_csda_kernel<Cfg><<<(unsigned int)blocks, CSDA_BLOCK, 0,
            stream>>>(p, cfg, ug, a);
        return cudaGetLastError();
}

/* Closed form CSDA in a plain uniform medium, decays weighted (the default).
 * Returns false when the settings / geometry are not of that kind. */
PB_SPECIALISATION(launch_spec_csda)
{
        (void)workspace;
        if ((cfg.energy_loss != MODE_CSDA) ||
            (cfg.scattering != MODE_DISABLED) ||
            (cfg.decay != MODE_WEIGHTED) || !geometry_is_plain_uniform(geo))
                return false;
        UniformGeometry ug;
        ug.material = geo.layer[0].material;
        ug.step_type = geo.step_type;
        BatchArgs a;
        a.seed = seed;
        a.offset = offset;
        a.n = n;
        a.states = states;
        a.events = events;
        a.media = media;
        a.status = status;
        a.n_random = n_random;
        a.n_steps = n_steps;
        a.cursor = nullptr;
        a.stats = nullptr;
        a.ehs_period = 0;
        a.del_threshold = 0;
        if (cfg.direction == MODE_FORWARD) {
                typedef StaticSettings<MODE_CSDA, MODE_DISABLED, MODE_FORWARD,
                    MODE_WEIGHTED>
                    Cfg;
                *error = launch_csda(p, Cfg(cfg), ug, a, stream);
        } else {
                typedef StaticSettings<MODE_CSDA, MODE_DISABLED, MODE_BACKWARD,
                    MODE_WEIGHTED>
                    Cfg;
                *error = launch_csda(p, Cfg(cfg), ug, a, stream);
        }
        return true;
}

} /* namespace pb */
