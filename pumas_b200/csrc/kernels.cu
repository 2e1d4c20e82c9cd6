/*
 * kernels.cu -- the general transport kernels and the kernel dispatch.
 *
 *   transport_resident_kernel<Settings, Geometry>  (resident.cuh) the product
 *     path for any configuration;
 *   specialised instantiations of the same template for the common
 *     configurations live in kernels_spec*.cu and are tried first;
 *   transport_csda_kernel (kernels_csda.cu): closed form CSDA in a uniform
 *     medium, one thread per track, HBM bound;
 *   transport_kernel: one thread runs one track to completion; kept for A/B
 *     parity and profiling (divergence: 10.4 of 32 lanes active,
 *     profiles/r1_v0).
 * The physics tables stay in HBM / L2 (1.6 MB) and are read through L1; a
 * track state crosses HBM twice (104 B in, 104 B out).
 */
#include "resident.cuh"

namespace pb {

/* ------------------------------------------------------------------------
 * Reference kernel: one thread, one track, whole loop.
 * ------------------------------------------------------------------------ */
__global__ void __launch_bounds__(TRANSPORT_BLOCK)
    transport_kernel(const Physics p, const Settings cfg, const Geometry geo,
        const uint64_t seed, const uint64_t offset, const long n,
        State * __restrict__ states, int * __restrict__ events,
        int * __restrict__ media, int * __restrict__ status,
        unsigned int * __restrict__ n_random, unsigned int * __restrict__ n_steps)
{
        const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;

        State state = states[i];
        DeviceRng rng;
        rng.init(seed, offset + (uint64_t)i);
        unsigned int steps = 0;
        const Outcome out = transport_track(p, cfg, geo, rng, state, &steps);
        states[i] = state;
        if (events != nullptr) events[i] = out.event;
        if (media != nullptr) {
                media[2 * i] = out.medium[0];
                media[2 * i + 1] = out.medium[1];
        }
        if (out.rc != RETURN_SUCCESS) atomicCAS(status, 0, out.rc);
        if (n_random != nullptr) n_random[i] = rng.count;
        if (n_steps != nullptr) n_steps[i] = steps;
}

/* ------------------------------------------------------------------------
 * The recorder on the device (pumas.h:432-458; hooks of pb_transport.h): one
 * thread per track, the general code instantiated on a recorder that appends
 * frames to the track's segment of a buffer in HBM. Contexts without a
 * recorder never run this kernel: recording costs them nothing.
 * ------------------------------------------------------------------------ */
struct DeviceRecorder {
        static constexpr bool enabled = true;
        DeviceFrame * frames; /* this track's segment, or NULL */
        int capacity;
        int * count;
        int period_;
        __device__ __forceinline__ int period() const { return period_; }
        __device__ __forceinline__ void frame(
            int medium, int event, const State & state) const
        {
                const int k = (*count)++;
                if ((frames != nullptr) && (k < capacity)) {
                        frames[k].state = state;
                        frames[k].medium = medium;
                        frames[k].event = event;
                }
        }
};

__global__ void __launch_bounds__(TRANSPORT_BLOCK)
    transport_record_kernel(const Physics p, const Settings cfg,
        const Geometry geo, const uint64_t seed, const uint64_t offset,
        const long n, State * __restrict__ states, int * __restrict__ events,
        int * __restrict__ media, int * __restrict__ status,
        unsigned int * __restrict__ n_random,
        unsigned int * __restrict__ n_steps, int period,
        DeviceFrame * __restrict__ frames, const long * __restrict__ first,
        int * __restrict__ counts)
{
        const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;

        State state = states[i];
        DeviceRng rng;
        rng.init(seed, offset + (uint64_t)i);
        unsigned int steps = 0;
        int count = 0;
        DeviceRecorder rec;
        rec.frames = (frames != nullptr) ? frames + first[i] : nullptr;
        rec.capacity = (frames != nullptr) ? (int)(first[i + 1] - first[i]) : 0;
        rec.count = &count;
        rec.period_ = period;
        const Outcome out =
            transport_track(p, cfg, geo, rng, state, &steps, rec);
        states[i] = state;
        counts[i] = count;
        if (events != nullptr) events[i] = out.event;
        if (media != nullptr) {
                media[2 * i] = out.medium[0];
                media[2 * i + 1] = out.medium[1];
        }
        if (out.rc != RETURN_SUCCESS) atomicCAS(status, 0, out.rc);
        if (n_random != nullptr) n_random[i] = rng.count;
        if (n_steps != nullptr) n_steps[i] = steps;
}

cudaError_t launch_transport_recorded(const Physics & p, const Settings & cfg,
    const Geometry & geo, uint64_t seed, uint64_t offset, long n,
    State * states, int * events, int * media, int * status,
    unsigned int * n_random, unsigned int * n_steps, int period,
    DeviceFrame * frames, const long * first, int * counts,
    cudaStream_t stream)
{
        if (n <= 0) return cudaSuccess;
        last_launch_count() = 1;
        const long blocks = (n + TRANSPORT_BLOCK - 1) / TRANSPORT_BLOCK;
        transport_record_kernel<<<(unsigned int)blocks, TRANSPORT_BLOCK, 0,
            stream>>>(p, cfg, geo, seed, offset, n, states, events, media,
            status, n_random, n_steps, period, frames, first, counts);
        return cudaGetLastError();
}

PB_SPECIALISATION(launch_spec_straggled_mixed_forward);
PB_SPECIALISATION(launch_spec_straggled_mixed_backward);
PB_SPECIALISATION(launch_spec_straggled_noscat_forward);
PB_SPECIALISATION(launch_spec_straggled_noscat_backward);
PB_SPECIALISATION(launch_spec_geo_straggled_mixed_backward);
bool launch_spec_csda(const Physics & p, const Physics & host_view,
    const Settings & cfg, const Geometry & geo, uint64_t seed, uint64_t offset,
    long n, State * states, int * events, int * media, int * status,
    unsigned int * n_random, unsigned int * n_steps, void * workspace,
    int kernel, cudaStream_t stream, cudaError_t * error);

/* hot_kernel.cuh instantiations (kernels_hot*.cu) */
#define PB_HOT_SPECIALISATION(NAME)                                            \
        bool NAME(const Physics & p, const Physics & host_view,                \
            const Settings & cfg, const Geometry & geo, uint64_t seed,         \
            uint64_t offset, long n, State * states, int * events,             \
            int * media, int * status, unsigned int * n_random,                \
            unsigned int * n_steps, void * workspace, cudaStream_t stream,     \
            cudaError_t * error)
PB_HOT_SPECIALISATION(launch_hot_straggled_mixed_forward);
PB_HOT_SPECIALISATION(launch_hot_straggled_mixed_backward);
PB_HOT_SPECIALISATION(launch_hot_straggled_noscat_forward);
PB_HOT_SPECIALISATION(launch_hot_straggled_noscat_backward);

int transport_last_launches() { return last_launch_count(); }

cudaError_t transport_workspace_bytes(size_t * bytes)
{
        *bytes = workspace_bytes();
        return cudaSuccess;
}

cudaError_t launch_transport(const Physics & p, const Physics & host_view,
    const Settings & cfg, const Geometry & geo, uint64_t seed, uint64_t offset, long n,
    State * states, int * events, int * media, int * status,
    unsigned int * n_random, unsigned int * n_steps, void * workspace,
    int kernel, cudaStream_t stream)
{
        if (n <= 0) return cudaSuccess;
        last_launch_count() = 1;
        if (kernel == KERNEL_MONOLITHIC) {
                const long blocks = (n + TRANSPORT_BLOCK - 1) / TRANSPORT_BLOCK;
                transport_kernel<<<(unsigned int)blocks, TRANSPORT_BLOCK, 0,
                    stream>>>(p, cfg, geo, seed, offset, n, states, events,
                    media, status, n_random, n_steps);
                return cudaGetLastError();
        }

        if (kernel == KERNEL_RESIDENT) {
                cudaError_t e = cudaSuccess;
#define TRY_HOT(NAME)                                                          \
        if (NAME(p, host_view, cfg, geo, seed, offset, n, states, events,      \
                media, status, n_random, n_steps, workspace, stream, &e))      \
                return e;
                TRY_HOT(launch_hot_straggled_mixed_forward)
                TRY_HOT(launch_hot_straggled_mixed_backward)
                TRY_HOT(launch_hot_straggled_noscat_forward)
                TRY_HOT(launch_hot_straggled_noscat_backward)
#undef TRY_HOT
        }
        if ((kernel == KERNEL_RESIDENT) || (kernel == KERNEL_RESIDENT_SPEC)) {
                cudaError_t e = cudaSuccess;
                if (launch_spec_csda(p, host_view, cfg, geo, seed, offset, n,
                        states, events, media, status, n_random, n_steps,
                        workspace, kernel, stream, &e))
                        return e;
#define TRY(NAME)                                                              \
        if (NAME(p, cfg, geo, seed, offset, n, states, events, media, status,  \
                n_random, n_steps, workspace, stream, &e))                     \
                return e;
                TRY(launch_spec_straggled_mixed_forward)
                TRY(launch_spec_straggled_mixed_backward)
                TRY(launch_spec_straggled_noscat_forward)
                TRY(launch_spec_straggled_noscat_backward)
                /* (any descriptor geometry) */
                TRY(launch_spec_geo_straggled_mixed_backward)
#undef TRY
        }
        /* KERNEL_RESIDENT_GENERAL, or no specialisation applies. */
        return launch_resident(p, cfg, geo, seed, offset, n, states, events,
            media, status, n_random, n_steps, workspace, stream);
}

} /* namespace pb */
