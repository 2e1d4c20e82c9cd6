/*
 * kernels.h -- host interface of the transport kernels (kernels.cu).
 */
#ifndef PB_KERNELS_H
#define PB_KERNELS_H

#include <cuda_runtime.h>

#include "pb_media.h"
#include "pb_physics.h"

namespace pb {

constexpr int TRANSPORT_BLOCK = 128;

enum {
        /* persistent CTAs, one track per lane kept in registers, events
         * regrouped through shared memory pools and queues (resident.cuh);
         * picks a specialised instantiation when one matches the settings */
        KERNEL_RESIDENT = 0,
        /* one thread runs one track to completion */
        KERNEL_MONOLITHIC = 1,
        /* the resident kernel, general instantiation only */
        KERNEL_RESIDENT_GENERAL = 2,
        /* the register resident kernel with its specialised instantiations,
         * without the shared memory resident kernel (hot_kernel.cuh) that
         * KERNEL_RESIDENT tries first (A/B) */
        KERNEL_RESIDENT_SPEC = 3
};

/* Size of the device scratch buffer launch_transport needs (work cursor and
 * the global work cursor). */
cudaError_t transport_workspace_bytes(size_t * bytes);
/* Kernels launched by the last launch_transport (1, or 2 when the queues of
 * the DEL CTAs are prepared by resident_offload_init). */
int transport_last_launches();

/* A recorded frame on the device: struct pumas_frame (pumas.h:380-392) with the
 * medium as a layer index (-1 = none). */
struct DeviceFrame {
        State state;
        int medium;
        int event;
};

/* Transport n tracks with a recorder of the given period (pumas.h:432-458),
 * one thread per track on the general code. Track i appends its frames at
 * frames[first[i] ...] (at most first[i + 1] - first[i] of them are stored)
 * and writes how many it produced to counts[i]. `frames` / `first` may be
 * NULL: count only. */
cudaError_t launch_transport_recorded(const Physics & p, const Settings & cfg,
    const Geometry & geo, uint64_t seed, uint64_t offset, long n,
    State * states, int * events, int * media, int * status,
    unsigned int * n_random, unsigned int * n_steps, int period,
    DeviceFrame * frames, const long * first, int * counts,
    cudaStream_t stream);

/* Transport n tracks. `p` is the device view of the physics tables and
 * `host_view` the host one (per launch constants are read from it). All
 * other pointers are device pointers; `status` receives the
 * first non-zero pumas_return raised by a track, `workspace` is a scratch
 * buffer of transport_workspace_bytes(). Track i draws from the Philox stream
 * (seed, offset + i). */
cudaError_t launch_transport(const Physics & p, const Physics & host_view,
    const Settings & cfg, const Geometry & geo, uint64_t seed, uint64_t offset, long n,
    State * states, int * events, int * media, int * status,
    unsigned int * n_random, unsigned int * n_steps, void * workspace,
    int kernel, cudaStream_t stream);

/* The same over `__device__` medium / locals callbacks (kernels_callback.cu):
 * the general resident kernel instantiated on CallbackGeometry. */
cudaError_t launch_transport_callbacks(const Physics & p, const Settings & cfg,
    const struct pumas_b200_callbacks & callbacks, uint64_t seed,
    uint64_t offset, long n, State * states, int * events, int * media,
    int * status, unsigned int * n_random, unsigned int * n_steps,
    void * workspace, cudaStream_t stream);
/* The planar media of `geo` as callbacks (user_data: a device copy of `geo`,
 * to be released with cudaFree). */
cudaError_t planar_callbacks(
    const Geometry & geo, struct pumas_b200_callbacks * cb);

} /* namespace pb */

#endif
