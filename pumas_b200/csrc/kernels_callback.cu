/*
 * kernels_callback.cu -- the transport kernel over `__device__` medium / locals
 * callbacks: the device counterparts of `pumas_medium_cb` (pumas.h:509-511)
 * and `pumas_locals_cb` (pumas.h:348-349), SURVEY.md section 8f.3.
 *
 * The general resident kernel (resident.cuh) instantiated on CallbackGeometry
 * (pb_media.h): every medium look-up and every locals evaluation of the
 * stepping (pumas.c:5528-5552, 6296-6925) is an indirect call through the
 * function table of `struct pumas_b200_callbacks`.
 *
 * A device function pointer can only be called by kernels of the same device
 * link unit (the callee's registers and stack are accounted for at link time).
 * So:
 *   - the built-in client below -- the planar media of pumas_b200_geometry.h
 *     behind the callback interface, user_data = a device copy of the
 *     descriptor -- lives in this file;
 *   - user callbacks are compiled with this file: `python -m pumas_b200.build
 *     --user my_media.cu` builds a library with `-rdc=true` where the user's
 *     translation unit is device-linked with the kernels (INTEGRATION.md
 *     section 6; examples/user_media.cu).
 */
/* This kernel calls the `__device__` functions of user registered DCS models
 * (pumas_dcs_register + pumas_b200_dcs_register_device; pb_dcs.h: dcs_model). */
#define PB_USER_DCS 1
#include "resident.cuh"

namespace pb {

/* ---- built-in client: the planar media as callbacks ----------------------- */
__device__ int planar_medium_fn(const void * user_data,
    const struct pumas_state * state, int backward, int * medium, double * step)
{
        const Geometry & g = *static_cast<const Geometry *>(user_data);
        return medium_locate(g, backward ? MODE_BACKWARD : MODE_FORWARD,
            *reinterpret_cast<const State *>(state), medium != nullptr, medium,
            step != nullptr, step);
}

__device__ double planar_locals_fn(const void * user_data, int medium,
    const struct pumas_state * state, double * density, double * magnet)
{
        const Geometry & g = *static_cast<const Geometry *>(user_data);
        return layer_locals(g.layer[medium],
            *reinterpret_cast<const State *>(state), density, magnet);
}

__device__ pumas_b200_medium_fn * const planar_medium_ptr = planar_medium_fn;
__device__ pumas_b200_locals_fn * const planar_locals_ptr = planar_locals_fn;

cudaError_t planar_callbacks(
    const Geometry & geo, struct pumas_b200_callbacks * cb)
{
        memset(cb, 0x0, sizeof *cb);
        cudaError_t e = cudaMemcpyFromSymbol(
            &cb->medium, planar_medium_ptr, sizeof cb->medium);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyFromSymbol(
            &cb->locals, planar_locals_ptr, sizeof cb->locals);
        if (e != cudaSuccess) return e;
        void * d = nullptr;
        e = cudaMalloc(&d, sizeof geo);
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(d, &geo, sizeof geo, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
                cudaFree(d);
                return e;
        }
        cb->user_data = d;
        cb->n_media =
            (geo.kind == PUMAS_B200_GEOMETRY_UNIFORM) ? 1 : geo.n_layers;
        for (int i = 0; i < cb->n_media; i++) {
                cb->material[i] = geo.layer[i].material;
                cb->has_locals[i] = geo.layer[i].has_locals;
        }
        return cudaSuccess;
}

/* ---- the kernel ------------------------------------------------------------ */
cudaError_t launch_transport_callbacks(const Physics & p, const Settings & cfg,
    const struct pumas_b200_callbacks & callbacks, uint64_t seed,
    uint64_t offset, long n, State * states, int * events, int * media,
    int * status, unsigned int * n_random, unsigned int * n_steps,
    void * workspace, cudaStream_t stream)
{
        if (n <= 0) return cudaSuccess;
        last_launch_count() = 1;
        CallbackGeometry geo;
        geo.cb = callbacks;
        return launch_resident(p, cfg, geo, seed, offset, n, states, events,
            media, status, n_random, n_steps, workspace, stream);
}

} /* namespace pb */
