/*
 * user_media.cu -- user `__device__` code for libpumas_b200:
 *   - medium / locals callbacks (the device form of pumas_medium_cb /
 *     pumas_locals_cb, pumas.h:509-511, 348-349): the spherical shells of
 *     user_media.h;
 *   - a DCS model for pumas_dcs_register (pumas.h:2150-2167): user_dcs.h.
 *
 * Build (device-links this file with the transport kernels):
 *     python -m pumas_b200.build --user examples/user_media.cu \
 *         -o examples/libpumas_b200_user.so
 * then, from the host program:
 *     struct pumas_b200_callbacks cb;
 *     user_media_callbacks(&shells, materials, &cb);
 *     pumas_b200_context_callbacks_set(context, &cb);
 *     pumas_context_transport_n(context, n, states, events, media);
 */
#include <cuda_runtime.h>

#include "pumas_b200.h"

#define USER_MEDIA_FN static __host__ __device__ __forceinline__
#include "user_media.h"
#define USER_DCS_FN static __host__ __device__ __forceinline__
#include "user_dcs.h"

__device__ int user_media_medium_fn(const void * user_data,
    const struct pumas_state * state, int backward, int * medium, double * step)
{
        return user_media_locate(static_cast<const struct user_media *>(user_data),
            state->position, state->direction, backward, medium, step);
}

__device__ double user_media_locals_fn(const void * user_data, int medium,
    const struct pumas_state * state, double * density, double * magnet)
{
        magnet[0] = magnet[1] = magnet[2] = 0.;
        return user_media_density(
            static_cast<const struct user_media *>(user_data), medium,
            state->position, density);
}

/* Device function pointers are read from device variables. */
__device__ pumas_b200_medium_fn * const user_media_medium_ptr =
    user_media_medium_fn;
__device__ pumas_b200_locals_fn * const user_media_locals_ptr =
    user_media_locals_fn;

/* Fill the table for `shells` (copied to the device); materials[i] is the
 * physics material of shell i. Returns a cudaError_t. */
extern "C" int user_media_callbacks(const struct user_media * shells,
    const int * materials, struct pumas_b200_callbacks * cb)
{
        cudaError_t e = cudaMemcpyFromSymbol(
            &cb->medium, user_media_medium_ptr, sizeof cb->medium);
        if (e != cudaSuccess) return (int)e;
        e = cudaMemcpyFromSymbol(
            &cb->locals, user_media_locals_ptr, sizeof cb->locals);
        if (e != cudaSuccess) return (int)e;
        void * d = nullptr;
        e = cudaMalloc(&d, sizeof *shells);
        if (e != cudaSuccess) return (int)e;
        e = cudaMemcpy(d, shells, sizeof *shells, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return (int)e;
        cb->user_data = d;
        cb->n_media = shells->n_shells;
        for (int i = 0; i < shells->n_shells; i++) {
                cb->material[i] = materials[i];
                cb->has_locals[i] = 1;
        }
        return 0;
}

/* ---- a user DCS model -------------------------------------------------------
 * Host function for pumas_dcs_register, device function for
 * pumas_b200_dcs_register_device. */
static double user_dcs_host(double Z, double A, double m, double K, double q)
{
        return user_dcs_bremsstrahlung(Z, A, m, K, q);
}

__device__ double user_dcs_device_fn(
    double Z, double A, double m, double K, double q)
{
        return user_dcs_bremsstrahlung(Z, A, m, K, q);
}

__device__ pumas_dcs_t * const user_dcs_device_ptr = user_dcs_device_fn;

/* Register the model USER_DCS_MODEL for bremsstrahlung, on both sides.
 * Returns a pumas_return. */
extern "C" int user_dcs_register(void)
{
        pumas_dcs_t * device = nullptr;
        if (cudaMemcpyFromSymbol(&device, user_dcs_device_ptr, sizeof device) !=
            cudaSuccess)
                return (int)PUMAS_RETURN_INTERNAL_ERROR;
        enum pumas_return rc = pumas_dcs_register(
            PUMAS_PROCESS_BREMSSTRAHLUNG, USER_DCS_MODEL, &user_dcs_host);
        if (rc != PUMAS_RETURN_SUCCESS) return (int)rc;
        return (int)pumas_b200_dcs_register_device(
            PUMAS_PROCESS_BREMSSTRAHLUNG, USER_DCS_MODEL, device);
}

/* The host registration alone (no device needed): loading a dump, the host
 * API. */
extern "C" int user_dcs_register_host(void)
{
        return (int)pumas_dcs_register(
            PUMAS_PROCESS_BREMSSTRAHLUNG, USER_DCS_MODEL, &user_dcs_host);
}
