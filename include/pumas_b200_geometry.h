/*
 * pumas_b200_geometry.h -- POD description of the built-in device media.
 *
 * Split from pumas_b200.h so that the CPU oracle (oracle/oracle_driver.c, which
 * compiles the reference with its own pumas.h) can interpret the very same
 * descriptor through host callbacks.
 */
#ifndef PUMAS_B200_GEOMETRY_H
#define PUMAS_B200_GEOMETRY_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * Built-in device media (replace the host `pumas_medium_cb` /
 * `pumas_locals_cb` callbacks, pumas.h:348, pumas.h:509).
 * ------------------------------------------------------------------------ */
#define PUMAS_B200_MAX_LAYERS 8

enum pumas_b200_geometry_kind {
        /* One infinite medium, medium callback proposes step 0 (no boundary)
         * with PUMAS_STEP_CHECK: examples/pumas/straight.c:61-74. */
        PUMAS_B200_GEOMETRY_UNIFORM = 0,
        /* Stack of horizontal layers z[i] <= z < z[i+1]; outside is the NULL
         * medium. Step proposal = distance to the next plane along the motion
         * (+ step_epsilon), `step_default` when moving parallel to the planes:
         * examples/pumas/geometry.c:104-163. */
        PUMAS_B200_GEOMETRY_LAYERS = 1
};

enum pumas_b200_density_model {
        PUMAS_B200_DENSITY_CONSTANT = 0,   /* geometry.c:64-76 (locals_rock) */
        PUMAS_B200_DENSITY_EXPONENTIAL = 1 /* geometry.c:79-97 (locals_air)  */
};

struct pumas_b200_layer {
        int material;       /* physics material index */
        int has_locals;     /* 0: pumas_medium::locals == NULL (pumas.c:5533) */
        int density_model;  /* enum pumas_b200_density_model */
        int reserved;
        double density;     /* rho0 [kg/m^3]; <= 0 selects the material default */
        double scale;       /* lambda [m] in rho0 * exp(-z / lambda) */
        double eps;         /* floor on |u_z| for the proposed locals step */
        double magnet[3];   /* uniform field [T] */
};

struct pumas_b200_geometry {
        int kind;           /* enum pumas_b200_geometry_kind */
        int n_layers;
        int step_type;      /* enum pumas_step returned by the medium function */
        int reserved;
        double step_default;  /* geometry.c:120 (1E+03) */
        double step_epsilon;  /* geometry.c:153 (1E-07) */
        double z[PUMAS_B200_MAX_LAYERS + 1];
        struct pumas_b200_layer layer[PUMAS_B200_MAX_LAYERS];
};

/* ------------------------------------------------------------------------
 * Device callbacks: the counterparts of `pumas_medium_cb` (pumas.h:509-511)
 * and `pumas_locals_cb` (pumas.h:348-349) as `__device__` functions, for
 * geometries the descriptors above do not cover. The pointers are DEVICE
 * function pointers (read from a `__device__` variable with
 * cudaMemcpyFromSymbol); they are only callable by kernels of the same device
 * link unit: the built-in clients of the library (pumas_b200_planar_callbacks),
 * or user functions linked with the relocatable build of the kernels
 * (libpumas_b200_rdc.a).
 * ------------------------------------------------------------------------ */
struct pumas_state;

/* Locate `state` and / or propose a step. `medium`, when not NULL, receives
 * the index of the medium (0 .. n_media - 1; -1 outside); `step`, when not
 * NULL, the proposed stepping distance [m] along the motion (<= 0: no
 * boundary). `backward` != 0: the particle moves opposite to state->direction.
 * Returns an enum pumas_step (PUMAS_STEP_CHECK = 0, PUMAS_STEP_RAW = 1). */
typedef int pumas_b200_medium_fn(const void * user_data,
    const struct pumas_state * state, int backward, int * medium,
    double * step);

/* Local properties of medium `medium` at `state`: density [kg/m^3] (<= 0
 * selects the material's default) and magnetic field [T]. Returns the size of
 * the inhomogeneities [m] (<= 0: uniform), as pumas_locals_cb does. */
typedef double pumas_b200_locals_fn(const void * user_data, int medium,
    const struct pumas_state * state, double * density, double * magnet);

struct pumas_b200_callbacks {
        pumas_b200_medium_fn * medium; /* device function */
        pumas_b200_locals_fn * locals; /* device function, or NULL */
        const void * user_data;        /* device pointer handed back */
        int n_media;
        int material[PUMAS_B200_MAX_LAYERS];   /* physics material of medium i */
        int has_locals[PUMAS_B200_MAX_LAYERS]; /* 0: pumas_medium::locals == NULL
                                                * (pumas.c:5533) */
};

#ifdef __cplusplus
}
#endif
#endif
