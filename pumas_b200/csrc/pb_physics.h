/*
 * pb_physics.h -- read-only view of the physics tables.
 *
 * The tables are the bytes of a `pumas_physics_dump` (src/pumas.c:2167): the
 * blob is kept as one allocation (host mirror, and one device copy in HBM) and
 * this POD holds typed pointers into it, valid for the memory space the code
 * runs in. Shapes follow the reference accessors (src/pumas.c:4260-4752):
 * rows are contiguous, so every (scheme, material) column is a dense array of
 * n_energies doubles -- already struct-of-arrays per material.
 */
#ifndef PB_PHYSICS_H
#define PB_PHYSICS_H

#include "pb_common.h"

namespace pb {

/* Per element constants. Z, A, I come from the dump (struct atomic_element,
 * pumas.c:510-523); the rest is folded at load time from pb_atomic_data.h so
 * that the kernels never touch the Z-indexed tables. */
struct Element {
        double Z, A, I;
        /* Coulomb screening, pumas.c:7351-7386 */
        int n_screening;          /* number of atomic terms: 1, 2 or 3 */
        int pad_;
        double amplitude[3];
        double alpha_over_a0[3];  /* exponent[i][iZ] / BOHR_RADIUS */
        double nuclear_radius;    /* data_nuclear_radius(Z, A), m; pumas.c:7143 */
        double radiation_log;     /* radiation_logarithm(Z); pumas.c:13709 */
        /* Powers the exact DCS models recompute on every call. */
        double z13;               /* pow(Z, 1. / 3.)   pumas.c:13922, 14234 */
        double z13_inv;           /* pow(Z, -1. / 3.)  pumas.c:14417 */
        double z23;               /* pow(Z, 2. / 3.)   pumas.c:14434 */
        double Dn;                /* 1.54 * pow(A, 0.27) pumas.c:13943, 14418 */
        double lnA;               /* log(A)            pumas.c:14694 */
        double a13;               /* pow(A, 1. / 3.)   pumas.c:14738, 14803, 15247 */
        double wood_saxon;        /* dcs_wood_saxon(Z, A) pumas.c:14732-14748 */
};

/* Index of a user registered DCS model in Physics::model (the built-in ones
 * are 0 .. 2, pb_dcs_models.h). */
constexpr int MODEL_USER = 3;

struct Physics {
        int n_energies;
        int n_materials;
        int n_composites;
        int n_elements;
        int n_components;
        int max_components;
        int n_table_dcs;
        int particle;
        /* DCS models the tables were built with (pb_dcs_models.h): [0]
         * bremsstrahlung, [1] pair production, [2] photonuclear */
        int model[3];
        int pad_;
        /* model[i] == MODEL_USER: a model registered by the user
         * (pumas_dcs_register, pumas.c:15422), as a function of the side the
         * view belongs to: a host function in the host view, a `__device__`
         * function (pumas_b200_dcs_register_device) in a device view. Only
         * kernels compiled with PB_USER_DCS (kernels_callback.cu) call it. */
        double (*user_dcs[3])(double Z, double A, double m, double K, double q);
        double ctau;
        double mass;
        double cutoff;
        double elastic_ratio;
        /* exp() of literal arguments, evaluated once by the host libm. */
        double c_exp_1;    /* exp(1.)       pumas.c:13945 */
        double c_exp_1_2;  /* exp(0.5)      pumas.c:13957, 14468 */
        double c_exp_m1_6; /* exp(-1. / 6.) pumas.c:13958, 14472 */
        double c_exp_1_3;  /* exp(1. / 3.)  pumas.c:13959, 14473 */
        double c_exp_m1_3; /* exp(-1. / 3.) pumas.c:14476 */

        const double * K;
        const double * K_dX;
        const double * K_dNI_in;
        const double * K_dNI_el;
        const double * X;
        const double * X_dK;
        const double * X_dT;
        const double * T;
        const double * T_dK;
        const double * dE;
        const double * dE_dK;
        const double * Omega;
        const double * Omega_dK;
        const double * NI_el;
        const double * NI_el_dK;
        const double * NI_in;
        const double * NI_in_dK;
        const double * CS;
        const double * CS_dK;
        const double * CSf;
        const double * CSf_dK;
        const double * CSn;
        const double * CSn_dK;
        const float * DCS;
        const float * DCS_x;
        const float * DCS_envelope;
        const double * Xt;
        const double * Kt;
        const double * Li;
        const double * Li_dK;
        const double * a_max;
        const double * b_max;
        const double * Mu0;
        const double * Mu0_dK;
        const double * Lb;
        const double * Lb_dK;
        const double * Ms1;
        const double * Ms1_dK;
        const int * elements_in;
        const double * material_density;

        /* Flattened composition (struct material_component, pumas.c:527):
         * components of material m are [comp_offset[m], comp_offset[m + 1]).
         * The same running index addresses table_CSf (pumas.c:6179). */
        const int * comp_offset;
        const int * comp_element;
        const double * comp_fraction;
        const Element * element;

        /* The tables of a Monte Carlo step re-packed per grid cell
         * (pb_hot_tables.h): [material][cell][HOTK_WORDS] and
         * [material][cell][HOTX_WORDS]. */
        const double * hotK;
        const double * hotX;
        /* The same for the closed form CSDA (pb_csda.h): packed records of
         * the CSDA scheme and the O(1) locators of the kinetic grid and of the
         * grammage column of every material (a common key range). */
        const double * csdaK;
        const double * csdaX;
        const unsigned short * locK;
        const unsigned short * locX; /* [material][locX_keys] */
        int locK_min, locK_keys, locX_min, locX_keys;
};

/* ---- table accessors, pumas.c:4260-4752 -------------------------------- */
PB_FN int clamp_scheme(int scheme)
{
        return (scheme > MODE_MIXED) ? MODE_MIXED : scheme;
}

/* Column of a [scheme][material][row] table. */
PB_FN const double * col_sm(
    const Physics & p, const double * table, int scheme, int material)
{
        return table +
            (clamp_scheme(scheme) * p.n_materials + material) * p.n_energies;
}

/* Column of a [material][row] table. */
PB_FN const double * col_m(const Physics & p, const double * table, int material)
{
        return table + material * p.n_energies;
}

/* NI_el uses `scheme >= MIXED ? MIXED : CSDA` (pumas.c:4410). */
PB_FN const double * col_NI_el(
    const Physics & p, const double * table, int scheme, int material)
{
        const int s = (scheme >= MODE_MIXED) ? MODE_MIXED : MODE_CSDA;
        return table + (s * p.n_materials + material) * p.n_energies;
}

/* Ms1 has three slabs indexed by scheme + 1 (pumas.c:4684). */
PB_FN const double * col_Ms1(
    const Physics & p, const double * table, int scheme, int material)
{
        return table +
            ((clamp_scheme(scheme) + 1) * p.n_materials + material) *
            p.n_energies;
}

PB_FN double b_max(const Physics & p, int scheme, int material)
{
        return p.b_max[clamp_scheme(scheme) * p.n_materials + material];
}

} /* namespace pb */

#endif
