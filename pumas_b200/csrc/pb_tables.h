/*
 * pb_tables.h -- table lookups: bisection, cubic Hermite (PCHIP) evaluation
 * and the tabulated CEL / DEL / EHS properties with their analytic
 * extrapolations above the last grid energy.
 *
 * Restates src/pumas.c:3644-4244 (a11, a12, a13 of SURVEY.md section 8a). The
 * arithmetic is written in the reference's order of operations: results must
 * agree to the last bit when compiled without FMA contraction.
 */
#ifndef PB_TABLES_H
#define PB_TABLES_H

#include "pb_physics.h"

namespace pb {

/* pumas.c:4118-4125 */
PB_FN double pchip(double t, double p0, double p1, double m0, double m1)
{
        const double c2 = -3 * (p0 - p1) - 2 * m0 - m1;
        const double c3 = 2 * (p0 - p1) + m0 + m1;
        return p0 + t * (m0 + t * (c2 + t * c3));
}

/* Index of the grid cell holding `value`, from below: -1 on underflow, imax
 * on overflow. pumas.c:4174-4233. The two-entry index memory of the reference
 * (pumas.c:4187-4190) only short-cuts the search and cannot change its result
 * on a non-decreasing table, so a plain bisection is used. */
PB_FN int table_index(const Physics & p, const double * table, double value)
{
        const int imax = p.n_energies - 1;
        if (value < table[0]) return -1;
        if (value >= table[imax]) return imax;
        int i1 = 0, i2 = imax;
        do {
                const int i3 = (i1 + i2) / 2;
                if (value >= table[i3])
                        i1 = i3;
                else
                        i2 = i3;
        } while (i2 - i1 >= 2);
        return i1;
}

/* pumas.c:4146-4157. No bound check: x must lie inside table_X. */
PB_FN double interpolate_pchip(const Physics & p, const double * table_X,
    const double * table_Y, const double * table_M, double x)
{
        const int i1 = table_index(p, table_X, x);
        const int i2 = i1 + 1;
        const double dX = table_X[i2] - table_X[i1];
        const double t = (x - table_X[i1]) / dX;
        const double m1 = table_M[i1] * dX;
        const double m2 = (i2 > 1) ? table_M[i2] * dX : m1;
        return pchip(t, table_Y[i1], table_Y[i2], m1, m2);
}

/* pumas.c:3644-3668 */
PB_FN double cel_grammage(
    const Physics & p, int scheme, int material, double kinetic)
{
        const int imax = p.n_energies - 1;
        if (kinetic < p.K[0]) return 0.;
        const double * X = col_sm(p, p.X, scheme, material);
        if (kinetic >= p.K[imax]) {
                const double a = p.a_max[material];
                const double b = b_max(p, scheme, material);
                const double K0 = p.K[imax];
                const double K1 = a / b + p.mass;
                return X[imax] + 1. / b * log((kinetic + K1) / (K0 + K1));
        }
        return interpolate_pchip(
            p, p.K, X, col_sm(p, p.X_dK, scheme, material), kinetic);
}

/* pumas.c:3680-3708 */
PB_FN double cel_grammage_as_time(
    const Physics & p, int scheme, int material, double time)
{
        const int imax = p.n_energies - 1;
        const double * T = col_sm(p, p.T, scheme, material);
        const double * X = col_sm(p, p.X, scheme, material);
        if (time < T[0]) return 0.;
        if (time >= T[imax]) {
                const double a = p.a_max[material];
                const double b = b_max(p, scheme, material);
                const double E0 = p.K[imax] + p.mass;
                const double tmax = T[imax];
                const double r =
                    E0 / (a + b * E0) * exp(a * (time - tmax) / p.mass);
                const double kinetic = a * r / (1. - b * r) - p.mass;
                const double K1 = a / b + p.mass;
                return X[imax] +
                    1. / b * log((kinetic + K1) / (E0 - p.mass + K1));
        }
        return interpolate_pchip(
            p, T, X, col_sm(p, p.X_dT, scheme, material), time);
}

/* pumas.c:3720-3745 */
PB_FN double cel_proper_time(
    const Physics & p, int scheme, int material, double kinetic)
{
        const int imax = p.n_energies - 1;
        if (kinetic < p.K[0]) return 0.;
        const double * T = col_sm(p, p.T, scheme, material);
        if (kinetic >= p.K[imax]) {
                const double a = p.a_max[material];
                const double b = b_max(p, scheme, material);
                const double E0 = p.K[imax] + p.mass;
                const double E1 = kinetic + p.mass;
                return T[imax] +
                    p.mass / a * log((E1 / E0) * (a + b * E0) / (a + b * E1));
        }
        return interpolate_pchip(
            p, p.K, T, col_sm(p, p.T_dK, scheme, material), kinetic);
}

/* pumas.c:3758-3784 */
PB_FN double cel_kinetic_energy(
    const Physics & p, int scheme, int material, double grammage)
{
        const int imax = p.n_energies - 1;
        const double * X = col_sm(p, p.X, scheme, material);
        if (grammage < X[0]) return 0.;
        if (grammage >= X[imax]) {
                const double a = p.a_max[material];
                const double b = b_max(p, scheme, material);
                const double K0 = p.K[imax];
                const double K1 = a / b + p.mass;
                return (K0 + K1) * exp(b * (grammage - X[imax])) - K1;
        }
        return interpolate_pchip(
            p, X, p.K, col_sm(p, p.K_dX, scheme, material), grammage);
}

/* pumas.c:3796-3815 */
PB_FN double cel_energy_loss(
    const Physics & p, int scheme, int material, double kinetic)
{
        const int imax = p.n_energies - 1;
        if (kinetic < p.K[0]) return 0.;
        if (kinetic >= p.K[imax]) {
                return p.a_max[material] +
                    b_max(p, scheme, material) * (kinetic + p.mass);
        }
        return interpolate_pchip(p, p.K, col_sm(p, p.dE, scheme, material),
            col_sm(p, p.dE_dK, scheme, material), kinetic);
}

/* pumas.c:3826-3846 */
PB_FN double cel_straggling(const Physics & p, int material, double kinetic)
{
        const int imax = p.n_energies - 1;
        if (kinetic < p.K[0]) return 0.;
        const double Kmax = p.K[imax];
        const double * Omega = col_m(p, p.Omega, material);
        if (kinetic >= Kmax) {
                const double r = (kinetic + p.mass) / (Kmax + p.mass);
                return Omega[imax] * r * r;
        }
        return interpolate_pchip(
            p, p.K, Omega, col_m(p, p.Omega_dK, material), kinetic);
}

/* pumas.c:3861-3884 */
PB_FN double cel_magnetic_rotation(
    const Physics & p, int material, double kinetic)
{
        const int imax = p.n_energies - 1;
        const double factor = LARMOR_FACTOR / p.mass;
        const double * T = col_sm(p, p.T, MODE_CSDA, material);
        if (kinetic <= p.K[0]) return T[imax] * factor;
        if (kinetic >= p.K[imax]) return 0.;
        const double t = interpolate_pchip(
            p, p.K, T, col_sm(p, p.T_dK, MODE_CSDA, material), kinetic);
        return (T[imax] - t) * factor;
}

/* pumas.c:3895-3912 */
PB_FN double del_cross_section(const Physics & p, int material, double kinetic)
{
        const int imax = p.n_energies - 1;
        if (kinetic < p.K[0]) return 0.;
        const double * CS = col_m(p, p.CS, material);
        if (kinetic >= p.K[imax]) return CS[imax];
        return interpolate_pchip(
            p, p.K, CS, col_m(p, p.CS_dK, material), kinetic);
}

/* pumas.c:3923-3948 */
PB_FN double del_interaction_length(
    const Physics & p, int material, double kinetic)
{
        const int imax = p.n_energies - 1;
        if (kinetic < p.K[0]) return 0.;
        const double * NI = col_m(p, p.NI_in, material);
        if (kinetic >= p.K[imax]) {
                const double k0 = p.K[imax];
                const double a0 = p.a_max[material];
                const double b0 = b_max(p, MODE_MIXED, material);
                const double cs = col_m(p, p.CS, material)[imax];
                const double dZ = cs / b0 *
                    log((a0 + b0 * (kinetic + p.mass)) /
                        (a0 + b0 * (k0 + p.mass)));
                return NI[imax] + dZ;
        }
        return interpolate_pchip(
            p, p.K, NI, col_m(p, p.NI_in_dK, material), kinetic);
}

/* pumas.c:3959-3986 */
PB_FN double del_kinetic_from_interaction_length(
    const Physics & p, int material, double nI)
{
        const int imax = p.n_energies - 1;
        const double * NI = col_m(p, p.NI_in, material);
        if (nI < NI[0]) return 0.;
        if (nI >= NI[imax]) {
                const double k0 = p.K[imax];
                const double a0 = p.a_max[material];
                const double b0 = b_max(p, MODE_MIXED, material);
                const double cs = col_m(p, p.CS, material)[imax];
                const double nI0 = NI[imax];
                return ((a0 + b0 * (k0 + p.mass)) * exp(b0 * (nI - nI0) / cs) -
                           a0) /
                    b0 -
                    p.mass;
        }
        return interpolate_pchip(
            p, NI, p.K, col_m(p, p.K_dNI_in, material), nI);
}

/* pumas.c:3998-4022 */
PB_FN double ehs_interaction_length(
    const Physics & p, int scheme, int material, double kinetic)
{
        const int imax = p.n_energies - 1;
        if (kinetic < p.K[0]) return 0.;
        const double * NI = col_NI_el(p, p.NI_el, scheme, material);
        if (kinetic >= p.K[imax]) {
                const double k0 = p.K[imax - 1];
                const double k1 = p.K[imax];
                const double nI0 = NI[imax - 1];
                const double nI1 = NI[imax];
                return nI1 + (kinetic - k1) * (nI1 - nI0) / (k1 - k0);
        }
        return interpolate_pchip(
            p, p.K, NI, col_NI_el(p, p.NI_el_dK, scheme, material), kinetic);
}

/* pumas.c:4034-4058 */
PB_FN double ehs_kinetic_from_interaction_length(
    const Physics & p, int scheme, int material, double nI)
{
        const int imax = p.n_energies - 1;
        const double * NI = col_NI_el(p, p.NI_el, scheme, material);
        if (nI < NI[0]) return 0.;
        if (nI >= NI[imax]) {
                const double k0 = p.K[imax - 1];
                const double k1 = p.K[imax];
                const double nI0 = NI[imax - 1];
                const double nI1 = NI[imax];
                return k1 + (nI - nI1) * (k1 - k0) / (nI1 - nI0);
        }
        /* K_dNI_el is indexed like K_dX (scheme > MIXED -> MIXED,
         * pumas.c:4274-4281), unlike NI_el itself. */
        return interpolate_pchip(
            p, NI, p.K, col_sm(p, p.K_dNI_el, scheme, material), nI);
}

/* Multiple scattering parameters: angular cut-off mu0 and inverse first
 * transport grammage. pumas.c:4077-4111. `scheme` is the context energy loss
 * mode; DISABLED (-1) selects slab 0 of Ms1. */
PB_FN void table_get_msc(const Physics & p, int scheme, int material,
    double kinetic, double * mu0, double * invlb1)
{
        const int imax = p.n_energies - 1;
        const double * Mu0 = col_m(p, p.Mu0, material);
        const double * Ms1 = col_Ms1(p, p.Ms1, scheme, material);
        if (kinetic < p.K[1]) {
                *mu0 = Mu0[1];
                *invlb1 = Ms1[1] * sqrt(p.K[1] / kinetic);
        } else if (kinetic >= p.K[imax]) {
                *mu0 = Mu0[imax];
                *invlb1 = Ms1[imax] * p.K[imax] / kinetic;
        } else {
                const int i1 = table_index(p, p.K, kinetic);
                const double * x = p.K + i1;
                const double * y0 = Mu0 + i1;
                const double * m0 = col_m(p, p.Mu0_dK, material) + i1;
                const double * y1 = Ms1 + i1;
                const double * m1 = col_Ms1(p, p.Ms1_dK, scheme, material) + i1;
                const double dx = x[1] - x[0];
                const double t = (kinetic - x[0]) / dx;
                *mu0 = pchip(t, y0[0], y0[1], m0[0] * dx, m0[1] * dx);
                *invlb1 = pchip(t, y1[0], y1[1], m1[0] * dx, m1[1] * dx);
        }
}

/* Grammage mean free path of hard elastic events. pumas.c:7446-7466. */
PB_FN double coulomb_ehs_length(const Physics & p, int material, double kinetic)
{
        const double p2 = kinetic * (kinetic + 2. * p.mass);
        const int imax = p.n_energies - 1;
        const double * Lb = col_m(p, p.Lb, material);
        if (kinetic < p.K[1]) {
                return Lb[1] / p2;
        } else if (kinetic >= p.K[imax]) {
                return Lb[imax] / p2;
        } else {
                const int i1 = table_index(p, p.K, kinetic);
                const int i2 = i1 + 1;
                const double h = (kinetic - p.K[i1]) / (p.K[i2] - p.K[i1]);
                return (Lb[i1] + h * (Lb[i2] - Lb[i1])) / p2;
        }
}

} /* namespace pb */

#endif
