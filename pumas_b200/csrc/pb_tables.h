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

/* pumas.c:4146-4157. No bound check: x must lie inside table_X. Out of line:
 * only the cold call sites use the generic search. */
PB_NOINLINE double interpolate_pchip(const Physics & p, const double * table_X,
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
                return X[imax] + 1. / b * xlog((kinetic + K1) / (K0 + K1));
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
                    E0 / (a + b * E0) * xexp(a * (time - tmax) / p.mass);
                const double kinetic = a * r / (1. - b * r) - p.mass;
                const double K1 = a / b + p.mass;
                return X[imax] +
                    1. / b * xlog((kinetic + K1) / (E0 - p.mass + K1));
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
                    p.mass / a * xlog((E1 / E0) * (a + b * E0) / (a + b * E1));
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
                return (K0 + K1) * xexp(b * (grammage - X[imax])) - K1;
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
                    xlog((a0 + b0 * (kinetic + p.mass)) /
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
                return ((a0 + b0 * (k0 + p.mass)) * xexp(b0 * (nI - nI0) / cs) -
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

/* ------------------------------------------------------------------------
 * Located lookups (hot path).
 *
 * Inside a Monte Carlo step the reference interpolates a dozen columns at the
 * same two or three energies, searching the grid again every time (it only
 * keeps a two entry index memory, pumas.c:4187-4190). Here a value is located
 * ONCE in its grid (`Loc`: cell index, reduced abscissa t and cell width, the
 * exact quantities math_pchip_interpolate receives at pumas.c:4150-4156) and
 * every column is then evaluated from the `Loc`. The search starts from a hint
 * (the cell of the previous lookup: the energy moves by ~1 % per step, the
 * grid spacing is 9-25 %) and walks to the right cell, which on a strictly
 * increasing table is the cell the reference's bisection finds. Results are
 * bit identical; dependent loads per step drop from ~50 to ~10.
 * ------------------------------------------------------------------------ */
struct Loc {
        int i;     /* cell; -1: below table[0]; n_energies - 1: >= last node */
        double t;  /* (x - X[i]) / dX, when 0 <= i < n_energies - 1 */
        double dX; /* X[i + 1] - X[i] */
        double x;  /* the located value */
};

/* Locate x in a strictly increasing table of n_energies nodes, starting the
 * walk at cell `hint`; falls back to a bisection when the hint is far off. */
PB_NOINLINE Loc locate_near_n(
    const double * table, int n_energies, double x, int hint)
{
        const int imax = n_energies - 1;
        Loc l;
        l.x = x;
        l.t = 0.;
        l.dX = 0.;
        if (x < table[0]) {
                l.i = -1;
                return l;
        }
        if (x >= table[imax]) {
                l.i = imax;
                return l;
        }
        int i = hint;
        if (i < 0) i = 0;
        if (i > imax - 1) i = imax - 1;
        int budget = 6;
        while ((x < table[i]) && (budget > 0)) {
                --i;
                --budget;
        }
        while ((x >= table[i + 1]) && (budget > 0)) {
                ++i;
                --budget;
        }
        if (!((x >= table[i]) && (x < table[i + 1]))) {
                /* far hint: the reference's bisection, pumas.c:4199-4200 */
                int i1 = 0, i2 = imax;
                do {
                        const int i3 = (i1 + i2) / 2;
                        if (x >= table[i3])
                                i1 = i3;
                        else
                                i2 = i3;
                } while (i2 - i1 >= 2);
                i = i1;
        }
        l.i = i;
        l.dX = table[i + 1] - table[i];
        l.t = (x - table[i]) / l.dX;
        return l;
}

PB_FN Loc locate_near(
    const Physics & p, const double * table, double x, int hint)
{
        return locate_near_n(table, p.n_energies, x, hint);
}

/* table_interpolate_pchip (pumas.c:4146-4157) from a located value. Inline:
 * 34 instructions, fewer than a call costs in register saves around it. */
PB_FN double pchip_at_n(
    int i1, double t, double dX, const double * table_Y, const double * table_M)
{
        const int i2 = i1 + 1;
        const double m1 = table_M[i1] * dX;
        const double m2 = (i2 > 1) ? table_M[i2] * dX : m1;
        return pchip(t, table_Y[i1], table_Y[i2], m1, m2);
}

PB_FN double pchip_at(const Loc & l, const double * table_Y, const double * table_M)
{
        return pchip_at_n(l.i, l.t, l.dX, table_Y, table_M);
}

/* Extrapolations above the last grid energy (10^9 GeV for the default grid):
 * kept out of line, they are never reached in practice. */
PB_NOINLINE double cel_grammage_above(
    const Physics & p, int scheme, int material, double kinetic)
{
        const int imax = p.n_energies - 1;
        const double a = p.a_max[material];
        const double b = b_max(p, scheme, material);
        const double K0 = p.K[imax];
        const double K1 = a / b + p.mass;
        return col_sm(p, p.X, scheme, material)[imax] +
            1. / b * xlog((kinetic + K1) / (K0 + K1));
}

PB_NOINLINE double cel_kinetic_energy_above(
    const Physics & p, int scheme, int material, double grammage)
{
        const int imax = p.n_energies - 1;
        const double a = p.a_max[material];
        const double b = b_max(p, scheme, material);
        const double K0 = p.K[imax];
        const double K1 = a / b + p.mass;
        return (K0 + K1) *
            xexp(b * (grammage - col_sm(p, p.X, scheme, material)[imax])) -
            K1;
}

/* cel_grammage, pumas.c:3644-3668, at a located kinetic energy. */
PB_FN double cel_grammage_at(
    const Physics & p, int scheme, int material, const Loc & l)
{
        if (l.i < 0) return 0.;
        if (l.i >= p.n_energies - 1)
                return cel_grammage_above(p, scheme, material, l.x);
        return pchip_at(l, col_sm(p, p.X, scheme, material),
            col_sm(p, p.X_dK, scheme, material));
}

/* cel_kinetic_energy, pumas.c:3758-3784, at a grammage located in the X
 * column of (scheme, material). */
PB_FN double cel_kinetic_energy_at(
    const Physics & p, int scheme, int material, const Loc & l)
{
        if (l.i < 0) return 0.;
        if (l.i >= p.n_energies - 1)
                return cel_kinetic_energy_above(p, scheme, material, l.x);
        return pchip_at(l, p.K, col_sm(p, p.K_dX, scheme, material));
}

/* cel_energy_loss, pumas.c:3796-3815. */
PB_FN double cel_energy_loss_at(
    const Physics & p, int scheme, int material, const Loc & l)
{
        if (l.i < 0) return 0.;
        if (l.i >= p.n_energies - 1)
                return p.a_max[material] +
                    b_max(p, scheme, material) * (l.x + p.mass);
        return pchip_at(l, col_sm(p, p.dE, scheme, material),
            col_sm(p, p.dE_dK, scheme, material));
}

/* cel_straggling, pumas.c:3826-3846. */
PB_FN double cel_straggling_at(const Physics & p, int material, const Loc & l)
{
        if (l.i < 0) return 0.;
        const int imax = p.n_energies - 1;
        const double * Omega = col_m(p, p.Omega, material);
        if (l.i >= imax) {
                const double r = (l.x + p.mass) / (p.K[imax] + p.mass);
                return Omega[imax] * r * r;
        }
        return pchip_at(l, Omega, col_m(p, p.Omega_dK, material));
}

/* del_cross_section, pumas.c:3895-3912. */
PB_FN double del_cross_section_at(const Physics & p, int material, const Loc & l)
{
        if (l.i < 0) return 0.;
        const double * CS = col_m(p, p.CS, material);
        if (l.i >= p.n_energies - 1) return CS[p.n_energies - 1];
        return pchip_at(l, CS, col_m(p, p.CS_dK, material));
}

/* table_get_msc, pumas.c:4077-4111, at a located kinetic energy. */
PB_FN void table_get_msc_at(const Physics & p, int scheme, int material,
    const Loc & l, double * mu0, double * invlb1)
{
        const int imax = p.n_energies - 1;
        const double kinetic = l.x;
        const double * Mu0 = col_m(p, p.Mu0, material);
        const double * Ms1 = col_Ms1(p, p.Ms1, scheme, material);
        if (kinetic < p.K[1]) {
                *mu0 = Mu0[1];
                *invlb1 = Ms1[1] * sqrt(p.K[1] / kinetic);
        } else if (l.i >= imax) {
                *mu0 = Mu0[imax];
                *invlb1 = Ms1[imax] * p.K[imax] / kinetic;
        } else {
                const int i1 = l.i;
                const double * y0 = Mu0 + i1;
                const double * m0 = col_m(p, p.Mu0_dK, material) + i1;
                const double * y1 = Ms1 + i1;
                const double * m1 = col_Ms1(p, p.Ms1_dK, scheme, material) + i1;
                const double dx = l.dX;
                const double t = l.t;
                *mu0 = pchip(t, y0[0], y0[1], m0[0] * dx, m0[1] * dx);
                *invlb1 = pchip(t, y1[0], y1[1], m1[0] * dx, m1[1] * dx);
        }
}

/* coulomb_ehs_length, pumas.c:7446-7466, at a located kinetic energy. */
PB_FN double coulomb_ehs_length_at(const Physics & p, int material, const Loc & l)
{
        const double kinetic = l.x;
        const double p2 = kinetic * (kinetic + 2. * p.mass);
        const int imax = p.n_energies - 1;
        const double * Lb = col_m(p, p.Lb, material);
        if (kinetic < p.K[1]) {
                return Lb[1] / p2;
        } else if (l.i >= imax) {
                return Lb[imax] / p2;
        } else {
                const int i1 = l.i;
                return (Lb[i1] + l.t * (Lb[i1 + 1] - Lb[i1])) / p2;
        }
}

/* A two entry memo of located kinetic energies, keyed by value: the energies
 * of a step (start, end before and after straggling) are each located once. */
struct KLocator {
        Loc a, b;
        int hint;
        int turn;
};

PB_FN void klocator_init(KLocator & k, int hint)
{
        k.a.x = k.b.x = -1.; /* kinetic energies are >= 0 */
        k.a.i = k.b.i = hint;
        k.a.t = k.b.t = k.a.dX = k.b.dX = 0.;
        k.hint = hint;
        k.turn = 0;
}

PB_FN Loc klocate(const Physics & p, KLocator & k, double kinetic)
{
        if (kinetic == k.a.x) return k.a;
        if (kinetic == k.b.x) return k.b;
        const Loc l = locate_near(p, p.K, kinetic, k.hint);
        if ((l.i >= 0) && (l.i < p.n_energies - 1)) k.hint = l.i;
        if (k.turn == 0)
                k.a = l;
        else
                k.b = l;
        k.turn ^= 1;
        return l;
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
