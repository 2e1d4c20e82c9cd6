/*
 * pb_hot_tables.h -- the physics tables of a Monte Carlo step, re-packed per
 * grid cell for the lean step of pb_hot.h (layout and rationale there). Built
 * once per physics object by the dump loader (pb_dump.hpp), on the host, with
 * the reference's own expressions (pumas.c:4118-4125, 4146-4157) and without
 * contraction: the device evaluates the same polynomial on the same
 * coefficients.
 */
#ifndef PB_HOT_TABLES_H
#define PB_HOT_TABLES_H

#include "pb_physics.h"

namespace pb {

/* ---- packed tables --------------------------------------------------------
 * K record of cell i (kinetic grid, material m, straggled / mixed scheme):
 *   [0] K[i]  [1] K[i+1]  [2] Lb[i]  [3] Lb[i+1]-Lb[i]
 *   [4..7] X  [8..11] Omega  [12..15] CS  [16..19] Ms1  [20..23] dE
 * X record of cell i (grammage column of the same scheme and material):
 *   [0] X[i]  [1] X[i+1]  [2] X[i+1]-X[i]  [3] unused  [4..7] K
 * each column as (p0, m0 * dx, c2, c3). */
constexpr int HOTK_WORDS = 24;
constexpr int HOTX_WORDS = 8;
constexpr int HK_LB = 2, HK_DLB = 3, HK_X = 4, HK_OMEGA = 8, HK_CS = 12,
              HK_MS1 = 16, HK_DE = 20;
constexpr int HX_K = 4;

/* The four coefficients of one cell of one column: table_interpolate_pchip
 * (pumas.c:4146-4157) up to the evaluation of the polynomial. */
inline void hot_pack_column(const double * table_Y, const double * table_M,
    int i1, double dX, double * out)
{
        const int i2 = i1 + 1;
        const double m1 = table_M[i1] * dX;
        const double m2 = (i2 > 1) ? table_M[i2] * dX : m1;
        const double p0 = table_Y[i1], p1 = table_Y[i2];
        /* math_pchip_interpolate, pumas.c:4118-4125 */
        const double c2 = -3 * (p0 - p1) - 2 * m1 - m2;
        const double c3 = 2 * (p0 - p1) + m1 + m2;
        out[0] = p0;
        out[1] = m1;
        out[2] = c2;
        out[3] = c3;
}

/* Number of doubles of the packed tables of a physics object. */
inline size_t hot_tables_words(const Physics & p)
{
        return (size_t)p.n_materials * (size_t)(p.n_energies - 1) *
            (HOTK_WORDS + HOTX_WORDS);
}

/* Fill the packed tables (host, at load time): first the K records of every
 * material, then the X records. Scheme index MODE_MIXED: the tables the
 * straggled and mixed modes read (pumas.c:4260-4281, 4684). */
inline void hot_tables_build(const Physics & p, double * hotK, double * hotX)
{
        const int nc = p.n_energies - 1;
        const int scheme = MODE_MIXED;
        for (int m = 0; m < p.n_materials; m++) {
                const double * X = col_sm(p, p.X, scheme, m);
                for (int i = 0; i < nc; i++) {
                        double * r =
                            hotK + ((size_t)m * nc + i) * HOTK_WORDS;
                        for (int k = 0; k < HOTK_WORDS; k++) r[k] = 0.;
                        const double dK = p.K[i + 1] - p.K[i];
                        r[0] = p.K[i];
                        r[1] = p.K[i + 1];
                        r[HK_LB] = col_m(p, p.Lb, m)[i];
                        hot_pack_column(
                            X, col_sm(p, p.X_dK, scheme, m), i, dK, r + HK_X);
                        hot_pack_column(col_m(p, p.Omega, m),
                            col_m(p, p.Omega_dK, m), i, dK, r + HK_OMEGA);
                        hot_pack_column(col_m(p, p.CS, m), col_m(p, p.CS_dK, m),
                            i, dK, r + HK_CS);
                        hot_pack_column(col_Ms1(p, p.Ms1, scheme, m),
                            col_Ms1(p, p.Ms1_dK, scheme, m), i, dK,
                            r + HK_MS1);
                        hot_pack_column(col_sm(p, p.dE, scheme, m),
                            col_sm(p, p.dE_dK, scheme, m), i, dK, r + HK_DE);
                        r[HK_DLB] =
                            col_m(p, p.Lb, m)[i + 1] - col_m(p, p.Lb, m)[i];

                        double * q =
                            hotX + ((size_t)m * nc + i) * HOTX_WORDS;
                        for (int k = 0; k < HOTX_WORDS; k++) q[k] = 0.;
                        const double dX = X[i + 1] - X[i];
                        q[0] = X[i];
                        q[1] = X[i + 1];
                        q[2] = dX;
                        hot_pack_column(p.K, col_sm(p, p.K_dX, scheme, m), i,
                            dX, q + HX_K);
                }
        }
}

} /* namespace pb */

#endif
