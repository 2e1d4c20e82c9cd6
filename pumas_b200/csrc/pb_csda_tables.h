/*
 * pb_csda_tables.h -- packed tables and O(1) locators of the closed form CSDA
 * (pb_csda.h), built once per physics object by the dump loader (pb_dump.hpp).
 */
#ifndef PB_CSDA_TABLES_H
#define PB_CSDA_TABLES_H

#include <cstring>

#include "pb_hot_tables.h"

namespace pb {

/* K record of cell i (scheme CSDA, material m):
 *   [0] K[i] [1] K[i+1] [2..3] unused [4..7] X [8..11] T [12..15] dE
 * X record: [0] X[i] [1] X[i+1] [2..3] unused [4..7] K. */
constexpr int CSDAK_WORDS = 16;
constexpr int CSDAX_WORDS = 8;
constexpr int CK_X = 4, CK_T = 8, CK_DE = 12, CX_K = 4;

/* O(1) locator of a strictly increasing table: key = the top bits of the
 * double; first[key - key_min] = first cell that can hold a value of that
 * key. */
constexpr int LOCATOR_SHIFT = 46; /* sign, exponent, 6 mantissa bits */
struct Locator {
        int key_min, n_keys;
        const unsigned short * first; /* [n_keys] */
};

inline size_t csda_tables_words(const Physics & p)
{
        return (size_t)p.n_materials * (size_t)(p.n_energies - 1) *
            (CSDAK_WORDS + CSDAX_WORDS);
}

/* Packed CSDA tables: K records of every material, then X records. */
inline void csda_tables_build(const Physics & p, double * cK, double * cX)
{
        const int nc = p.n_energies - 1;
        const int scheme = MODE_CSDA;
        for (int m = 0; m < p.n_materials; m++) {
                const double * X = col_sm(p, p.X, scheme, m);
                for (int i = 0; i < nc; i++) {
                        double * r = cK + ((size_t)m * nc + i) * CSDAK_WORDS;
                        for (int k = 0; k < CSDAK_WORDS; k++) r[k] = 0.;
                        const double dK = p.K[i + 1] - p.K[i];
                        r[0] = p.K[i];
                        r[1] = p.K[i + 1];
                        hot_pack_column(
                            X, col_sm(p, p.X_dK, scheme, m), i, dK, r + CK_X);
                        hot_pack_column(col_sm(p, p.T, scheme, m),
                            col_sm(p, p.T_dK, scheme, m), i, dK, r + CK_T);
                        hot_pack_column(col_sm(p, p.dE, scheme, m),
                            col_sm(p, p.dE_dK, scheme, m), i, dK, r + CK_DE);
                        double * q = cX + ((size_t)m * nc + i) * CSDAX_WORDS;
                        for (int k = 0; k < CSDAX_WORDS; k++) q[k] = 0.;
                        const double dX = X[i + 1] - X[i];
                        q[0] = X[i];
                        q[1] = X[i + 1];
                        hot_pack_column(p.K, col_sm(p, p.K_dX, scheme, m), i,
                            dX, q + CX_K);
                }
        }
}

PB_FN int locator_key(double x)
{
        uint64_t b;
        memcpy(&b, &x, sizeof b);
        return (int)(b >> LOCATOR_SHIFT);
}

/* Keys spanned by the positive nodes of a table, and the table of first
 * cells: first[k] = the cell of the smallest value of key k (cell of the node
 * range [table[i], table[i + 1]) holding it). */
inline void locator_range(
    const double * table, int n_nodes, int * key_min, int * n_keys)
{
        int first_positive = 0;
        while ((first_positive < n_nodes - 1) && !(table[first_positive] > 0.))
                first_positive++;
        *key_min = locator_key(table[first_positive]);
        *n_keys = locator_key(table[n_nodes - 1]) - *key_min + 1;
}

inline void locator_build(const double * table, int n_nodes, int key_min,
    int n_keys, unsigned short * first)
{
        int cell = 0;
        for (int k = 0; k < n_keys; k++) {
                /* smallest double of the key */
                const uint64_t b = (uint64_t)(key_min + k) << LOCATOR_SHIFT;
                double lo;
                memcpy(&lo, &b, sizeof lo);
                while ((cell < n_nodes - 2) && (lo >= table[cell + 1])) cell++;
                first[k] = (unsigned short)cell;
        }
}

} /* namespace pb */

#endif
