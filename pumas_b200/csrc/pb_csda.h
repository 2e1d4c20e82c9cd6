/*
 * pb_csda.h -- the closed form CSDA transport of one track in a uniform medium
 * (transport_with_csda, src/pumas.c:4777-4949, behind the prologue of
 * pumas_context_transport, pumas.c:3114-3235), in the lean form the CSDA
 * kernel runs (kernels_csda.cu; BASELINE.json configs[1]).
 *
 * Same results, bit for bit, as the general restatement (pb_transport.h), with
 * the two costs of a track removed:
 *  - the searches. The reference bisects the energy grid for every look-up (8
 *    dependent loads each, ~6 look-ups per track). Here a value is located in
 *    O(1): the top bits of the double (sign, exponent, a few mantissa bits)
 *    index a small table of first cells (`Locator`), followed by a walk of
 *    0-2 cells; on a strictly increasing grid that is the cell the bisection
 *    finds (pumas.c:4174-4233);
 *  - the gathers. The columns a track needs (X, T, dE over K; K over X) are
 *    packed per grid cell as the four PCHIP coefficients
 *    (pb_hot_tables.h: hot_pack_column), one 16 byte aligned record per cell.
 * Anything unusual (decayed state, bad direction, a time limit, energies off
 * the grid) is detected before the state is touched and handed to the general
 * transport_track.
 */
#ifndef PB_CSDA_H
#define PB_CSDA_H

#include "pb_csda_tables.h"
#include "pb_hot.h"

namespace pb {

/* ---- per launch constants --------------------------------------------------- */
struct CsdaCtx {
        HotPtr cK, cX;          /* records of the medium's material */
        const double * gK, * gX; /* the same in HBM */
        Locator locK, locX;
        int n_cells, material;
        double K0, Kmax, X0, Xmax;
        double density, mass, ctau;
        double X_limit_energy; /* cel_grammage(limit.energy), pumas.c:4901 */
};

/* `host_view` is read here; `view` holds the pointers the track function will
 * use (device or host). */
template <class Cfg>
inline CsdaCtx csda_context(const Physics & host_view, const Physics & view,
    const Cfg & cfg, int material, double density = 0.)
{
        CsdaCtx c;
        const Physics & p = host_view;
        const int nc = p.n_energies - 1;
        c.gK = view.csdaK + (size_t)material * nc * CSDAK_WORDS;
        c.gX = view.csdaX + (size_t)material * nc * CSDAX_WORDS;
#if defined(__CUDACC__) && defined(PB_HOT_SMEM_TABLES)
        c.cK = c.cX = 0; /* (set by the launcher: shared memory layout) */
#else
        c.cK = c.gK;
        c.cX = c.gX;
#endif
        c.locK.key_min = p.locK_min;
        c.locK.n_keys = p.locK_keys;
        c.locK.first = view.locK;
        c.locX.key_min = p.locX_min;
        c.locX.n_keys = p.locX_keys;
        c.locX.first = view.locX + (size_t)material * p.locX_keys;
        c.n_cells = nc;
        c.material = material;
        c.K0 = p.K[0];
        c.Kmax = p.K[nc];
        const double * X = col_sm(p, p.X, MODE_CSDA, material);
        c.X0 = X[0];
        c.Xmax = X[nc];
        c.density = (density > 0.) ? density : p.material_density[material];
        c.mass = p.mass;
        c.ctau = p.ctau;
        c.X_limit_energy = (cfg.event & EVENT_LIMIT_ENERGY) ?
            cel_grammage(p, MODE_CSDA, material, cfg.limit_energy) :
            0.;
        return c;
}

/* Cell of x, K0 <= x < Kmax (the caller checks). */
PB_FN HLoc csda_locate(HotPtr records, int words, const Locator & loc,
    int n_cells, double x)
{
        int k = locator_key(x) - loc.key_min;
        if (k < 0) k = 0;
        if (k > loc.n_keys - 1) k = loc.n_keys - 1;
#if defined(__CUDA_ARCH__)
        int i = (int)__ldg(loc.first + k);
#else
        int i = (int)loc.first[k];
#endif
        if (i > n_cells - 1) i = n_cells - 1;
        D2 b = hot_ld2(records + i * words);
        while (x >= b.y) {
                ++i;
                b = hot_ld2(records + i * words);
        }
        while (x < b.x) {
                --i;
                b = hot_ld2(records + i * words);
        }
        HLoc l;
        l.i = i;
        l.t = (x - b.x) / (b.y - b.x);
        return l;
}

/*
 * One track. Returns false, with the state untouched, when the general
 * transport_track must run instead. `out` as transport_track's.
 */
template <class Cfg>
PB_FN bool csda_track(const Cfg & cfg, const CsdaCtx & c, State & state,
    Outcome & out)
{
        static_assert(Cfg::energy_loss == MODE_CSDA, "CSDA only");
        static_assert(Cfg::scattering == MODE_DISABLED, "no scattering");
        static_assert(Cfg::decay == MODE_WEIGHTED, "weighted decays only");
        constexpr bool forward = (Cfg::direction == MODE_FORWARD);

        /* pumas_context_transport prologue, pumas.c:3135-3150: the unusual
         * cases go the general way */
        if (state.decayed || (cfg.event & EVENT_LIMIT_TIME)) return false;
        {
                const double * const u = state.direction;
                const double norm2 = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
                if (fabs(norm2 - 1) > (double)FLT_EPSILON) return false;
        }
        const double ki = state.energy;
        if (!((ki >= c.K0) && (ki < c.Kmax))) return false;

        /* transport_with_csda, pumas.c:4777-4949 */
        const double density = c.density;
        const double di = state.distance;
        const double ti = state.time;
        const HLoc li =
            csda_locate(c.cK, CSDAK_WORDS, c.locK, c.n_cells, ki);
        const HotPtr ri = c.cK + li.i * CSDAK_WORDS;
        const double xi = hot_pchip(ri + CK_X, li.t);
        const double Ti = hot_pchip(ri + CK_T, li.t);

        int event = EVENT_NONE;
        double xB;
        if (cfg.event & EVENT_LIMIT_GRAMMAGE) {
                xB = cfg.limit_grammage - state.grammage;
                event = EVENT_LIMIT_GRAMMAGE;
        } else
                xB = DBL_MAX;
        if (cfg.event & EVENT_LIMIT_DISTANCE) {
                const double xD = density * (cfg.limit_distance - di);
                if (xD < xB) {
                        xB = xD;
                        event = EVENT_LIMIT_DISTANCE;
                }
        }
        out.rc = RETURN_SUCCESS;
        out.medium[0] = out.medium[1] = 0;
        if (xB <= 0.) {
                out.event = event;
                return true;
        }

        double xf, kf;
        bool have_lf = false; /* kf comes from ki's cell data only */
        if (forward) {
                if (xi > xB) {
                        xf = xi - xB;
                        if (xB < EPSILON_X * xi) {
                                kf = ki - hot_pchip(ri + CK_DE, li.t) * xB;
                                if (kf < 0.) kf = 0.;
                        } else if (xf < c.X0) {
                                kf = 0.;
                        } else {
                                const HLoc lx = csda_locate(c.cX, CSDAX_WORDS,
                                    c.locX, c.n_cells, xf);
                                kf = hot_pchip(
                                    c.cX + lx.i * CSDAX_WORDS + CX_K, lx.t);
                        }
                } else {
                        xf = kf = 0.;
                        event = EVENT_LIMIT_ENERGY;
                }
                if (cfg.event & EVENT_LIMIT_ENERGY) {
                        if (ki <= cfg.limit_energy) {
                                out.event = EVENT_LIMIT_ENERGY;
                                return true;
                        }
                        if (kf < cfg.limit_energy) {
                                kf = cfg.limit_energy;
                                xf = c.X_limit_energy;
                                event = EVENT_LIMIT_ENERGY;
                        }
                }
        } else {
                if (xB == DBL_MAX) {
                        xf = xB;
                        kf = DBL_MAX;
                } else {
                        xf = xB + xi;
                        if (xB < EPSILON_X * xi) {
                                kf = ki + hot_pchip(ri + CK_DE, li.t) * xB;
                        } else {
                                if (!(xf < c.Xmax)) return false;
                                const HLoc lx = csda_locate(c.cX, CSDAX_WORDS,
                                    c.locX, c.n_cells, xf);
                                kf = hot_pchip(
                                    c.cX + lx.i * CSDAX_WORDS + CX_K, lx.t);
                        }
                }
                if (cfg.event & EVENT_LIMIT_ENERGY) {
                        if (ki >= cfg.limit_energy) {
                                out.event = EVENT_LIMIT_ENERGY;
                                return true;
                        }
                        if (kf > cfg.limit_energy) {
                                kf = cfg.limit_energy;
                                xf = c.X_limit_energy;
                                event = EVENT_LIMIT_ENERGY;
                        }
                }
        }
        /* the end energy in the grid (proper time, backward weight) */
        if (!((kf >= c.K0) && (kf < c.Kmax))) return false;
        (void)have_lf;
        const HLoc lf =
            csda_locate(c.cK, CSDAK_WORDS, c.locK, c.n_cells, kf);
        const HotPtr rf = c.cK + lf.i * CSDAK_WORDS;

        const double distance = fabs(xf - xi) / density;
        state.energy = kf;
        if (event == EVENT_LIMIT_DISTANCE)
                state.distance = cfg.limit_distance;
        else
                state.distance += distance;
        if (event == EVENT_LIMIT_GRAMMAGE)
                state.grammage = cfg.limit_grammage;
        else
                state.grammage += distance * density;
        state.time += fabs(Ti - hot_pchip(rf + CK_T, lf.t)) / density;
        if (!forward)
                state.weight *= hot_pchip(rf + CK_DE, lf.t) /
                    hot_pchip(ri + CK_DE, li.t);
        state.weight *= xexp(-fabs(ti - state.time) / c.ctau);

        const double path = forward ? distance : -distance;
        state.position[0] += path * state.direction[0];
        state.position[1] += path * state.direction[1];
        state.position[2] += path * state.direction[2];
        out.event = event;
        return true;
}

} /* namespace pb */

#endif
