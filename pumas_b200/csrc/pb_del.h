/*
 * pb_del.h -- sampling of a discrete (hard) inelastic energy loss.
 *
 * Restates src/pumas.c:5887-6270 (a15 of SURVEY.md section 8a): choice of the
 * target element and process, forward sampling under a two power law
 * envelope, backward (adjoint) sampling with its Jacobian weight.
 * The order and number of uniform variates drawn is part of the contract.
 */
#ifndef PB_DEL_H
#define PB_DEL_H

#include "pb_dcs.h"

namespace pb {

/* struct del_info, pumas.c:643-656. In the reference the energy transfer Q and
 * the weight factor are the two members of a UNION: writing the weight
 * (pumas.c:6231) clobbers Q, and the element / process selection loop that
 * follows (pumas.c:6254-6257) evaluates the DCS at q = weight. This is kept:
 * a single storage location plays both roles. */
struct DelInfo {
        double reverse; /* in: energy transfer Q; out: weight factor */
        int process;
        int element;
};

/* pumas.c:5887-5894 */
PB_FN double envelope_norm(double a, double xmin, double xmax)
{
        if (a == 1) {
                return xlog(xmax / xmin);
        } else {
                return (xpow(xmin, 1. - a) - xpow(xmax, 1. - a)) / (a - 1.);
        }
}

/* pumas.c:6154-6270 */
template <class Cfg, class Rng>
PB_FN void del_randomise_target(const Physics & p, const Cfg & cfg,
    Rng & rng, const State & state, int material, DelInfo * info)
{
        int i1, i2;
        double h, dk;
        i1 = table_index(p, p.K, state.energy);
        if (i1 < 0) {
                i1 = i2 = 0;
                h = dk = 0.;
        } else if (i1 >= p.n_energies - 1) {
                i1 = i2 = p.n_energies - 1;
                h = dk = 0.;
        } else {
                i2 = i1 + 1;
                const double K1 = p.K[i1];
                dk = p.K[i2] - K1;
                h = (state.energy - K1) / dk;
        }
        (void)i2;

        const int ic0 = p.comp_offset[material];
        const int ic1 = ic0 + p.elements_in[material];
        int ic, ip;
        if (cfg.direction == MODE_FORWARD) {
                const double zeta = rng.uniform();
                for (ic = ic0; ic < ic1; ic++)
                        for (ip = 0; ip < N_DEL_PROCESSES; ip++) {
                                const size_t k =
                                    (size_t)(ip * p.n_components + ic) *
                                        p.n_energies + i1;
                                const double * f = p.CSf + k;
                                const double * df = p.CSf_dK + k;
                                const double csf = pchip(
                                    h, f[0], f[1], df[0] * dk, df[1] * dk);
                                if (!(zeta > csf)) {
                                        const int iel = p.comp_element[ic];
                                        const size_t kn =
                                            (size_t)(ip * p.n_elements + iel) *
                                                p.n_energies + i1;
                                        const double * n = p.CSn + kn;
                                        const double * dn = p.CSn_dK + kn;
                                        const double csn = pchip(h, n[0], n[1],
                                            dn[0] * dk, dn[1] * dk);
                                        info->reverse = 1. / csn;
                                        info->element = iel;
                                        info->process = ip;
                                        return;
                                }
                        }
                /* Not reached for a cumulative table ending at 1. */
                info->element = p.comp_element[ic1 - 1];
                info->process = N_DEL_PROCESSES - 1;
                return;
        }

        /* Backward: randomise according to the DCS. */
        double stot = 0.;
        for (ip = 0; ip < N_DEL_PROCESSES; ip++) {
                for (ic = ic0; ic < ic1; ic++) {
                        const double d = dcs_evaluate(p, ip, p.comp_element[ic],
                            state.energy, info->reverse);
                        stot += d * p.comp_fraction[ic];
                }
        }
        if (stot == 0.) {
                info->reverse = 0.;
                return;
        } else {
                info->reverse = stot;
        }

        if (cfg.scattering == MODE_DISABLED) {
                info->element = p.n_elements;
                info->process = N_DEL_PROCESSES;
                return;
        }

        double zeta;
        for (;;) {
                zeta = rng.uniform() * stot;
                if ((zeta > 0.) && (zeta < stot)) break;
        }
        double s = 0.;
        for (ic = ic0; ic < ic1; ic++)
                for (ip = 0; ip < N_DEL_PROCESSES; ip++) {
                        const int iel = p.comp_element[ic];
                        const double si = dcs_evaluate(p, ip, iel, state.energy,
                                              info->reverse) *
                            p.comp_fraction[ic];
                        s += si;
                        if (!(zeta > s)) {
                                info->element = iel;
                                info->process = ip;
                                return;
                        }
                }
        info->element = p.comp_element[ic1 - 1];
        info->process = N_DEL_PROCESSES - 1;
}

/* pumas.c:5912-6017 in two parts (a scheduler may regroup the events by
 * process between them; tried and rejected on B200, profiles/r2_experiments.md
 * section 13): the threshold test and the choice of the target (false: below
 * the threshold, no event), then the sampling of the energy transfer for that
 * target. Returns true when a polar angle must be sampled for (*process,
 * *element); false is the reference's NULL polar function. */
template <class Cfg, class Rng>
PB_FN bool del_forward_target(const Physics & p, const Cfg & cfg, Rng & rng,
    const State & state, int material, DelInfo * info)
{
        if (state.energy <= p.Kt[material]) return false;
        info->reverse = 0.;
        info->process = 0;
        info->element = 0;
        del_randomise_target(p, cfg, rng, state, material, info);
        return true;
}

template <class Cfg, class Rng>
PB_FN bool del_forward_sample(const Physics & p, const Cfg & cfg, Rng & rng,
    State & state, const DelInfo & info, int * process, int * element)
{
        (void)cfg;
        const Element & el = p.element[info.element];
        *element = info.element;
        *process = info.process;

        if (info.process == 3) {
                state.energy = dcs_ionisation_randomise(
                    p, rng, el, state.energy, p.cutoff);
                return true;
        }

        double qmin, qmax;
        dcs_get_range(info.process, el, p.mass, state.energy, &qmin, &qmax);
        double xmin = qmin / state.energy;
        if (xmin < p.cutoff) xmin = p.cutoff;
        const double xmax = qmax / state.energy;
        if (xmin >= xmax) return false;

        double a0, a1, b0, b1, p0;
        const int imax = p.n_energies - 1;
        if (state.energy >= p.K[imax]) {
                const float * envelope = p.DCS_envelope +
                    2 * ((info.process * p.n_elements + info.element) *
                                p.n_energies + imax);
                a0 = (double)envelope[0];
                b0 = (double)envelope[1];
                p0 = 1.;
                a1 = b1 = 0.;
        } else {
                const int row = table_index(p, p.K, state.energy);
                const float * envelope0 = p.DCS_envelope +
                    2 * ((info.process * p.n_elements + info.element) *
                                p.n_energies + row);
                const float * envelope1 = envelope0 + 2;

                a0 = (double)envelope0[0];
                b0 = (double)envelope0[1];
                a1 = (double)envelope1[0];
                b1 = (double)envelope1[1];

                const double k0 = p.K[row];
                const double k1 = p.K[row + 1];
                const double h1 = (state.energy - k0) / (k1 - k0);
                const double h0 = 1. - h1;
                b0 *= h0;
                b1 *= h1;

                p0 = envelope_norm(a0, xmin, xmax) * b0;
                const double p1 = envelope_norm(a1, xmin, xmax) * b1;
                p0 /= p0 + p1;
        }

        double x = 0.;
        int i;
        for (i = 0; i < 100; i++) {
                double u = rng.uniform();
                double a;
                if (u < p0) {
                        a = a0;
                        u = u / p0;
                } else {
                        a = a1;
                        u = (u - p0) / (1. - p0);
                }
                if (a == 1) {
                        x = xmin * xexp(u * xlog(xmax / xmin));
                } else {
                        const double am = a - 1.;
                        const double pmin = xpow(xmin, -am);
                        const double pmax = xpow(xmax, -am);
                        x = xpow(pmin - u * (pmin - pmax), -1. / am);
                }

                const double penv = b0 * xpow(x, -a0) + b1 * xpow(x, -a1);
                const double d = dcs_evaluate(p, info.process, info.element,
                    state.energy, state.energy * x);
                if (rng.uniform() * penv <= d) break;
        }

        if (i == 100) {
                return false;
        } else {
                state.energy *= (1. - x);
                return true;
        }
}

template <class Cfg, class Rng>
PB_FN bool del_randomise_forward(const Physics & p, const Cfg & cfg,
    Rng & rng, State & state, int material, int * process, int * element)
{
        DelInfo info;
        if (!del_forward_target(p, cfg, rng, state, material, &info))
                return false;
        return del_forward_sample(p, cfg, rng, state, info, process, element);
}

/* pumas.c:6109-6138 */
template <class Rng>
PB_FN void del_randomise_power_law(Rng & rng, double alpha, double xmin,
    double xmax, double * p_r, double * p_w)
{
        double r = 0., w_bias;
        if (alpha == 1.) {
                const double lnq = xlog(xmax / xmin);
                for (;;) {
                        const double z = rng.uniform();
                        r = xmin * xexp(z * lnq);
                        if ((r < xmin) || (r >= xmax)) continue;
                        w_bias = lnq * r;
                        break;
                }
        } else {
                const double a1 = 1. - alpha;
                const double x0 = xpow(xmin, a1);
                const double x1 = xpow(xmax, a1);
                for (;;) {
                        const double z = rng.uniform();
                        const double tmp = x0 + z * (x1 - x0);
                        r = xpow(tmp, 1. / a1);
                        if ((r < xmin) || (r >= xmax)) continue;
                        w_bias = (x1 - x0) * r / (a1 * tmp);
                        break;
                }
        }
        *p_r = 1. - r;
        *p_w = w_bias;
}

/* pumas.c:6036-6093 */
template <class Cfg, class Rng>
PB_FN bool del_randomise_reverse(const Physics & p, const Cfg & cfg,
    Rng & rng, State & state, int material, int * process, int * element)
{
        const double lnq0 = -xlog(p.cutoff);
        const double kt = p.Kt[material];
        const double kf = state.energy;
        const double pCEL =
            (kf >= kt) ? 0. : lnq0 / (lnq0 + xlog(kt / (kt - kf)));
        if (pCEL && (rng.uniform() < pCEL)) {
                state.weight /= pCEL;
                return false;
        }
        const double xf =
            (kf >= kt * (1. - p.cutoff)) ? p.cutoff : 1. - kf / kt;

        double xmax;
        const double m1 = p.mass - ELECTRON_MASS;
        if (state.energy < 0.5 * m1 * m1 / ELECTRON_MASS) {
                const double m2 = p.mass + ELECTRON_MASS;
                xmax = 2. * ELECTRON_MASS * (state.energy + 2. * p.mass) /
                    (m2 * m2);
                if (xmax < xf)
                        return false;
                else if (xmax > 1.)
                        xmax = 1.;
        } else {
                xmax = 1.;
        }
        const double alpha = BMC_ALPHA;

        double r, w_bias;
        del_randomise_power_law(rng, alpha, xf, xmax, &r, &w_bias);
        w_bias /= r;
        state.energy /= r;

        DelInfo info;
        info.reverse = state.energy - kf;
        info.process = 0;
        info.element = 0;
        del_randomise_target(p, cfg, rng, state, material, &info);
        if (info.reverse <= 0) {
                state.weight = 0.;
                return false;
        }

        state.weight *= w_bias * info.reverse /
            (del_cross_section(p, material, kf) * (1. - pCEL));
        *process = info.process;

        if (cfg.scattering == MODE_DISABLED) {
                return false;
        } else {
                *element = info.element;
                return true;
        }
}

} /* namespace pb */

#endif
