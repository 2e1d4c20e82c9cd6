/*
 * pb_polar.h -- polar angle of the projectile in a discrete energy loss.
 *
 * Restates src/pumas.c:11736-11976 and the minimiser it relies on,
 * src/pumas.c:12105-12325 (a19 of SURVEY.md section 8a). All samplers return
 * mu = (1 - cos theta) / 2.
 */
#ifndef PB_POLAR_H
#define PB_POLAR_H

#include "pb_dcs.h"

namespace pb {

/* Bremsstrahlung (also used for pair production, pumas.c:11841): Tsai's DDCS,
 * two stage rejection. pumas.c:11759-11827. */
template <class Rng>
PB_NOINLINE double polar_bremsstrahlung(
    const Physics & p, Rng & rng, const Element & el, double ki, double kf)
{
        const double m = p.mass;
        const double E = ki + m;
        const double nu = ki - kf;
        const double y = nu / E;

        const double tmp = 0.5 * m * nu / (E * (E - nu));
        const double mu0 = tmp * tmp;
        const double c1 = (2 * (1. - y) + y * y) * mu0;
        const double c2 = 4 * (1. - y) * mu0 * mu0;

        const double RN = el.nuclear_radius;
        const double tmp2 = m * RN / HBAR_C;
        const double muni = tmp2 * tmp2 / 6.;
        const double muc = sqrt(mu0) - mu0;
        if (muc < 0) return 0;

        double x0;
        {
                const double q = muni;
                const double r = mu0;
                const double qr = q * r;
                x0 = (1. + 2 * qr) * xlog((1. + qr) / (r + qr)) -
                    (1. - r) * (1 + 2 * q) / (1 + q);
        }

        for (int i = 0; i < 100; i++) {
                const double r0 = rng.uniform();
                const double r1 = rng.uniform();

                const double mu = r0 * mu0 / (mu0 + 1. - r0);

                const double mus = mu0 + mu;
                const double mu2 = mus * mus;
                const double d1 = 1 / mu2;
                const double d2 = d1 * d1;
                const double score = r1 * c1 * d1;

                double pdf = c1 * d1 - mu * c2 * d2;
                if (score > pdf) continue;

                const double l2 = mu2 / (mu0 * mu0);
                const double q = muni * l2;
                const double r = mu0 * l2;
                const double qr = q * r;
                if (qr > 1E+05) continue;

                const double x = (1. + 2 * qr) * xlog((1. + qr) / (r + qr)) -
                    (1. - r) * (1 + 2 * q) / (1 + q);

                pdf *= x / x0;
                if (score <= pdf) return mu;
        }
        return 0.;
}

/* Objective of the photonuclear maximisation, pumas.c:11855-11862. */
struct PhotonuclearObjective {
        const Physics * p;
        const Element * el;
        double ml, K, q;
        PB_FN double operator()(double lnQ2) const
        {
                const double Q2 = xexp(lnQ2);
                return -dcs_photonuclear_d2(*p, *el, ml, K, q, Q2) * Q2;
        }
};

/* Brent's minimiser as written in the reference (the bracket it refines is
 * [xa, xb] with x starting at xb). pumas.c:12233-12325. */
template <class F>
PB_FN int math_find_minimum_brent(const F & f, double xa, double xb, double xc,
    double fb, double tol, int max_iter, double * xopt, double * fopt)
{
        const double EPS = 1E-15;
        const double GOLD = 0.3819660;

        tol *= fabs(xc - xa);

        double x = xb, w = xb, v = xb;
        double fw = fb, fv = fb, fx = fb;
        double d = 0., e = 0.;
        int i;
        for (i = 0; i < max_iter; i++) {
                const double xm = 0.5 * (xa + xb);
                const double tol1 = tol * fabs(x) + EPS;
                const double tol2 = 2 * tol1;
                if (fabs(x - xm) <= (tol2 - 0.5 * (xb - xa))) break;
                if (fabs(e) > tol1) {
                        const double r = (x - w) * (fx - fv);
                        double q = (x - v) * (fx - fw);
                        double pp = (x - v) * q - (x - w) * r;
                        q = 2. * (q - r);
                        if (q > 0.) pp = -pp;
                        q = fabs(q);
                        const double etemp = e;
                        e = d;
                        if (fabs(pp) >= fabs(0.5 * q * etemp) ||
                            (pp <= q * (xa - x)) || (pp >= q * (xb - x))) {
                                e = (x >= xm) ? xa - x : xb - x;
                                d = GOLD * e;
                        } else {
                                d = pp / q;
                                const double u = x + d;
                                if ((u - xa < tol2) || (xb - u < tol2)) {
                                        d = ((xm - x) >= 0.0) ? fabs(tol1) :
                                                                -fabs(tol1);
                                }
                        }
                } else {
                        e = (x >= xm) ? xa - x : xb - x;
                        d = GOLD * e;
                }
                const double u = (fabs(d) >= tol1) ?
                    x + d :
                    x + ((d >= 0.0) ? fabs(tol1) : -fabs(tol1));
                const double fu = f(u);
                if (fu <= fx) {
                        if (u >= x) {
                                xa = x;
                        } else {
                                xb = x;
                        }
                        v = w, w = x, x = u;
                        fv = fw, fw = fx, fx = fu;
                } else {
                        if (u < x) {
                                xa = u;
                        } else {
                                xb = u;
                        }
                        if ((fu <= fw) || (w == x)) {
                                v = w;
                                w = u;
                                fv = fw;
                                fw = fu;
                        } else if ((fu <= fv) || (v == x) || (v == w)) {
                                v = u;
                                fv = fu;
                        }
                }
        }
        *xopt = x;
        *fopt = fx;
        return (i < max_iter) ? i : -1;
}

/* Bisection bracketing then Brent. pumas.c:12105-12174 with algo = 1 and no
 * precomputed end values. */
template <class F>
PB_FN int math_find_minimum(const F & f, double xa, double xc, double tol,
    int max_iter, double * xopt, double * fopt)
{
        double fa = f(xa);
        double fc = f(xc);
        const double atol = tol * fabs(xc - xa);

        double xb = 0, fb = 0;
        int i;
        for (i = 0; i < max_iter; i++) {
                xb = 0.5 * (xa + xc);
                fb = f(xb);
                if ((fb < fa) && (fb < fc)) break;

                if (fabs(xc - xa) <= atol) {
                        *xopt = xb;
                        *fopt = fb;
                        return i;
                }
                if (fb < fc) {
                        xc = xb;
                        fc = fb;
                } else {
                        xa = xb;
                        fa = fb;
                }
        }
        if (i == max_iter) {
                if (fa < fb) {
                        fb = fa;
                        xb = xa;
                } else if (fc < fb) {
                        fb = fc;
                        xb = xc;
                }
                *xopt = xb;
                *fopt = fb;
                return -1;
        }

        const int status = math_find_minimum_brent(
            f, xa, xb, xc, fb, tol, max_iter - i, xopt, fopt);
        if (status < 0) {
                return -2;
        } else {
                return status + i;
        }
}

/* pumas.c:11874-11950 */
template <class Rng>
PB_NOINLINE double polar_photonuclear(
    const Physics & p, Rng & rng, const Element & el, double ki, double kf)
{
        const double M = 0.5 * (PROTON_MASS + NEUTRON_MASS);
        const double q = ki - kf;
        const double ml = p.mass;
        const double E = ki + ml;
        const double ml2 = ml * ml;
        const double Q2min = ml2 * (q * q - 0.5 * ml2) / (E * (E - q));
        const double Q2max = 2.0 * M * (q - PION_MASS) - PION_MASS * PION_MASS;
        if ((Q2max < Q2min) | (Q2min < 0)) return 0.;

        PhotonuclearObjective objective;
        objective.el = &el;
        objective.ml = ml;
        objective.p = &p;
        objective.K = ki;
        objective.q = q;

        double fopt = 0., xopt = 0.;
        const double lnQ2min = xlog(Q2min);
        const double lnQ2max = xlog(Q2max);
        const int status = math_find_minimum(
            objective, lnQ2min, lnQ2max, 1E-06, 100, &xopt, &fopt);
        fopt = -fopt;
        if (status == -2) return 0.;

        const double x = xexp(xopt) / (2 * q * M);
        const double lne = 0.1;
        if ((x <= 0.04) && (xopt + lne < lnQ2max)) {
                double f2 = 0., x2 = 0.;
                const int status2 = math_find_minimum(
                    objective, xopt + lne, lnQ2max, 1E-06, 100, &x2, &f2);
                if (status2 >= 0) {
                        f2 = -f2;
                        if (f2 > fopt) {
                                xopt = x2;
                                fopt = f2;
                        }
                }
        }

        const double rQ2 = lnQ2max - lnQ2min;
        double Q2 = 0.;
        int i;
        for (i = 0; i < 100; i++) {
                const double u = rng.uniform();
                Q2 = Q2min * xexp(rQ2 * u);
                const double r =
                    dcs_photonuclear_d2(p, el, ml, ki, q, Q2) * Q2;
                if (rng.uniform() * fopt <= r) break;
        }
        if (i == 100) return 0.;

        const double pm = sqrt(ki * (ki + 2 * ml));
        const double E1 = E - q;
        const double eps = ml / E1;
        const double p1 = E1 * sqrt((1. + eps) * (1. - eps));
        const double p2 = pm * p1;
        double tmp = p2 + ml * ml - E * E1;
        if (fabs(tmp) <= 3 * DBL_EPSILON * p2) tmp = 0.;
        const double a_mu = 0.5 * tmp / p2;
        const double b_mu = 0.25 / p2;
        double mu = a_mu + b_mu * Q2;
        if (mu < 0)
                mu = 0.;
        else if (mu > 1.)
                mu = 1.;
        return mu;
}

/* pumas.c:11965-11976 */
PB_FN double polar_ionisation(const Physics & p, double ki, double kf)
{
        const double nu = ki - kf;
        const double p2 = ki * (ki + 2 * p.mass);
        const double E = ki + p.mass;
        const double c = (p2 - nu * (E + ELECTRON_MASS)) /
            sqrt(p2 * (p2 + nu * nu - 2 * nu * E));
        return 0.5 * (1. - c);
}

/* polar_get dispatch, pumas.c:11736-11741. */
template <class Rng>
PB_FN double polar_sample(const Physics & p, Rng & rng, int process,
    const Element & el, double ki, double kf)
{
        if (process <= 1)
                return polar_bremsstrahlung(p, rng, el, ki, kf);
        else if (process == 2)
                return polar_photonuclear(p, rng, el, ki, kf);
        else
                return polar_ionisation(p, ki, kf);
}

} /* namespace pb */

#endif
