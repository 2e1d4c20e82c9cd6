/*
 * pb_coulomb.h -- single (hard) elastic Coulomb scattering.
 *
 * Restates src/pumas.c:7203-7683 (a14 of SURVEY.md section 8a): screening
 * parameters, nuclear form factor, CM frame parameters, pole reduction of the
 * screened DCS and its restricted cross-section.
 *
 * Extended precision. The reference evaluates the pole reduction and the
 * transport coefficients in x87 `long double` (pumas.c:7524, 7623). Their only
 * consumer in the transport is the choice of the struck element
 * (pumas.c:5816-5826); the sampled angle uses binary64 data. The functions are
 * therefore templated on the real type `R`: the CPU debugging build uses
 * `long double` (same as the reference), the kernels use the double-double
 * type of pb_dd.h (106 bit significand, more than the 64 bit of x87), and
 * single element materials skip the computation altogether since the choice
 * is forced (the uniform variate is still drawn).
 */
#ifndef PB_COULOMB_H
#define PB_COULOMB_H

#include "pb_dd.h"
#include "pb_physics.h"

namespace pb {

constexpr int MAX_SCREENING = 4; /* 3 atomic terms + 1 nuclear */

/* Scattering data of one element at one kinetic energy: the double precision
 * members of struct coulomb_data (pumas.c:355-379). */
struct CoulombData {
        int n_parameters;
        double amplitude[3];
        double screening[MAX_SCREENING];
        double fspin;
        double fCM[2];
};

/* pumas.c:7492-7508 */
PB_FN void coulomb_frame_parameters(double Z, double A, double mass,
    double kinetic, double * kinetic0, double * parameters)
{
        const double Ma = Z * PROTON_MASS + (A - Z) * NEUTRON_MASS;
        double M2 = mass + Ma;
        M2 *= M2;
        const double sCM12i = 1. / sqrt(M2 + 2. * Ma * kinetic);
        parameters[0] = (kinetic + mass + Ma) * sCM12i;
        *kinetic0 = (kinetic * Ma + mass * (mass + Ma)) * sCM12i - mass;
        if (*kinetic0 < 1E-09) *kinetic0 = 1E-09;
        const double etot = kinetic + mass + Ma;
        const double betaCM2 = kinetic * (kinetic + 2. * mass) / (etot * etot);
        double rM2 = mass / Ma;
        rM2 *= rM2;
        parameters[1] = sqrt(rM2 * (1. - betaCM2) + betaCM2);
}

/* pumas.c:7475-7479 */
PB_FN double coulomb_spin_factor(double mass, double kinetic)
{
        const double e = kinetic + mass;
        return kinetic * (e + mass) / (e * e);
}

/* pumas.c:7203-7387, with the Z-indexed tables folded into `Element`. */
PB_FN void coulomb_screening_parameters(const Element & el, double mass,
    double kinetic, double kinetic0, int * n_parameters, double * amplitude,
    double * screening)
{
        const double Z = el.Z;
        /* KTV correction to the first Born approximation. */
        const double aZE = ALPHA_EM * Z * (kinetic + mass);
        const double p2 = kinetic * (kinetic + 2. * mass);
        const double zeta2 = aZE * aZE / p2;
        double f;
        if (zeta2 < 1) {
                f = 1. / (1. + zeta2) + 0.20205690315959424 -
                    0.03692775514336999 * zeta2 +
                    0.008349277381922926 * zeta2 * zeta2;
                f *= zeta2;
        } else {
                const double r2i = 1. / (1. + zeta2);
                const double phi = -xatan(sqrt(zeta2));
                f = -0.5 * xlog(r2i) + 0.5 * (1. - r2i) +
                    1. / 12. * (1. - xcos(2 * phi) * r2i) -
                    1. / 120. * (1. - xcos(4 * phi) * r2i * r2i) +
                    1. / 252. * (1. - xcos(6 * phi) * r2i * r2i * r2i);
        }
        const double ktv = xexp(2 * f);

        const double p02 = kinetic0 * (kinetic0 + 2. * mass);
        const double d = 0.25 * HBAR_C * HBAR_C / p02 * ktv;
        int n = el.n_screening;
        int i;
        for (i = 0; i < 3; i++) {
                if (i < n) {
                        amplitude[i] = el.amplitude[i];
                        const double a = el.alpha_over_a0[i];
                        screening[i] = d * a * a;
                }
        }
        const double RN = el.nuclear_radius;
        screening[n++] = 12. * d / (RN * RN);
        *n_parameters = n;
}

/* pumas.c:7396-7416 */
PB_FN double coulomb_nuclear_form_factor(double mu, double N)
{
        if (mu <= 0.)
                return 1.;
        else if (mu > 1.)
                return 0.;
        const double x2 = 10. * mu / N;
        double d;
        if (x2 <= 1E-04) {
                d = 1. / (1. + 0.1 * x2);
        } else {
                const double x = sqrt(x2);
                /* (one argument reduction for both; same values) */
                const SinCos sc = xsincos(x);
                d = 3 * (sc.s - x * sc.c) / (x2 * x);
        }
        return d * d;
}

/* pumas.c:7428-7435 */
PB_FN double coulomb_normalisation(
    double Z, double A, double mass, double kinetic, double kinetic0)
{
        const double p2 = kinetic * (kinetic + 2. * mass);
        const double p02 = kinetic0 * (kinetic0 + 2. * mass);
        double d = ALPHA_EM * Z * HBAR_C * (kinetic + mass);
        return A * 1E-03 * p2 * p02 / (d * d * PI * AVOGADRO_NUMBER);
}

/* Fill the binary64 scattering data of one element. */
PB_FN void coulomb_data_set(
    const Physics & p, const Element & el, double kinetic, CoulombData * data,
    double * kinetic0)
{
        coulomb_frame_parameters(el.Z, el.A, p.mass, kinetic, kinetic0,
            data->fCM);
        data->fspin = coulomb_spin_factor(p.mass, kinetic);
        coulomb_screening_parameters(el, p.mass, kinetic, *kinetic0,
            &data->n_parameters, data->amplitude, data->screening);
}

/* Logarithm in the working precision R. */
PB_FN double real_log(double x) { return xlog(x); }
PB_FN dd real_log(const dd & x) { return dd_log(x); }
#if !defined(__CUDA_ARCH__)
inline long double real_log(long double x) { return logl(x); }
#endif
#if defined(PB_FORCE_QUAD)
extern "C" __float128 logq(__float128);
inline __float128 real_log(__float128 x) { return logq(x); }
#endif

/* Coefficients of the pole reduction, pumas.c:355-379 (a, b, c). */
template <class R> struct CoulombPoles {
        R a[3];
        R b[3];
        R c[4];
};

/* pumas.c:7524-7576 */
template <class R>
PB_FN void coulomb_pole_reduction(int n_parameters, const double * amplitude,
    const double * screening, CoulombPoles<R> * poles)
{
        R * a = poles->a;
        R * b = poles->b;
        R * c = poles->c;
        const int n = n_parameters - 1;
        int i, j;
        for (i = 0; i < 3; i++) a[i] = b[i] = R(0.);
        for (i = 0; i < 4; i++) c[i] = R(0.);

        for (i = 0; i < n; i++) {
                const R Ai = R(amplitude[i]);
                const R Bi = R(screening[i]);
                b[i] = Ai * Ai;
                for (j = i + 1; j < n; j++) {
                        const R Aj = R(amplitude[j]);
                        const R Bj = R(screening[j]);
                        const R d = R(2.) * Ai * Aj / (Bj - Bi);
                        a[i] = a[i] + d;
                        a[j] = a[j] - d;
                }
        }

        const R N = R(screening[n]);
        R Sa[4] = { R(0.), R(0.), R(0.), R(0.) };
        R Sb[4] = { R(0.), R(0.), R(0.), R(0.) };
        for (i = 0; i < n; i++) {
                const R Bi = R(screening[i]);
                const R x = R(1.) - Bi / N;
                R r = x;
                for (j = 0; j < 4; j++) {
                        Sa[j] = Sa[j] + a[i] / r;
                        r = r * x;
                        Sb[j] = Sb[j] + b[i] / r;
                }
        }

        R tmp = R(1.);
        for (i = 0; i < 4; i++) {
                c[i] = (R((double)(4 - i)) * Sb[3 - i] / N - Sa[3 - i]) * tmp;
                tmp = tmp * N;
        }

        for (i = 0; i < n; i++) {
                const R Bi = R(screening[i]);
                R d = R(1.) - Bi / N;
                d = d * d;
                d = d * d;
                a[i] = (a[i] - R(4.) * b[i] / (N - Bi)) / d;
                b[i] = b[i] / d;
        }
}

/* Order 0 transport coefficient (the cross-section integrated up to mu).
 * pumas.c:7623-7683; the order 1 coefficient is not needed at run time. The
 * accumulator is binary64 and each term is rounded when added, as in the
 * reference (`double cs0; cs0 += <long double expression>`). */
template <class R>
PB_FN double coulomb_transport_cs0(double mu_, double fspin, int n_parameters,
    const double * screening, const CoulombPoles<R> & poles)
{
        const R S = R(fspin);
        const R mu = R(mu_);
        const int n = n_parameters - 1;
        double cs0 = 0.;
        int i;
        for (i = 0; i < n; i++) {
                const R alp = R(screening[i]);
                const R r = mu / (mu + alp);
                const R L = real_log(R(1.) + mu / alp);
                const R I0 = r / alp;
                const R J0 = L;
                const R I1 = L - r;
                const R J1 = -alp * L;
                cs0 = (double)(R(cs0) +
                    (poles.a[i] * (J0 - S * J1) + poles.b[i] * (I0 - S * I1)));
        }

        const R N = R(screening[n]);
        const R r = mu / (mu + N);
        const R L = real_log(R(1.) + mu / N);
        const R I0 = r / N;
        const R J0 = L;
        const R I1 = L - r;
        const R J1 = -N * L;

        const R rn = R(1.) / (R(1.) + mu / N);
        const R K0 = (R(1.) - rn) * (R(1.) + rn) / (R(2.) * N * N);
        const R L0 = (R(1.) - rn) * ((R(1.) + rn) * (R(1.) + rn) - rn) /
            (R(3.) * N * N * N);
        const R K1 = I0 - N * K0;
        const R L1 = K0 - N * L0;

        cs0 = (double)(R(cs0) +
            (poles.c[0] * (J0 - S * J1) + poles.c[1] * (I0 - S * I1) +
                poles.c[2] * (K0 - S * K1) + poles.c[3] * (L0 - S * L1)));
        return cs0;
}

/* pumas.c:7591-7609 */
template <class R>
PB_FN double coulomb_restricted_cs(double mu, double fspin, int n_parameters,
    const double * screening, const CoulombPoles<R> & poles)
{
        if ((mu >= 1.) || (mu >= 1E+06 * screening[n_parameters - 1]))
                return 0.;
        const double cs0 =
            coulomb_transport_cs0<R>(1., fspin, n_parameters, screening, poles);
        const double cs1 =
            coulomb_transport_cs0<R>(mu, fspin, n_parameters, screening, poles);
        if (cs1 > cs0)
                return 0;
        else
                return cs0 - cs1;
}

} /* namespace pb */

#endif
