/*
 * user_dcs.h -- a user differential cross-section for pumas_dcs_register
 * (pumas.h:2150-2167): the bremsstrahlung of Bethe & Heitler with complete
 * screening, Petrukhin & Shestakov's nuclear size correction left out. Not a
 * recommended model: a closed form that differs from the built-in ones, to
 * exercise the extension point. Written once; compiled as a host function (for
 * the table generation of the reference and the host API) and as a `__device__`
 * function (for the transport kernels, examples/user_media.cu).
 */
#ifndef USER_DCS_H
#define USER_DCS_H

#include <math.h>

#ifndef USER_DCS_FN
#define USER_DCS_FN static inline
#endif

#define USER_DCS_MODEL "BHS"

/* d(sigma)/dq [m^2 / GeV] for a projectile of mass m [GeV] and kinetic energy
 * K [GeV] losing q [GeV] to a photon on a nucleus (Z, A). */
USER_DCS_FN double user_dcs_bremsstrahlung(
    double Z, double A, double m, double K, double q)
{
        (void)A;
        const double me = 0.510998910E-03; /* electron mass [GeV] */
        const double re = 2.817940285E-15; /* classical electron radius [m] */
        const double alpha = 7.2973525376E-03;
        const double E = K + m;
        if ((q <= 0.) || (q >= K)) return 0.;
        const double v = q / E;
        const double z13 = cbrt(Z);
        /* complete screening, with the minimum momentum transfer
         * delta = m^2 v / (2 E (1 - v)) */
        const double delta = 0.5 * m * m * v / (E * (1. - v));
        const double sqrte = 1.6487212707001282; /* exp(1 / 2) */
        const double phi = log(189. * m / (me * z13) /
            (1. + 189. * sqrte * delta / (me * z13)));
        if (phi <= 0.) return 0.;
        const double r = re * me / m;
        return 4. * alpha * Z * (Z + 1.) * r * r *
            (4. / 3. * (1. - v) + v * v) * phi / q;
}

#endif
