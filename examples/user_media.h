/*
 * user_media.h -- a user geometry the planar descriptors of
 * pumas_b200_geometry.h do not cover: concentric spherical shells around the
 * origin, each with a density that varies linearly with the radius. Written
 * once, compiled twice:
 *   - by examples/user_media.cu as `__device__` functions behind
 *     `struct pumas_b200_callbacks` (the device form of pumas_medium_cb /
 *     pumas_locals_cb, pumas.h:509-511, 348-349);
 *   - by oracle/oracle_driver.c (test infrastructure) as the host callbacks of
 *     the unmodified reference, so that tests/test_user_media.py can compare
 *     the two.
 * Plain C99 / C++.
 */
#ifndef USER_MEDIA_H
#define USER_MEDIA_H

#include <math.h>

#ifndef USER_MEDIA_FN
#define USER_MEDIA_FN static inline
#endif

#define USER_MEDIA_MAX_SHELLS 8

struct user_media {
        int n_shells;
        int reserved;
        /* shell i: r[i] <= |x| < r[i + 1] [m] */
        double r[USER_MEDIA_MAX_SHELLS + 1];
        /* rho(|x|) = density[i] * (1 + gradient[i] * (r[i + 1] - |x|)) */
        double density[USER_MEDIA_MAX_SHELLS];  /* [kg/m^3] */
        double gradient[USER_MEDIA_MAX_SHELLS]; /* [1/m]; 0: uniform */
        double step_epsilon;                    /* [m] */
};

/* Shell of `position` (-1 outside) and distance to the next shell boundary
 * along the motion. Returns 0 (PUMAS_STEP_CHECK): the proposal may be short
 * of, or beyond, the boundary by rounding; the stepping checks. */
USER_MEDIA_FN int user_media_locate(const struct user_media * g,
    const double * position, const double * direction, int backward,
    int * medium, double * step)
{
        const double x = position[0], y = position[1], z = position[2];
        const double r2 = x * x + y * y + z * z;
        const int n = g->n_shells;
        if ((r2 < g->r[0] * g->r[0]) || !(r2 < g->r[n] * g->r[n])) {
                if (medium != 0) *medium = -1;
                if (step != 0) *step = -1.;
                return 0;
        }
        int i = 0;
        while ((i < n - 1) && !(r2 < g->r[i + 1] * g->r[i + 1])) i++;
        if (medium != 0) *medium = i;
        if (step == 0) return 0;

        const double sgn = backward ? -1. : 1.;
        /* |x + s u|^2 = R^2: s^2 + 2 b s + (r2 - R^2) = 0 */
        const double b =
            sgn * (x * direction[0] + y * direction[1] + z * direction[2]);
        double s = -1.;
        if (b < 0.) {
                /* moving inward: the inner sphere, if the line meets it */
                const double ci = r2 - g->r[i] * g->r[i];
                const double di = b * b - ci;
                if (di > 0.) s = -b - sqrt(di);
        }
        if (s <= 0.) {
                const double co = r2 - g->r[i + 1] * g->r[i + 1];
                double d = b * b - co;
                if (d < 0.) d = 0.;
                s = -b + sqrt(d);
        }
        if (s < 0.) s = 0.;
        *step = s + g->step_epsilon;
        return 0;
}

/* Density of shell `medium` at `position`; returns the size of the
 * inhomogeneity [m] (0: uniform), as pumas_locals_cb does. */
USER_MEDIA_FN double user_media_density(const struct user_media * g, int medium,
    const double * position, double * density)
{
        const double a = g->gradient[medium];
        if (a == 0.) {
                *density = g->density[medium];
                return 0.;
        }
        const double r = sqrt(position[0] * position[0] +
            position[1] * position[1] + position[2] * position[2]);
        *density = g->density[medium] * (1. + a * (g->r[medium + 1] - r));
        return 1. / fabs(a);
}

#endif
