/*
 * pb_dcs.h -- differential cross-sections (DCS) of the discrete energy losses.
 *
 * Restates, for the default models (SSR bremsstrahlung, SSR pair production,
 * DRSS photonuclear; src/pumas.c:192-200; the other models of the reference
 * are in pb_dcs_models.h):
 *   - dcs_evaluate, the single precision log-log PCHIP over table_DCS with its
 *     fall back to the exact model (pumas.c:11243-11347)           [a16]
 *   - the ionisation DCS and its sampler (pumas.c:11082-11226)      [a17]
 *   - the exact models (pumas.c:13915, 14394-14604, 14626-15083),
 *     the dilogarithm (pumas.c:12609) and the kinematic ranges      [a18]
 * Element constants that the reference recomputes with xpow() on every call
 * (Z^(1/3), A^0.27 ...) are taken from `Element`, where the host evaluated the
 * same expressions with the same libm at load time.
 */
#ifndef PB_DCS_H
#define PB_DCS_H

#include <cstring>

#include "pb_tables.h"

namespace pb {

/* Gauss-Legendre nodes and weights on [0, 1], 9 and 12 points
 * (pumas.c:12490-12553). */
#define PB_GQ9_X                                                               \
        {                                                                      \
                0.0159198802461870, 0.0819844463366821, 0.1933142836497048,    \
                    0.3378732882980955, 0.5000000000000000,                    \
                    0.6621267117019045, 0.8066857163502952,                    \
                    0.9180155536633179, 0.9840801197538130                     \
        }
#define PB_GQ9_W                                                               \
        {                                                                      \
                0.0406371941807872, 0.0903240803474287, 0.1303053482014677,    \
                    0.1561735385200015, 0.1651196775006299,                    \
                    0.1561735385200015, 0.1303053482014677,                    \
                    0.0903240803474287, 0.0406371941807872                     \
        }
#define PB_GQ12_X                                                              \
        {                                                                      \
                0.0092196828766404, 0.0479413718147625, 0.1150486629028477,    \
                    0.2063410228566913, 0.3160842505009099,                    \
                    0.4373832957442655, 0.5626167042557344,                    \
                    0.6839157494990901, 0.7936589771433087,                    \
                    0.8849513370971523, 0.9520586281852375,                    \
                    0.9907803171233596                                         \
        }
#define PB_GQ12_W                                                              \
        {                                                                      \
                0.0235876681932559, 0.0534696629976592, 0.0800391642716731,    \
                    0.1015837133615330, 0.1167462682691774,                    \
                    0.1245735229067014, 0.1245735229067014,                    \
                    0.1167462682691774, 0.1015837133615330,                    \
                    0.0800391642716731, 0.0534696629976592,                    \
                    0.0235876681932559                                         \
        }

/* Single precision log / exp. dcs_evaluate (pumas.c:11279, 11320, 11337)
 * calls the C library's logf / expf and, in backward mode, its value
 * multiplies the Monte Carlo weight: a 1 ulp(float) = 6E-08 difference breaks
 * the 1E-09 parity criterion. The reference platform's libm (glibc >= 2.27,
 * which adopted the Arm Optimized Routines) is not correctly rounded (it
 * differs from the rounded binary64 result for 0.02 % of the arguments), so
 * its algorithm is restated here: table driven, evaluated in binary64, rounded
 * once at the end. Published algorithm and constants: Szabolcs Nagy, Arm
 * Optimized Routines, math/logf.c, math/expf.c, math/logf_data.c, math/exp2f_data.c (also musl
 * src/math). tests/test_libm_float.py pins both functions against the system
 * libm. Only finite, normal arguments occur here. */
PB_FN uint32_t bits_of_float(float x)
{
#if defined(__CUDA_ARCH__)
        return __float_as_uint(x);
#else
        uint32_t u;
        memcpy(&u, &x, sizeof u);
        return u;
#endif
}

PB_FN float float_of_bits(uint32_t u)
{
#if defined(__CUDA_ARCH__)
        return __uint_as_float(u);
#else
        float x;
        memcpy(&x, &u, sizeof x);
        return x;
#endif
}

PB_FN double double_of_bits(uint64_t u)
{
#if defined(__CUDA_ARCH__)
        return __longlong_as_double((long long)u);
#else
        double x;
        memcpy(&x, &u, sizeof x);
        return x;
#endif
}

PB_FN uint64_t bits_of_double(double x)
{
#if defined(__CUDA_ARCH__)
        return (uint64_t)__double_as_longlong(x);
#else
        uint64_t u;
        memcpy(&u, &x, sizeof u);
        return u;
#endif
}

/* {1 / c, log(c)} for the 16 sub-intervals of [sqrt(2)/2, sqrt(2)]. */
#define PB_LOGF_TABLE                                                          \
        {                                                                      \
                1.398907162146528, -0.33569133332882284,                   \
                    1.3403141896637998, -0.2929040563774074,               \
                    1.286432210124115, -0.2518726580937369,                \
                    1.2367150214269895, -0.21245868807117255,               \
                    1.1906977166711752, -0.17453945183745634,               \
                    1.1479821020556429, -0.1380057072319758,                  \
                    1.1082251448272158, -0.10275976698545139,               \
                    1.0711297413057381, -0.06871392447020525,               \
                    1.036437278977283, -0.0357891387398228, 1.0,       \
                    0.0, 0.9492859795739057, 0.05204517742929496,        \
                    0.8951049428609004, 0.11081431298787942,                 \
                    0.8476821620351103, 0.1652495223695143,                \
                    0.8050314851692001, 0.21687389031699977,                 \
                    0.7664671008843108, 0.2659635028121397,                \
                    0.731428603316328, 0.3127556664073557                 \
        }

PB_FN float logf_libm(float x)
{
        uint32_t ix = bits_of_float(x);
        if (ix == 0x3f800000u) return 0.f;
        if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
                /* zero, subnormal, negative, inf or nan: never on this path */
                if (ix * 2 == 0) return -HUGE_VALF;
                if (ix == 0x7f800000u) return x;
                if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u)
                        return (x - x) / (x - x);
                ix = bits_of_float(x * 8388608.0f);
                ix -= 23u << 23;
        }
        const double tab[32] = PB_LOGF_TABLE;
        const uint32_t tmp = ix - 0x3f330000u;
        const int i = (int)((tmp >> 19) % 16);
        const int k = (int32_t)tmp >> 23;
        const uint32_t iz = ix - (tmp & 0xff800000u);
        const double invc = tab[2 * i];
        const double logc = tab[2 * i + 1];
        const double z = (double)float_of_bits(iz);
        const double r = z * invc - 1;
        const double y0 = logc + (double)k * 0.6931471805599453;
        const double r2 = r * r;
        double y = 0.333456765744066 * r + -0.4999997485802103;
        y = -0.25089342214237154 * r2 + y;
        y = y * r2 + (y0 + r);
        return (float)y;
}

/* bits(2^(i/32)) - (i << 47), i = 0..31 */
#define PB_EXP2F_TABLE                                                         \
        {                                                                      \
                0x3ff0000000000000, 0x3fefd9b0d3158574, 0x3fefb5586cf9890f,    \
                    0x3fef9301d0125b51, 0x3fef72b83c7d517b,                    \
                    0x3fef54873168b9aa, 0x3fef387a6e756238,                    \
                    0x3fef1e9df51fdee1, 0x3fef06fe0a31b715,                    \
                    0x3feef1a7373aa9cb, 0x3feedea64c123422,                    \
                    0x3feece086061892d, 0x3feebfdad5362a27,                    \
                    0x3feeb42b569d4f82, 0x3feeab07dd485429,                    \
                    0x3feea47eb03a5585, 0x3feea09e667f3bcd,                    \
                    0x3fee9f75e8ec5f74, 0x3feea11473eb0187,                    \
                    0x3feea589994cce13, 0x3feeace5422aa0db,                    \
                    0x3feeb737b0cdc5e5, 0x3feec49182a3f090,                    \
                    0x3feed503b23e255d, 0x3feee89f995ad3ad,                    \
                    0x3feeff76f2fb5e47, 0x3fef199bdd85529c,                    \
                    0x3fef3720dcef9069, 0x3fef5818dcfba487,                    \
                    0x3fef7c97337b9b5f, 0x3fefa4afa2a490da,                    \
                    0x3fefd0765b6e4540                                         \
        }

PB_FN float expf_libm(float x)
{
        const double xd = (double)x;
        const uint32_t abstop = (bits_of_float(x) >> 20) & 0x7ff;
        if (abstop >= 0x42b) { /* |x| >= 88 */
                if (bits_of_float(x) == 0xff800000u) return 0.0f;
                if (abstop >= 0x7f8) return x + x;
                if (x > 88.72283172607422f) return HUGE_VALF;
                if (x < -103.97207641601562f) return 0.0f;
        }
        const uint64_t tab[32] = PB_EXP2F_TABLE;
        const double N = 32;
        double z = (1.4426950408889634 * N) * xd;
        double kd = z + 6755399441055744.0;
        const uint64_t ki = bits_of_double(kd);
        kd -= 6755399441055744.0;
        const double r = z - kd;
        uint64_t t = tab[ki % 32];
        t += ki << 47;
        const double s = double_of_bits(t);
        z = (0.05550361559341535 / N / N / N) * r +
            (0.2402284522445722 / N / N);
        const double r2 = r * r;
        double y = (0.6931471806916203 / N) * r + 1;
        y = z * r2 + y;
        y = y * s;
        return (float)y;
}

PB_FN float logf_cr(float x) { return logf_libm(x); }
PB_FN float expf_cr(float x) { return expf_libm(x); }

/* ---- kinematic ranges --------------------------------------------------- */
/* pumas.c:14230-14239 */
PB_FN void dcs_pair_production_range(
    const Element & el, double m, double K, double * qmin, double * qmax)
{
        const double sqrte = 1.6487212707;
        *qmax = K + m * (1. - 0.75 * sqrte * el.z13);
        *qmin = 4. * ELECTRON_MASS;
}

/* pumas.c:11058-11064 */
PB_FN void dcs_photonuclear_range(double m, double K, double * qmin, double * qmax)
{
        const double M = 0.5 * (NEUTRON_MASS + PROTON_MASS);
        const double mpi = PION_MASS;
        *qmin = mpi + 0.5 * mpi * mpi / M;
        *qmax = K + m - 0.5 * (M + m * m / M);
}

/* pumas.c:10983-10992 */
PB_FN void dcs_get_range(int process, const Element & el, double mass,
    double energy, double * qmin, double * qmax)
{
        if (process == 2) {
                dcs_photonuclear_range(mass, energy, qmin, qmax);
        } else {
                dcs_pair_production_range(el, mass, energy, qmin, qmax);
                if (process == 0) *qmin = 0.;
        }
}

/* ---- ionisation --------------------------------------------------------- */
/* pumas.c:11082-11101 */
PB_FN double electronic_dcs(double Z, double I, double m, double K, double q)
{
        const double P2 = K * (K + 2. * m);
        const double E = K + m;
        const double Wmax = 2. * ELECTRON_MASS * P2 /
            (m * m + ELECTRON_MASS * (ELECTRON_MASS + 2. * E));
        if (q > Wmax) return 0.;
        const double Wmin = (13.6 / 19.2) * I;
        if (q <= Wmin) return 0.;
        const double E2 = E * E;
        const double beta2 = P2 / E2;
        const double a0 = 0.5 / E2;
        const double a1 = beta2 / Wmax;
        const double cs = 2 * PI * ELECTRON_MASS * ELECTRON_RADIUS *
            ELECTRON_RADIUS * Z * (a0 + 1. / q * (1 / q - a1)) / beta2;
        return cs;
}

/* pumas.c:11117-11122 */
PB_FN double dcs_ionisation(
    const Physics & p, const Element & el, double K, double q)
{
        return electronic_dcs(el.Z, el.I, p.mass, K, q) * (p.mass + K) *
            AVOGADRO_NUMBER / (el.A * 1E-03);
}

/* pumas.c:11185-11226 */
template <class Rng>
PB_FN double dcs_ionisation_randomise(
    const Physics & p, Rng & rng, const Element & el, double K, double xlow)
{
        const double P2 = K * (K + 2. * p.mass);
        const double E = K + p.mass;
        const double Wmax = 2. * ELECTRON_MASS * P2 /
            (p.mass * p.mass + ELECTRON_MASS * (ELECTRON_MASS + 2. * E));
        if (Wmax < p.cutoff * K) return K;
        double Wmin = (13.6 / 19.2) * el.I;
        const double qlow = K * xlow;
        if (qlow >= Wmin) Wmin = qlow;

        const double E2 = E * E;
        const double beta2 = P2 / E2;
        const double a0 = 0.5 / E2;
        const double a1 = beta2 / Wmax;

        const double p0 = a0 * (Wmax - Wmin);
        const double p2 = 1. / Wmin - 1. / Wmax;
        double q;
        for (;;) {
                double z = rng.uniform() * (p0 + p2);
                if (z <= p0) {
                        z /= p0;
                        q = (Wmax - Wmin) * z;
                } else {
                        z = (z - p0) / p2;
                        q = Wmin * Wmax / (Wmax - z * (Wmax - Wmin));
                }
                const double r0 = a0 + 1 / (q * q);
                const double r1 = r0 - a1 / q;
                if (rng.uniform() * r0 <= r1) break;
        }
        return K - q;
}

/* ---- bremsstrahlung, Sandrock-Soedingrekso-Rhode ------------------------- */
/* pumas.c:13915-14000. Atomic DCS in m^2 / GeV. */
PB_NOINLINE double dcs_bremsstrahlung_SSR(
    const Physics & p, const Element & el, double m, double K, double q)
{
        const double SQRTE = 1.648721270700128;
        const double ME = 0.5109989461;
        const double RE = 2.8179403227E-13;
        const double MMU = 105.6583745;
        const double Z = el.Z, A = el.A;

        if ((Z <= 0) || (A <= 0) || (m <= 0) || (K <= 0) || (q <= 0)) return 0.;
        const double z13 = el.z13;
        if (q >= K + m * (1. - 0.75 * SQRTE * z13)) return 0.;

        const double energy = (K + m) * 1E+03;
        const double v = q * 1E+03 / energy;
        m *= 1E+03;

        const double Z13 = 1. / z13;
        const double rad_log = el.radiation_log;
        const double rad_log_inel = (Z == 1.) ? 446. : 1429.;
        const double Dn = el.Dn;

        const double mu_qc = m / (MMU * p.c_exp_1 / Dn);
        const double rho = sqrt(1. + 4. * mu_qc * mu_qc);

        const double log_rho = xlog((rho + 1.) / (rho - 1.));
        const double delta1 = xlog(mu_qc) + 0.5 * rho * log_rho;
        const double delta2 = xlog(mu_qc) +
            0.25 * (3. * rho - rho * rho * rho) * log_rho + 2. * mu_qc * mu_qc;

        const double delta = m * m * v / (2. * energy * (1. - v));

        double phi1 = xlog(rad_log * Z13 * (m / ME) /
            (1. + rad_log * Z13 * p.c_exp_1_2 * delta / ME));
        double phi2 = xlog(rad_log * Z13 * p.c_exp_m1_6 * (m / ME) /
            (1. + rad_log * Z13 * p.c_exp_1_3 * delta / ME));
        phi1 -= delta1 * (1. - 1. / Z);
        phi2 -= delta2 * (1. - 1. / Z);

        const double s_atomic_1 =
            xlog(m / delta / (m * delta / (ME * ME) + SQRTE));
        const double s_atomic_2 =
            xlog(1. + ME / (delta * rad_log_inel * Z13 * Z13 * SQRTE));
        const double s_atomic =
            (4. / 3. * (1. - v) + v * v) * (s_atomic_1 - s_atomic_2);

        double s_rad;
        if (v < .0 || v > 1.) {
                s_rad = 0.;
        } else if (v < 0.02) {
                s_rad = -0.00349 + 148.84 * v + -987.531 * v * v;
        } else if (v >= 0.02 && v < 0.1) {
                s_rad = 0.1642 + 132.573 * v + -585.361 * v * v +
                    1407.77 * v * v * v;
        } else if (v >= 0.01 && v < 0.9) {
                s_rad = -2.8922 + -19.0156 * v + 57.698 * v * v;
                const double tmp = xlog(1. - v);
                s_rad += -63.418 * v * xlog(v) + 14.1166 * tmp +
                    1.84206 * tmp * tmp;
        } else {
                s_rad = 2134.19 + 581.823 * v + -2708.85 * v * v;
                const double tmp = xlog(1. - v);
                s_rad += 4767.05 * v * xlog(v) + 1.52918 * tmp +
                    0.361933 * tmp * tmp;
        }

        const double result =
            ((2. - 2. * v + v * v) * phi1 - 2. / 3. * (1. - v) * phi2) +
            1. / Z * s_atomic + 0.25 * ALPHA_EM * phi1 * s_rad;
        if (result <= 0.) return 0.;

        const double aux = 2 * (ME / m) * RE * Z;
        return aux * aux * (ALPHA_EM / q) * result * 1E-04;
}

/* ---- dilogarithm, CERNLIB C332 (pumas.c:12609-12673) --------------------- */
PB_NOINLINE double math_dilog(double x)
{
        const double C[20] = { 0.42996693560813697, 0.40975987533077105,
                -0.01858843665014592, 0.00145751084062268,
                -0.00014304184442340, 0.00001588415541880,
                -0.00000190784959387, 0.00000024195180854,
                -0.00000003193341274, 0.00000000434545063,
                -0.00000000060578480, 0.00000000008612098,
                -0.00000000001244332, 0.00000000000182256,
                -0.00000000000027007, 0.00000000000004042,
                -0.00000000000000610, 0.00000000000000093,
                -0.00000000000000014, 0.00000000000000002 };
        const double PI3 = PI * PI / 3.;
        const double PI6 = PI * PI / 6.;
        const double PI12 = PI * PI / 12.;

        if (x == 10) {
                return PI6;
        } else if ((x == -1.) || (1. - x == 0.)) {
                return -PI12;
        }
        const double T = -x;
        double A, S, Y;
        if (T <= -2) {
                Y = -1. / (1. + T);
                S = 1.;
                const double l0 = xlog(-T), l1 = xlog(1. + 1. / T);
                A = -PI3 + 0.5 * (l0 * l0 - l1 * l1);
        } else if (T < -1.) {
                Y = -1. - T;
                S = -1.;
                A = xlog(-T);
                A = -PI6 + A * (A + xlog(1. + 1. / T));
        } else if (T <= -0.5) {
                Y = -(1. + T) / T;
                S = 1.;
                A = xlog(-T);
                A = -PI6 + A * (-0.5 * A + xlog(1. + T));
        } else if (T < 0.) {
                Y = -T / (1. + T);
                S = -1.;
                const double l = xlog(1 + T);
                A = 0.5 * (l * l);
        } else if (T <= 1.) {
                Y = T;
                S = 1.;
                A = 0.;
        } else {
                Y = 1. / T;
                S = -1.;
                const double l = xlog(T);
                A = PI6 + 0.5 * (l * l);
        }

        const double H = Y + Y - 1;
        const double ALFA = H + H;
        double B0 = 0., B1 = 0., B2 = 0.;
        for (int i = 19; i >= 0; i--) {
                B0 = C[i] + ALFA * B1 - B2;
                B2 = B1;
                B1 = B0;
        }
        return -(S * (B0 - H * B2) + A);
}

/* ---- pair production, Sandrock-Soedingrekso-Rhode ------------------------ */
/* Doubly differential in (q, rho). pumas.c:14394-14544. */
PB_NOINLINE double dcs_pair_production_d2_SSR(const Physics & p,
    const Element & el, double m, double K, double q, double rho)
{
        const double ME = 0.5109989461;
        const double RE = 2.8179403227E-13;
        const double Z = el.Z, A = el.A;

        if ((Z <= 0) || (A <= 0) || (m <= 0) || (K <= 0) || (q <= 0)) return 0.;
        double qmin, qmax;
        dcs_pair_production_range(el, m, K, &qmin, &qmax);
        if ((q < qmin) || (q > qmax)) return 0.;

        const double energy = (K + m) * 1E+03;
        const double v = q / (K + m);
        m *= 1E+03;
        const double rad_log = el.radiation_log;

        const double aR = ALPHA_EM * RE;
        const double const_prefactor = 4. / (3. * PI) * Z * (aR * aR);
        const double Z13 = el.z13_inv;
        const double d_n = el.Dn;

        rho = 1 - rho;
        const double rho2 = rho * rho;

        double g1, g2;
        if (Z == 1.) {
                g1 = 4.4E-05;
                g2 = 4.8E-05;
        } else {
                g1 = 1.95E-05;
                g2 = 5.3E-05;
        }

        const double zeta1 = (0.073 * xlog(energy / m /
                                      (1. + g1 * el.z23 * energy / m)) -
            0.26);
        const double zeta2 =
            (0.058 * xlog(energy / m / (1 + g2 / Z13 * energy / m)) - 0.14);
        double zeta;
        if ((zeta1 > 0.) && (zeta2 > 0.)) {
                zeta = zeta1 / zeta2;
        } else {
                zeta = 0.;
        }

        const double beta = v * v / (2. * (1. - v));
        const double mv = m * v / (2. * ME);
        const double xi = (mv * mv) * (1. - rho2) / (1. - v);

        /* Diagram e */
        const double Be = ((2. + rho2) * (1. + beta) + xi * (3. + rho2)) *
                xlog(1. + 1. / xi) +
            (1. - rho2 - beta) / (1. + xi) - (3. + rho2);
        const double Ce2 = ((1. - rho2) * (1. + beta) + xi * (3. - rho2)) *
                xlog(1. + 1. / xi) +
            2. * (1. - beta - rho2) / (1. + xi) - (3. - rho2);
        const double Ce1 = Be - Ce2;
        const double De = ((2. + rho2) * (1. + beta) + xi * (3. + rho2)) *
                math_dilog(1. / (1. + xi)) -
            (2. + rho2) * xi * xlog(1. + 1. / xi) -
            (xi + rho2 + beta) / (1. + xi);

        const double med = ME / m * d_n;
        const double med2 = med * med;
        double Le1, Le2;
        if (De / Be > 0.) {
                const double tmp = De / Be;
                const double Xe = (tmp < 100) ? xexp(-tmp) : 0.;
                Le1 = xlog(rad_log * Z13 * sqrt(1. + xi) /
                          (Xe + 2. * ME * p.c_exp_1_2 * rad_log * Z13 *
                                  (1. + xi) / (energy * v * (1. - rho2)))) -
                    tmp - 0.5 * xlog(Xe + med2 * (1. + xi));
                Le2 = xlog(rad_log * Z13 * p.c_exp_m1_6 * sqrt(1 + xi) /
                          (Xe + 2. * ME * p.c_exp_1_3 * rad_log * Z13 *
                                  (1. + xi) / (energy * v * (1. - rho2)))) -
                    tmp - 0.5 * xlog(Xe + med2 * p.c_exp_m1_3 * (1. + xi));
        } else {
                const double tmp = De / Be;
                const double Xe_inv = (tmp > -100) ? xexp(De / Be) : 0;
                Le1 = xlog(rad_log * Z13 * sqrt(1. + xi) /
                          (1. + Xe_inv * 2. * ME * p.c_exp_1_2 * rad_log * Z13 *
                                  (1. + xi) / (energy * v * (1. - rho2)))) -
                    0.5 * De / Be - 0.5 * xlog(1. + Xe_inv * med2 * (1. + xi));
                Le2 = xlog(rad_log * Z13 * p.c_exp_m1_6 * sqrt(1 + xi) /
                          (1. + Xe_inv * 2. * ME * p.c_exp_1_3 * rad_log * Z13 *
                                  (1. + xi) / (energy * v * (1. - rho2)))) -
                    0.5 * De / Be -
                    0.5 * xlog(1. + Xe_inv * med2 * p.c_exp_m1_3 * (1. + xi));
        }

        double diagram_e = const_prefactor * (Z + zeta) * (1. - v) / v *
            (Ce1 * Le1 + Ce2 * Le2);
        if (diagram_e < 0.) diagram_e = 0.;

        /* Diagram mu */
        const double Bm = ((1. + rho2) * (1. + (3. * beta) / 2) -
                              1. / xi * (1. + 2. * beta) * (1. - rho2)) *
                xlog(1. + xi) +
            xi * (1. - rho2 - beta) / (1. + xi) + (1. + 2. * beta) * (1. - rho2);
        const double Cm2 = ((1. - beta) * (1. - rho2) - xi * (1. + rho2)) *
                xlog(1. + xi) / xi -
            2. * (1. - beta - rho2) / (1. + xi) + 1. - beta -
            (1. + beta) * rho2;
        const double Cm1 = Bm - Cm2;
        const double Dm = ((1. + rho2) * (1. + (3. * beta) / 2.) -
                              1. / xi * (1. + 2. * beta) * (1. - rho2)) *
                math_dilog(xi / (1. + xi)) +
            (1. + (3. * beta) / 2.) * (1. - rho2) / xi * xlog(1. + xi) +
            (1. - rho2 - beta / 2. * (1. + rho2) +
                (1. - rho2) / (2. * xi) * beta) *
                xi / (1. + xi);

        double Lm1, Lm2;
        if (Dm / Bm > 0.) {
                const double tmp = Dm / Bm;
                const double Xm = (tmp < 100) ? xexp(-tmp) : 0.;
                Lm1 = xlog(m / ME * rad_log * Z13 / d_n /
                          (Xm + 2. * ME * p.c_exp_1_2 * rad_log * Z13 *
                                  (1. + xi) / (energy * v * (1. - rho2)))) -
                    tmp;
                Lm2 = xlog(m / ME * rad_log * Z13 / d_n /
                          (Xm + 2. * ME * p.c_exp_1_3 * rad_log * Z13 *
                                  (1. + xi) / (energy * v * (1. - rho2)))) -
                    tmp;
        } else {
                const double tmp = Dm / Bm;
                const double Xmv = (tmp > -100) ? xexp(tmp) : 0.;
                Lm1 = xlog(m / ME * rad_log * Z13 / d_n /
                    (1. + 2. * ME * p.c_exp_1_2 * rad_log * Z13 * (1. + xi) /
                            (energy * v * (1. - rho2)) * Xmv));
                Lm2 = xlog(m / ME * rad_log * Z13 / d_n /
                    (1. + 2. * ME * p.c_exp_1_3 * rad_log * Z13 * (1. + xi) /
                            (energy * v * (1. - rho2)) * Xmv));
        }

        const double mem = ME / m;
        double diagram_mu = const_prefactor * (Z + zeta) * (1. - v) / v *
            (mem * mem) * (Cm1 * Lm1 + Cm2 * Lm2);
        if (diagram_mu < 0.) diagram_mu = 0.;

        return (diagram_e + diagram_mu) * 1E-01 / energy;
}

/* Integrated over the pair asymmetry with a 12 point quadrature in ln(rho).
 * pumas.c:14560-14604. */
PB_NOINLINE double dcs_pair_production_SSR(
    const Physics & p, const Element & el, double mass, double K, double q)
{
        const double Z = el.Z, A_ = el.A;
        if ((Z <= 0) || (A_ <= 0) || (mass <= 0) || (K <= 0) || (q <= 0))
                return 0.;
        if (q <= 4. * ELECTRON_MASS) return 0.;
        const double sqrte = 1.6487212707;
        if (q >= K + mass * (1. - 0.75 * sqrte * el.z13)) return 0.;

        const double gamma = 1. + K / mass;
        const double x0 = 1. - 4. * ELECTRON_MASS / q;
        const double x1 = 1. - 6. / (gamma * (gamma - q / mass));
        double rmax;
        if ((x0 > 0) && (x1 > 0)) {
                rmax = sqrt(x0) * x1;
        } else {
                return 0.;
        }
        const double ri = 1. - rmax;
        if ((ri <= 0.) || (ri >= 1.)) return 0.;
        const double tmin = xlog(ri);

        const double xGQ[12] = PB_GQ12_X;
        const double wGQ[12] = PB_GQ12_W;
        double I = 0.;
        for (int i = 0; i < 12; i++) {
                const double rho = xexp(xGQ[i] * tmin);
                I -= dcs_pair_production_d2_SSR(p, el, mass, K, q, rho) * rho *
                    wGQ[i] * tmin;
        }
        return (I < 0) ? 0. : I;
}

/* ---- photonuclear, Dutta-Reno-Sarcevic-Seckel with ALLM97 ---------------- */
/* pumas.c:14626-14673 */
PB_NOINLINE double dcs_photonuclear_f2p_ALLM97(double x, double Q2)
{
        const double m02 = 0.31985;
        const double mP2 = 49.457;
        const double mR2 = 0.15052;
        const double Q02 = 0.52544;
        const double Lambda2 = 0.06527;

        const double cP1 = 0.28067, cP2 = 0.22291, cP3 = 2.1979;
        const double aP1 = -0.0808, aP2 = -0.44812, aP3 = 1.1709;
        const double bP1 = 0.36292, bP2 = 1.8917, bP3 = 1.8439;
        const double cR1 = 0.80107, cR2 = 0.97307, cR3 = 3.4942;
        const double aR1 = 0.58400, aR2 = 0.37888, aR3 = 2.6063;
        const double bR1 = 0.01147, bR2 = 3.7582, bR3 = 0.49338;

        const double M = 0.5 * (PROTON_MASS + NEUTRON_MASS);
        const double M2 = M * M;
        const double W2 = M2 + Q2 * (1.0 / x - 1.0);
        const double t = xlog(xlog((Q2 + Q02) / Lambda2) / xlog(Q02 / Lambda2));
        const double xP = (Q2 + mP2) / (Q2 + mP2 + W2 - M2);
        const double xR = (Q2 + mR2) / (Q2 + mR2 + W2 - M2);
        const double cP = cP1 + (cP1 - cP2) * (1.0 / (1.0 + xpow(t, cP3)) - 1.0);
        const double aP = aP1 + (aP1 - aP2) * (1.0 / (1.0 + xpow(t, aP3)) - 1.0);
        const double bP = bP1 + bP2 * xpow(t, bP3);
        const double cR = cR1 + cR2 * xpow(t, cR3);
        const double aR = aR1 + aR2 * xpow(t, aR3);
        const double bR = bR1 + bR2 * xpow(t, bR3);

        const double F2P = cP * xexp(aP * xlog(xP) + bP * xlog(1 - x));
        const double F2R = cR * xexp(aR * xlog(xR) + bR * xlog(1 - x));
        return Q2 / (Q2 + m02) * (F2P + F2R);
}

/* pumas.c:14688-14699 */
PB_FN double dcs_photonuclear_shadowing_DRSS(const Element & el, double x)
{
        if (el.Z == 1.) return 1.;
        if (x < 0.0014)
                return xexp(-0.1 * el.lnA);
        else if (x < 0.04)
                return xexp((0.069 * xlog10(x) + 0.097) * el.lnA);
        else
                return 1.;
}

/* d2S / (dq dQ2), pumas.c:14716-14721 and 14923-14943. */
PB_FN double dcs_photonuclear_d2_DRSS(
    const Element & el, double ml, double K, double q, double Q2)
{
        const double Z = el.Z, A = el.A;
        const double cf = 4 * PI * ALPHA_EM * ALPHA_EM * HBAR_C * HBAR_C;
        const double M = 0.5 * (PROTON_MASS + NEUTRON_MASS);
        const double E = K + ml;
        const double y = q / E;
        const double x = 0.5 * Q2 / (M * q);
        const double F2p = dcs_photonuclear_f2p_ALLM97(x, Q2);
        const double shadowing = dcs_photonuclear_shadowing_DRSS(el, x);
        const double F2A = F2p * shadowing *
            (Z + (A - Z) * (1.0 + x * (-1.85 + x * (2.45 + x * (-2.35 + x)))));
        const double R = 0.;
        const double dds =
            (1 - y +
                0.5 * (1 - 2 * ml * ml / Q2) * (y * y + Q2 / (E * E)) /
                    (1 + R)) /
                (Q2 * Q2) -
            0.25 / (E * E * Q2);
        return cf * F2A * dds / q;
}

/* Integral over Q2, 9 points in ln(Q2). pumas.c:15002-15060 with mode 0 and
 * math_gauss_quad_initialise (pumas.c:12575-12602) inlined. */
PB_NOINLINE double dcs_photonuclear_DRSS(
    const Element & el, double ml, double K, double q)
{
        const double Z = el.Z, A = el.A;
        if ((Z <= 0) || (A <= 0) || (ml <= 0) || (K <= 0)) return 0.;
        const double M = 0.5 * (NEUTRON_MASS + PROTON_MASS);
        const double mpi = PION_MASS;
        if ((q <= mpi + 0.5 * mpi * mpi / M) ||
            (q >= K + ml - 0.5 * (M + ml * ml / M)))
                return 0.;

        const double E = K + ml;
        const double ml2 = ml * ml;
        const double Q2min = ml2 * (q * q - 0.5 * ml2) / (E * (E - q));
        const double Q2max = 2.0 * M * (q - mpi) - mpi * mpi;
        if ((Q2max < Q2min) | (Q2min < 0)) return 0.;

        const double x9[9] = PB_GQ9_X;
        const double w9[9] = PB_GQ9_W;
        const double xmin = xlog(Q2min);
        const double xmax = xlog(Q2max);
        const double dx = xmax - xmin;
        double ds = 0.;
        for (int i = 0; i < 9; i++) {
                const double Q2 = xexp(xmin + dx * x9[i]);
                const double w = w9[i] * (dx * Q2);
                const double tmp = dcs_photonuclear_d2_DRSS(el, ml, K, q, Q2);
                if (tmp <= 0.) continue;
                ds += tmp * w;
        }
        return (ds < 0.) ? 0. : ds;
}

} /* namespace pb */
#include "pb_dcs_models.h"
namespace pb {

/* The doubly differential photonuclear cross-section of the physics
 * (dcs_photonuclear_ddcs, pumas.c:15160-15170): BM has its own, DRSS and
 * BBKS share the DRSS one. */
PB_FN double dcs_photonuclear_d2(const Physics & p, const Element & el,
    double ml, double K, double q, double Q2)
{
        if (p.model[2] == PHOTONUCLEAR_BM)
                return dcs_photonuclear_d2_BM(el, ml, K, q, Q2);
        return dcs_photonuclear_d2_DRSS(el, ml, K, q, Q2);
}

/* ---- process dispatch (dcs_get, pumas.c:10954; wrappers 11003-11048) ----- */
/* Is the user's model of the process callable here? On the host always (the
 * host view holds host functions); on the device in the kernels device-linked
 * with user code only. */
#if !defined(__CUDA_ARCH__) || defined(PB_USER_DCS)
#define PB_USER_DCS_CALL(p, i, el, K, q)                                       \
        (((p).model[i] == MODEL_USER) && ((p).user_dcs[i] != nullptr))
#else
#define PB_USER_DCS_CALL(p, i, el, K, q) false
#endif

PB_FN double dcs_model(
    const Physics & p, int process, const Element & el, double K, double q)
{
        if (process == 3) return dcs_ionisation(p, el, K, q);
        double d;
        if (process == 0) {
                if (PB_USER_DCS_CALL(p, 0, el, K, q))
                        d = p.user_dcs[0](el.Z, el.A, p.mass, K, q);
                else if (p.model[0] == BREMSSTRAHLUNG_KKP)
                        d = dcs_bremsstrahlung_KKP(el, p.mass, K, q);
                else if (p.model[0] == BREMSSTRAHLUNG_ABB)
                        d = dcs_bremsstrahlung_ABB(el, p.mass, K, q);
                else
                        d = dcs_bremsstrahlung_SSR(p, el, p.mass, K, q);
        } else if (process == 1) {
                if (PB_USER_DCS_CALL(p, 1, el, K, q))
                        d = p.user_dcs[1](el.Z, el.A, p.mass, K, q);
                else if (p.model[1] == PAIR_PRODUCTION_KKP)
                        d = dcs_pair_production_KKP(el, p.mass, K, q);
                else
                        d = dcs_pair_production_SSR(p, el, p.mass, K, q);
        } else {
                double qmin, qmax;
                dcs_photonuclear_range(p.mass, K, &qmin, &qmax);
                if ((q < qmin) || (q > qmax)) return 0.;
                if (PB_USER_DCS_CALL(p, 2, el, K, q))
                        d = p.user_dcs[2](el.Z, el.A, p.mass, K, q);
                else if (p.model[2] == PHOTONUCLEAR_BM)
                        d = dcs_photonuclear_BM(el, p.mass, K, q);
                else if (p.model[2] == PHOTONUCLEAR_BBKS)
                        d = dcs_photonuclear_BBKS(el, p.mass, K, q);
                else
                        d = dcs_photonuclear_DRSS(el, p.mass, K, q);
        }
        return d * 1E+03 * AVOGADRO_NUMBER * (p.mass + K) / el.A;
}

/* Index of x in the (float) table_DCS_x, pumas.c:4236-4244. */
PB_FN int table_bracketf(const float * table, float value, int n)
{
        int i1 = 0, i2 = n - 1;
        do {
                const int i3 = (i1 + i2) / 2;
                if (value >= table[i3])
                        i1 = i3;
                else
                        i2 = i3;
        } while (i2 - i1 >= 2);
        return i1;
}

/* pumas.c:11243-11347. `iel` is the element index. */
PB_FN double dcs_evaluate(
    const Physics & p, int process, int iel, double K, double q)
{
        const Element & el = p.element[iel];
        const double wj = K / (K + p.mass);
        if (process == 3) return dcs_model(p, process, el, K, q) * wj;

        double qlow = 0;
        if (process == 2) {
                double qmin, qmax;
                dcs_photonuclear_range(p.mass, K, &qmin, &qmax);
                if ((q < qmin) || (q > qmax))
                        return 0.;
                else
                        qlow = 2 * qmin;
        } else {
                double qmin, qmax;
                dcs_pair_production_range(el, p.mass, K, &qmin, &qmax);
                if (process == 1)
                        qlow = 4 * qmin;
                else
                        qmin = 0.;
                if ((q < qmin) || (q > qmax)) return 0.;
        }
        if (q < qlow) return dcs_model(p, process, el, K, q) * wj;

        const int nx = p.n_table_dcs;
        const float xmin = p.DCS_x[0];
        const float xmax = p.DCS_x[nx - 1];
        const float x = logf_cr((float)(q / K));
        if ((x < xmin) || (x >= xmax))
                return dcs_model(p, process, el, K, q) * wj;
        const int ix = table_bracketf(p.DCS_x, x, nx);
        if ((ix < 0) || (ix >= nx - 1))
                return dcs_model(p, process, el, K, q) * wj;

        float p00, p01, m00, m01, p10, p11, m10, m11, h0, h1;
        const int imax = p.n_energies - 1;
        const double Kmax = p.K[imax];
        int i0 = 0;
        if ((K >= Kmax) || ((i0 = table_index(p, p.K, K)) >= imax)) {
                const float * data = p.DCS +
                    (size_t)nx * 2 *
                        ((process * p.n_elements + iel) * p.n_energies + imax);
                p00 = p10 = data[2 * ix];
                m00 = m10 = data[2 * ix + 1];
                p01 = p11 = data[2 * ix + 2];
                m01 = m11 = data[2 * ix + 3];
                h0 = 1.f;
                h1 = 0.f;
        } else {
                const float * data = p.DCS +
                    (size_t)nx * 2 *
                        ((process * p.n_elements + iel) * p.n_energies + i0);
                p00 = data[2 * ix];
                m00 = data[2 * ix + 1];
                p01 = data[2 * ix + 2];
                m01 = data[2 * ix + 3];
                data += 2 * nx;
                p10 = data[2 * ix];
                m10 = data[2 * ix + 1];
                p11 = data[2 * ix + 2];
                m11 = data[2 * ix + 3];

                const float K0 = (float)(p.K[i0]);
                const float K1 = (float)(p.K[i0 + 1]);
                h1 = logf_cr(((float)K) / K0) / logf_cr(K1 / K0);
                h0 = 1.f - h1;
        }

        if (((p00 == 0.f) && (m00 == 0.f)) || ((p01 == 0.f) && (m01 == 0.f)) ||
            ((p10 == 0.f) && (m10 == 0.f)) || ((p11 == 0.f) && (m11 == 0.f))) {
                return dcs_model(p, process, el, K, q) * wj;
        }

        const float x0 = p.DCS_x[ix];
        const float x1 = p.DCS_x[ix + 1];
        const float dx = x1 - x0;
        m00 *= dx;
        m01 *= dx;
        const float t = (x - x0) / dx;
        const float c02 = -3 * (p00 - p01) - 2 * m00 - m01;
        const float c03 = 2 * (p00 - p01) + m00 + m01;
        float r = expf_cr(p00 + t * (m00 + t * (c02 + t * c03))) * h0;
        if (h1 > 0.f) {
                m10 *= dx;
                m11 *= dx;
                const float c12 = -3 * (p10 - p11) - 2 * m10 - m11;
                const float c13 = 2 * (p10 - p11) + m10 + m11;
                r += expf_cr(p10 + t * (m10 + t * (c12 + t * c13))) * h1;
        }
        return r * wj;
}

} /* namespace pb */

#endif
