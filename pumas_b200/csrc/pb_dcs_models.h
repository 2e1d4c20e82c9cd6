/*
 * pb_dcs_models.h -- the non-default differential cross-section models
 * (SURVEY.md section 8f.4). A physics dump records the models its tables were
 * built with (pumas.c:2124-2149); at transport time the model itself is only
 * evaluated where the tabulation gives up (dcs_evaluate fall backs,
 * pumas.c:11279-11346) and, for the photonuclear process, in the polar angle
 * sampler (pumas.c:11858, 11920).
 *
 * Restates, atomic DCS in m^2 / GeV:
 *   bremsstrahlung   KKP   pumas.c:13759-13808   Kelner, Kokoulin & Petrukhin
 *                    ABB   pumas.c:13824-13899   Andreev, Bezrukov & Bugaev
 *   pair production  KKP   pumas.c:14255-14376
 *   photonuclear     BM    pumas.c:14732-14981   Butkevich & Mikheyev
 *                    BBKS  pumas.c:15193-15372   Bezrukov, Bugaev + hard part
 * Powers of Z and A that the reference evaluates with the run time libm on
 * every call are taken from `Element` (folded by the host libm at load time,
 * pb_dump.hpp), like for the default models in pb_dcs.h.
 */
#ifndef PB_DCS_MODELS_H
#define PB_DCS_MODELS_H

namespace pb {

/* Model indices, per process (Physics::model[]). */
enum {
        BREMSSTRAHLUNG_SSR = 0,
        BREMSSTRAHLUNG_KKP = 1,
        BREMSSTRAHLUNG_ABB = 2,
        PAIR_PRODUCTION_SSR = 0,
        PAIR_PRODUCTION_KKP = 1,
        PHOTONUCLEAR_DRSS = 0,
        PHOTONUCLEAR_BM = 1,
        PHOTONUCLEAR_BBKS = 2
};

/* pumas.c:13759-13808 */
PB_NOINLINE double dcs_bremsstrahlung_KKP(
    const Element & el, double mu, double K, double q)
{
        const double Z = el.Z, A = el.A;
        if ((Z <= 0) || (A <= 0) || (mu <= 0) || (K <= 0) || (q <= 0)) return 0.;

        const double Z13 = el.z13;
        const double sqrte = 1.648721271;
        if (q >= K + mu * (1. - 0.75 * sqrte * Z13)) return 0.;

        const double me = ELECTRON_MASS;
        const double phie_factor = mu / (me * me * sqrte);
        const double rem = 5.63588E-13 * me / mu;

        const double BZ_n = el.radiation_log / Z13;
        const double BZ_e = (Z == 1.) ? 446. : 1429. / (Z13 * Z13);
        const double D_n = el.Dn;
        const double E = K + mu;
        const double dcs_factor = 7.297182E-07 * rem * rem * Z / E;

        const double delta_factor = 0.5 * mu * mu / E;
        const double qe_max = E / (1. + 0.5 * mu * mu / (me * E));

        const double nu = q / E;
        const double delta = delta_factor * nu / (1. - nu);
        const double muD_factor = mu + delta * (D_n * sqrte - 2.);
        double Phi_n =
            xlog(BZ_n * muD_factor / (D_n * (me + delta * sqrte * BZ_n)));
        if (Phi_n < 0.) Phi_n = 0.;
        double Phi_x = 0.;
        if (Z >= 2) {
                Phi_x = xlog(mu * D_n / muD_factor);
                if (Phi_x < 0.) Phi_x = 0.;
        }
        double Phi_e = 0.;
        if (q < qe_max) {
                Phi_e = xlog(BZ_e * mu /
                    ((1. + delta * phie_factor) * (me + delta * sqrte * BZ_e)));
                if (Phi_e < 0.) Phi_e = 0.;
        }

        const double dcs = dcs_factor * (Z * Phi_n + Phi_x + Phi_e) *
            (4. / 3. * (1. / nu - 1.) + nu);
        return (dcs < 0.) ? 0. : dcs;
}

/* pumas.c:13824-13899 */
PB_NOINLINE double dcs_bremsstrahlung_ABB(
    const Element & el, double m, double K, double q)
{
        const double SQRTE = 1.648721270700128;
        const double ME = 0.5109989461;
        const double RE = 2.8179403227E-13;
        const double MMU = 105.6583745;
        const double Z = el.Z, A = el.A;
        if ((Z <= 0) || (A <= 0) || (m <= 0) || (K <= 0) || (q <= 0)) return 0.;

        const double Z13 = el.z13;
        if (q >= K + m * (1. - 0.75 * SQRTE * Z13)) return 0.;

        /* GeV to MeV */
        const double energy = (K + m) * 1E+03;
        const double v = q * 1E+03 / energy;
        m *= 1E+03;

        const double Z3 = 1. / Z13;
        const double a1 = 184.15 * Z3 / (SQRTE * ME);
        const double a2 = 1194 * Z3 * Z3 / (SQRTE * ME);

        const double qc = 1.9 * MMU * Z3;
        double aux = 2 * m / qc;
        const double zeta = sqrt(1 + aux * aux);

        const double delta = m * m * v / (2 * energy * (1 - v));
        const double x1 = a1 * delta;
        const double x2 = a2 * delta;

        double aux1, aux2, d1, d2, psi1, psi2;
        if (Z == 1) {
                d1 = 0;
                d2 = 0;
        } else {
                aux1 = xlog(m / qc);
                aux2 = 0.5 * zeta * xlog((zeta + 1) / (zeta - 1));
                d1 = aux1 + aux2;
                d2 = aux1 + 0.5 * ((3 - zeta * zeta) * aux2 + aux * aux);
        }

        aux = m * a1;
        aux1 = xlog(aux * aux / (1 + x1 * x1));
        aux = m * a2;
        aux2 = xlog(aux * aux / (1 + x2 * x2));
        psi1 = 0.5 * ((1 + aux1) + (1 + aux2) / Z);
        psi2 = 0.5 * ((2. / 3 + aux1) + (2. / 3 + aux2) / Z);

        aux1 = x1 * xatan(1 / x1);
        aux2 = x2 * xatan(1 / x2);
        psi1 -= aux1 + aux2 / Z;
        aux = x1 * x1;
        psi2 += 2 * aux * (1 - aux1 + 0.75 * xlog(aux / (1 + aux)));
        aux = x2 * x2;
        psi2 += 2 * aux * (1 - aux2 + 0.75 * xlog(aux / (1 + aux))) / Z;

        psi1 -= d1;
        psi2 -= d2;
        const double result =
            (2 - 2 * v + v * v) * psi1 - (2. / 3) * (1 - v) * psi2;
        if (result < 0) return 0;

        aux = 2 * (ME / m) * RE * Z;
        return aux * aux * (ALPHA_EM / q) * result * 1E-04;
}

/* pumas.c:14255-14376: integral over t = ln(1 - rho), 12 points. */
PB_NOINLINE double dcs_pair_production_KKP(
    const Element & el, double mass, double K, double q)
{
        const double Z = el.Z, A_ = el.A;
        if ((Z <= 0) || (A_ <= 0) || (mass <= 0) || (K <= 0) || (q <= 0))
                return 0.;

        double qmin, qmax;
        dcs_pair_production_range(el, mass, K, &qmin, &qmax);
        if ((q < qmin) || (q > qmax)) return 0.;

        const double nu = q / (K + mass);
        const double r = mass / ELECTRON_MASS;
        const double beta = 0.5 * nu * nu / (1. - nu);
        const double xi_factor = 0.5 * r * r * beta;
        const double A = el.radiation_log;
        const double Z13 = el.z13;
        const double AZ13 = A / Z13;
        const double sqrte = 1.6487212707;
        const double cL = 2. * sqrte * ELECTRON_MASS * AZ13;
        const double cLe = 2.25 * Z13 * Z13 / (r * r);

        const double gamma = 1. + K / mass;
        const double x0 = 4. * ELECTRON_MASS / q;
        const double x1 = 6. / (gamma * (gamma - q / mass));
        const double argmin =
            (x0 + 2. * (1. - x0) * x1) / (1. + (1. - x1) * sqrt(1. - x0));
        if ((argmin >= 1.) || (argmin <= 0.)) return 0.;
        const double tmin = xlog(argmin);

        const double xGQ[12] = PB_GQ12_X;
        const double wGQ[12] = PB_GQ12_W;
        double I = 0.;
        for (int i = 0; i < 12; i++) {
                const double eps = xexp(xGQ[i] * tmin);
                const double rho = 1. - eps;
                const double rho2 = rho * rho;
                const double rho21 = eps * (2. - eps);
                const double xi = xi_factor * rho21;
                const double xi_i = 1. / xi;

                /* the electron diagrams */
                double Be;
                if (xi >= 1E+03)
                        Be =
                            0.5 * xi_i * ((3 - rho2) + 2. * beta * (1. + rho2));
                else
                        Be = ((2. + rho2) * (1. + beta) + xi * (3. + rho2)) *
                                xlog(1. + xi_i) +
                            (rho21 - beta) / (1. + xi) - 3. - rho2;
                const double Ye = (5. - rho2 + 4. * beta * (1. + rho2)) /
                    (2. * (1. + 3. * beta) * xlog(3. + xi_i) - rho2 -
                        2. * beta * (2. - rho2));
                const double xe = (1. + xi) * (1. + Ye);
                const double cLi = cL / rho21;
                const double Le = xlog(AZ13 * sqrt(xe) * q / (q + cLi * xe)) -
                    0.5 * xlog(1. + cLe * xe);
                double Phi_e = Be * Le;
                if (Phi_e < 0.) Phi_e = 0.;

                /* the muon diagrams */
                double Bmu;
                if (xi <= 1E-03)
                        Bmu = 0.5 * xi * (5. - rho2 + beta * (3. + rho2));
                else
                        Bmu = ((1. + rho2) * (1. + 1.5 * beta) -
                                  xi_i * (1. + 2. * beta) * rho21) *
                                xlog(1. + xi) +
                            xi * (rho21 - beta) / (1. + xi) +
                            (1. + 2. * beta) * rho21;
                const double Ymu = (4. + rho2 + 3. * beta * (1. + rho2)) /
                    ((1. + rho2) * (1.5 + 2. * beta) * xlog(3. + xi) + 1. -
                        1.5 * rho2);
                const double xmu = (1. + xi) * (1. + Ymu);
                const double Lmu =
                    xlog(r * AZ13 * q / (1.5 * Z13 * (q + cLi * xmu)));
                double Phi_mu = Bmu * Lmu;
                if (Phi_mu < 0.) Phi_mu = 0.;

                I -= (Phi_e + Phi_mu / (r * r)) * (1. - rho) * wGQ[i] * tmin;
        }

        /* atomic electrons */
        double zeta;
        if (gamma <= 35.) {
                zeta = 0.;
        } else {
                const double gamma1 = (Z == 1.) ? 4.4E-05 : 1.95E-05;
                const double gamma2 = (Z == 1.) ? 4.8E-05 : 5.30E-05;
                zeta = 0.073 * xlog(gamma / (1. + gamma1 * gamma * Z13 * Z13)) -
                    0.26;
                if (zeta <= 0.)
                        zeta = 0.;
                else
                        zeta /= 0.058 *
                                xlog(gamma / (1. + gamma2 * gamma * Z13)) -
                            0.14;
        }

        const double E = K + mass;
        const double dcs =
            1.794664E-34 * Z * (Z + zeta) * (E - q) * I / (q * E);
        return (dcs < 0.) ? 0. : dcs;
}

/* Shadowing of Butkevich & Mikheyev, pumas.c:14768-14822. The Wood-Saxon
 * factor (pumas.c:14732-14748) is a constant of the element. */
PB_FN double dcs_photonuclear_shadowing_BM(
    const Element & el, double x, double q)
{
        const double Z = el.Z, A = el.A;
        if (Z == 1.) return 1;

        if (x > 0.3) {
                const double Mb = 0.437;
                const double la = 0.5;
                const double x2 = 0.278;
                const double au = 1. / (1. - x);
                const double ac = 1. / (1. - x2);
                const double MPI = 0.13957018;
                const double M = ((A - Z) * NEUTRON_MASS + Z * PROTON_MASS) / A;
                const double Aosc = (1. - la * x) *
                    (au - ac - MPI / M * (au * au - ac * ac));
                return 1. - Mb * el.wood_saxon * Aosc;
        } else {
                const double M1 = 0.129;
                const double M2 = 0.456;
                const double M3 = 0.553;
                const double ws = el.wood_saxon;
                const double m1 = M1 * ws;
                const double m2 = M2 * ws;
                const double m3 = M3 * ws;
                const double sgn = 112.2 *
                    (0.609 * xpow(q, 0.0988) + 1.037 * xpow(q, -0.5944));
                const double tmp = 0.00282 * el.a13 * sgn;
                double G = (3. / tmp) *
                    (0.5 + ((1. + tmp) * xexp(-tmp) - 1.) / (tmp * tmp));
                G = 0.75 * G + 0.25;
                const double x0 = xpow(G / (1. + m2), 1. / m1);
                if (x >= x0)
                        return xpow(x, m1) * (1 + m2) * (1 - m3 * x);
                else
                        return G;
        }
}

/* Nuclear structure function of Butkevich & Mikheyev, pumas.c:14837-14921. */
PB_FN double dcs_photonuclear_f2a_BM(
    const Element & el, double shadowing, double x, double Q2)
{
        const double Z = el.Z, A = el.A;
        Q2 *= 1E+06; /* GeV^2 to MeV^2 */

        const double a = 0.2513E+06;
        const double b = 0.6186E+06;
        const double c = 3.0292E+06;
        const double d = 1.4817E+06;
        const double intercept_pomeron = 0.0988;
        const double intercept_reggeon = 0.4056;
        const double tau = 1.8152;
        const double A_s = 0.12;
        const double B_up = 1.2437;
        const double B_down = 0.1853;

        const double n = 1.5 * (1. + Q2 / (Q2 + c));
        const double dl = intercept_pomeron * (1. + 2. * Q2 / (Q2 + d));
        /* sea quarks and gluons */
        double aux = A_s * xpow(x, -dl) * xpow(Q2 / (Q2 + a), 1. + dl);
        const double F_proton_singlet = aux * xpow(1. - x, n + 4.);
        const double F_neutron_singlet = aux * xpow(1. - x, n + tau);
        /* valence quarks */
        aux = xpow(x, 1. - intercept_reggeon) * xpow(1. - x, n) *
            xpow(Q2 / (Q2 + b), intercept_reggeon);
        const double Up_valence = B_up * aux;
        const double Down_valence = B_down * aux * (1. - x);
        const double F_proton_non_singlet = Up_valence + Down_valence;
        const double F_neutron_non_singlet =
            Up_valence / 4. + Down_valence * 4.;
        const double structure_function_proton =
            F_proton_singlet + F_proton_non_singlet;
        const double structure_function_neutron =
            F_neutron_singlet + F_neutron_non_singlet;

        return shadowing *
            (Z * structure_function_proton +
                (A - Z) * structure_function_neutron);
}

/* d2S / (dq dQ2), pumas.c:14963-14981. */
PB_NOINLINE double dcs_photonuclear_d2_BM(
    const Element & el, double ml, double K, double q, double Q2)
{
        const double cf = 4 * PI * ALPHA_EM * ALPHA_EM * HBAR_C * HBAR_C;
        const double M = 0.5 * (PROTON_MASS + NEUTRON_MASS);
        const double E = K + ml;
        const double y = q / E;
        const double x = 0.5 * Q2 / (M * q);
        const double shadowing = dcs_photonuclear_shadowing_BM(el, x, q);
        const double F2A = dcs_photonuclear_f2a_BM(el, shadowing, x, Q2);
        const double R = 0.25;
        const double dds =
            (1 - y +
                0.5 * (1 - 2 * ml * ml / Q2) * (y * y + Q2 / (E * E)) /
                    (1 + R)) /
                (Q2 * Q2) -
            0.25 / (E * E * Q2);
        return cf * F2A * dds / q;
}

/* Integral over Q2 (pumas.c:15002-15060, mode 0) of the BM doubly
 * differential cross-section, pumas.c:15099-15104. */
PB_NOINLINE double dcs_photonuclear_BM(
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
                const double tmp = dcs_photonuclear_d2_BM(el, ml, K, q, Q2);
                if (tmp <= 0.) continue;
                ds += tmp * w;
        }
        return (ds < 0.) ? 0. : ds;
}

/* Photon-nucleon cross-section of Kokoulin, pumas.c:15193-15203. */
PB_FN double dcs_photonuclear_phn_Kokoulin(double q)
{
        if (q <= 17.) {
                return 96.1 + 82. / sqrt(q);
        } else if (q <= 200.) {
                const double aux = xlog(0.0213 * q);
                return 114.3 + 1.647 * aux * aux;
        } else {
                return 49.2 + 11.1 * xlog(q) + 151.8 / sqrt(q);
        }
}

/* Soft part of Bezrukov & Bugaev as enhanced by Bugaev & Shlepin, plus the
 * hard part of Bugaev, Montaruli, Shlepin & Sokalski. pumas.c:15223-15372. */
PB_NOINLINE double dcs_photonuclear_BBKS(
    const Element & el, double m, double K, double q)
{
        const double Z = el.Z, A = el.A;
        if ((Z <= 0) || (A <= 0) || (m <= 0) || (K <= 0)) return 0.;

        const double M = 0.5 * (NEUTRON_MASS + PROTON_MASS);
        const double mpi = PION_MASS;
        if ((q <= mpi + 0.5 * mpi * mpi / M) ||
            (q >= K + m - 0.5 * (M + m * m / M)))
                return 0.;

        const double m1 = 0.54;
        const double m2 = 1.80;
        const double sgn = dcs_photonuclear_phn_Kokoulin(q);
        const double v = q / (K + m);

        double G;
        if (Z == 1) {
                G = 1;
        } else {
                const double tmp = 0.00282 * el.a13 * sgn;
                G = (3. / tmp) *
                    (0.5 + ((1. + tmp) * xexp(-tmp) - 1.) / (tmp * tmp));
        }

        G *= 3.;
        double aux = v * m;
        const double t = aux * aux / (1. - v);
        double aum = m;
        aum *= aum;
        aux = 2. * aum / t;
        const double kappa = 1. - 2. / v + 2. / (v * v);
        aux = G * ((kappa + 4. * aum / m1) * xlog(1. + m1 / t) -
                      (kappa * m1) / (m1 + t) - aux) +
            ((kappa + 2. * aum / m2) * xlog(1. + m2 / t) - aux) +
            aux * (G * (m1 - 4. * t) / (m1 + t) + (m2 / t) * xlog(1. + t / m2));
        aux *= ALPHA_EM / (8. * PI) * A * v * sgn;

        /* hard component: polynomials in log10(v) per decade of energy */
        const bool muon = fabs(m - MUON_MASS) < fabs(m - TAU_MASS);
        const double table_mu[7][8] = {
                { 7.174409E-04, -0.2436045, -0.2942209, -0.1658391,
                    -0.05227727, -9.328318E-03, -8.751909E-04, -3.343145E-05 },
                { 1.7132E-03, -0.5756682, -0.68615, -0.3825223, -0.1196482,
                    -0.02124577, -1.987841E-03, -7.584046E-05 },
                { 4.082304E-03, -1.553973, -2.004218, -1.207777, -0.4033373,
                    -0.07555636, -7.399682E-03, -2.943396E-04 },
                { 8.628455E-03, -3.251305, -3.999623, -2.33175, -0.7614046,
                    -0.1402496, -0.01354059, -5.3155E-04 },
                { 0.01244159, -5.976818, -6.855045, -3.88775, -1.270677,
                    -0.2370768, -0.02325118, -9.265136E-04 },
                { 0.02204591, -9.495636, -10.05705, -5.636636, -1.883845,
                    -0.3614146, -0.03629659, -1.473118E-03 },
                { 0.03228755, -13.92918, -14.37232, -8.418409, -2.948277,
                    -0.5819409, -0.059275, -2.419946E-03 } };
        const double table_tau[7][8] = {
                { -1.269205E-04, -0.01563032, 0.04693954, 0.05338546,
                    0.02240132, 4.658909E-03, 4.822364E-04, 1.9837E-05 },
                { -2.843877E-04, -0.03589573, 0.1162945, 0.130975, 0.05496,
                    0.01146659, 1.193018E-03, 4.940182E-05 },
                { -5.761546E-04, -0.07768545, 0.3064255, 0.3410341, 0.144945,
                    0.03090286, 3.302773E-03, 1.409573E-04 },
                { -1.195445E-03, -0.157375, 0.7041273, 0.7529364, 0.3119032,
                    0.06514455, 6.843364E-03, 2.877909E-04 },
                { -1.317386E-03, -0.2720009, 1.440518, 1.425927, 0.5576727,
                    0.1109868, 0.011191, 4.544877E-04 },
                { -9.689228E-15, -0.4186136, 2.533355, 2.284968, 0.8360727,
                    0.1589677, 0.015614, 6.280818E-04 },
                { -6.4595E-15, -0.8045046, 3.217832, 2.5487, 0.8085682,
                    0.1344223, 0.01173827, 4.281932E-04 } };

        const double E = K + m;
        if ((E > 1E+02) && (v > 1E-07)) {
                const double lE = xlog10(E) - 3.;
                const double lv = xlog10(v);
                int iE = (int)lE;

                double f[2] = { 0., 0. };
                for (int i = 0; i < 2; i++, iE++) {
                        /* (sic: a negative decade indexes before the table in
                         * the reference, pumas.c:15330-15336; E > 100 GeV
                         * gives lE > -1, hence iE >= 0 here) */
                        const int j = (iE > 6) ? 6 : ((iE < 0) ? 0 : iE);
                        const double x = (lv > -6.) ? lv : -6.;
                        const double * const c =
                            muon ? table_mu[j] : table_tau[j];
                        f[i] = c[0] +
                            x * (c[1] + x * (c[2] + x * (c[3] + x * (c[4] + x * (c[5] + x * (c[6] + x * c[7]))))));
                        if (lv < -6.) f[i] *= 7 + lv;
                        if (lE < 0) f[i] *= 1. + lE;
                }

                const double h = lE - (int)lE;
                double fi;
                if ((f[0] > 0.) && (f[1] > 0.))
                        fi = xexp(xlog(f[0]) * (1. - h) + xlog(f[1]) * h);
                else
                        fi = f[0] * (1. - h) + f[1] * h;
                aux += A * fi / v;
        }

        return aux * 1E-34 / (K + m);
}

} /* namespace pb */

#endif
