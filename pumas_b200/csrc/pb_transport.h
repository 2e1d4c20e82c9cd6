/*
 * pb_transport.h -- the per-track transport engine.
 *
 * Restates, one track per call (one GPU thread per track):
 *   pumas_context_transport      src/pumas.c:3114-3251   [a4]
 *   transport_with_csda          src/pumas.c:4777-4949   [a5]
 *   transport_with_stepping      src/pumas.c:5194-5514   [a6]
 *   transport_limit              src/pumas.c:5569-5696   [a7]
 *   transport_do_del / do_ehs    src/pumas.c:5706-5885   [a14, a15]
 *   step_transport               src/pumas.c:6296-6925   [a8]
 *   step_fluctuate / step_randn  src/pumas.c:6943-7029   [a9]
 *   step_rotate_direction        src/pumas.c:7042-7095   [a10]
 * The hidden per-context stepping state of the reference
 * (struct simulation_context, pumas.c:426-477) lives in `Track`, private to the
 * thread; the recorder hooks are out of scope.
 *
 * The number and order of the uniform variates drawn from `Rng` is part of the
 * contract with the oracle.
 */
#ifndef PB_TRANSPORT_H
#define PB_TRANSPORT_H

#include "pb_coulomb.h"
#include "pb_del.h"
#include "pb_media.h"
#include "pb_polar.h"

namespace pb {



/* Working precision of the Coulomb pole reduction (see pb_coulomb.h). */
#if defined(__CUDA_ARCH__) || defined(PB_FORCE_DD)
typedef dd CoulombReal;
#elif defined(PB_FORCE_QUAD)
/* (test only: IEEE binary128 as the yardstick for the two others) */
typedef __float128 CoulombReal;
#else
typedef long double CoulombReal;
#endif

/* Recorder hooks (record_state, pumas.c:7738-7780, called at pumas.c:3191,
 * 4798, 4946, 5230, 5326, 5368, 5385, 5466, 5509). The engine is a template
 * on the recorder type; the default records nothing and compiles away (zero
 * cost when `context->recorder == NULL`). `medium` is a layer index, -1 for
 * the reference's NULL medium. */
constexpr int EVENT_START = 4096;
constexpr int EVENT_STOP = 8192;
struct NoRecorder {
        static constexpr bool enabled = false;
        PB_FN int period() const { return 0; }
        PB_FN void frame(int, int, const State &) const {}
};

/* Hidden stepping state, pumas.c:426-477. */
struct Track {
        double lifetime;
        int step_first;
        int step_event;
        int step_foreseen;
        double step_X_limit;
        double step_invlb1;
        int randn_done;
        double randn_next;
        int iK;              /* grid cell of the last located kinetic energy */
        int error;           /* enum pumas_return raised inside the stepping */
        unsigned int n_steps; /* diagnostics: calls of step_transport */
};

/* pumas.c:7016-7029 */
template <class Rng> PB_FN double step_randn(Track & trk, Rng & rng)
{
        trk.randn_done = !trk.randn_done;
        if (!trk.randn_done) return trk.randn_next;

        const double r = sqrt(-2. * xlog(rng.uniform()));
        const double phi = 2. * PI * rng.uniform();
        const SinCos sc = xsincos(phi);
        trk.randn_next = r * sc.c;
        return r * sc.s;
}

/* pumas.c:7042-7095 */
template <class Rng>
PB_FN void step_rotate_direction(Rng & rng, State & state, double mu)
{
        double * const direction = state.direction;

        const double cos_theta = 1 - 2 * mu;
        const double stsq = 4 * mu * (1. - mu);
        if (stsq <= 0.) return;
        const double st = sqrt(stsq);

        /* Co-vector of the local basis, orthogonal to the direction and to the
         * axis of its smallest component. Written without branches (the
         * reference has four, pumas.c:7059-7083; same operands, same
         * operations): lanes of a warp take all four cases. */
        const double a0 = fabs(direction[0]);
        const double a1 = fabs(direction[1]);
        const double a2 = fabs(direction[2]);
        const bool case_xz = (a0 > a1) && (a0 > a2);   /* uses (x, z) */
        const bool case_xy = !(a0 > a1) && (a1 > a2);  /* uses (x, y) */
        const double pa = (case_xz || case_xy) ? direction[0] : direction[1];
        const double pb_ = case_xy ? direction[1] : direction[2];
        const double nrm = 1. / sqrt(pa * pa + pb_ * pb_);
        const double u0x = case_xz ? -direction[2] * nrm :
            (case_xy ? direction[1] * nrm : 0.);
        const double u0y = case_xz ? 0. :
            (case_xy ? -direction[0] * nrm : direction[2] * nrm);
        const double u0z = case_xz ? direction[0] * nrm :
            (case_xy ? 0. : -direction[1] * nrm);
        const double u1x = u0y * direction[2] - u0z * direction[1];
        const double u1y = u0z * direction[0] - u0x * direction[2];
        const double u1z = u0x * direction[1] - u0y * direction[0];

        const double phi = PI * (1. - 2. * rng.uniform());
        const SinCos sc = xsincos(phi);
        const double cp = sc.c;
        const double sp = sc.s;
        direction[0] = cos_theta * direction[0] + st * (cp * u0x + sp * u1x);
        direction[1] = cos_theta * direction[1] + st * (cp * u0y + sp * u1y);
        direction[2] = cos_theta * direction[2] + st * (cp * u0z + sp * u1z);
}

/* PENELOPE style straggling of the continuous loss. pumas.c:6943-7005.
 * `l0` is the located start energy; the end energy of the continuous loss is
 * located through `kl`. */
template <class Cfg, class Rng>
PB_FN void step_fluctuate(const Physics & p, const Cfg & cfg, Track & trk,
    Rng & rng, const State & state, int material, KLocator & kl, const Loc & l0,
    double Xtot, double dX, double * kf, double * ratio)
{
        const int scheme = cfg.energy_loss;
        const double sgn = (cfg.direction == MODE_FORWARD) ? 1. : -1.;
        double k1, r = 0.;
        {
                const Loc lx = locate_near(p,
                    col_sm(p, p.X, scheme, material), Xtot - sgn * dX, l0.i);
                k1 = cel_kinetic_energy_at(p, scheme, material, lx);
        }
        const double dk0 = fabs(state.energy - k1);
        if ((k1 > 0.) && (dk0 > 0.)) {
                const Loc l1 = klocate(p, kl, k1);
                const double Omega0 = cel_straggling_at(p, material, l0);
                const double Omega1 = cel_straggling_at(p, material, l1);
                double dk12 = 0.5 * dX * (Omega0 + Omega1);
                const double tmp = dX / Xtot;
                if (tmp > 5E-02) {
                        const double de1 =
                            cel_energy_loss_at(p, scheme, material, l1);
                        const double de0 =
                            cel_energy_loss_at(p, scheme, material, l0);
                        dk12 *= 1. + sgn * dX / dk0 * (de1 - de0);
                }
                const double dk1 = sqrt(dk12);
                if (dk0 >= 3. * dk1) {
                        double u;
                        do
                                u = step_randn(trk, rng);
                        while (fabs(u) > 3.);
                        u /= 1.015387;
                        k1 += u * dk1;
                } else if (dk0 >= 1.7320508 * dk1) {
                        const double u = 1.7320508 * (1. - 2. * rng.uniform());
                        k1 += u * dk1;
                } else {
                        const double dk32 = 3. * dk12;
                        const double dk02 = dk0 * dk0;
                        const double a =
                            1. - (dk32 - dk02) / (dk32 + 3. * dk02);
                        if (rng.uniform() <= a) {
                                const double b = 0.5 * (dk32 + 3. * dk02) / dk0;
                                const double u = rng.uniform();
                                k1 = state.energy - sgn * b * u;
                                if (k1 < 0.) k1 = 0.;
                        } else {
                                k1 = state.energy;
                        }
                }
                r = fabs(state.energy - k1) / dk0;
        }
        if (r == 0.) r = 1.;
        *kf = k1;
        *ratio = r;
}

/* pumas.c:5569-5696 */
template <class Cfg, class Rng>
PB_FN void transport_limit(const Physics & p, const Cfg & cfg, Track & trk,
    Rng & rng, const State & state, int material, double Xi, double Xf,
    double * grammage_max)
{
        trk.step_event = trk.step_foreseen = EVENT_NONE;

        if (state.weight <= 0.) {
                trk.step_event = EVENT_WEIGHT;
                return;
        }

        if (cfg.direction == MODE_FORWARD) {
                const double kinetic_min =
                    (cfg.event & EVENT_LIMIT_ENERGY) ? cfg.limit_energy : 0.;
                if (state.energy <= kinetic_min) {
                        trk.step_event = EVENT_LIMIT_ENERGY;
                        return;
                }
        } else if ((cfg.event & EVENT_LIMIT_ENERGY) &&
            (state.energy >= cfg.limit_energy)) {
                trk.step_event = EVENT_LIMIT_ENERGY;
                return;
        }

        if (cfg.event & EVENT_LIMIT_GRAMMAGE) {
                *grammage_max = cfg.limit_grammage;
                trk.step_foreseen = EVENT_LIMIT_GRAMMAGE;
        } else {
                *grammage_max = 0.;
        }

        const int scheme = cfg.energy_loss;
        if (scheme == MODE_DISABLED) {
                if (cfg.scattering == MODE_MIXED) {
                        const double X = Xi -
                            coulomb_ehs_length(p, material, state.energy) *
                                xlog(rng.uniform());
                        if ((*grammage_max == 0.) || (X < *grammage_max)) {
                                *grammage_max = X;
                                trk.step_foreseen = EVENT_VERTEX_COULOMB;
                                return;
                        }
                }
                return;
        }

        const double sgn = (cfg.direction == MODE_FORWARD) ? 1. : -1.;
        int foreseen = EVENT_NONE;
        double kinetic_limit = 0.;
        if (scheme == MODE_MIXED) {
                for (int i = 0; i < 1000; i++) {
                        const double zeta = rng.uniform();
                        if ((zeta <= 0.) || (zeta >= 1.)) continue;

                        const double nI =
                            del_interaction_length(p, material, state.energy) +
                            sgn * xlog(zeta);
                        if (nI <= 0.) break;

                        const double k = del_kinetic_from_interaction_length(
                            p, material, nI);

                        if (cfg.direction == MODE_FORWARD) {
                                if (k >= state.energy) continue;
                        } else {
                                if (k <= state.energy) continue;
                        }

                        if ((cfg.direction == MODE_BACKWARD) ||
                            (k > p.Kt[material])) {
                                kinetic_limit = k;
                                foreseen = EVENT_VERTEX_DEL;
                        }
                        break;
                }
        }

        if ((scheme < MODE_STRAGGLED) && (cfg.scattering == MODE_MIXED)) {
                const double nI =
                    ehs_interaction_length(p, scheme, material, state.energy) +
                    sgn * xlog(rng.uniform());
                if (nI > 0.) {
                        const double k = ehs_kinetic_from_interaction_length(
                            p, scheme, material, nI);
                        if ((kinetic_limit <= 0.) ||
                            ((cfg.direction == MODE_FORWARD) &&
                                (k > kinetic_limit)) ||
                            ((cfg.direction == MODE_BACKWARD) &&
                                (k < kinetic_limit))) {
                                kinetic_limit = k;
                                foreseen = EVENT_VERTEX_COULOMB;
                        }
                }
        }

        if (foreseen == EVENT_NONE) return;

        const double X = Xi +
            sgn * (Xf - cel_grammage(p, scheme, material, kinetic_limit));
        if ((*grammage_max <= 0) || (X < *grammage_max)) {
                *grammage_max = X;
                trk.step_foreseen = foreseen;
        }
}

/* pumas.c:5723-5757: the event flag of a hard inelastic vertex and its polar
 * deflection. */
template <class Cfg, class Rng>
PB_FN void transport_do_del_tail(const Physics & p, const Cfg & cfg,
    Track & trk, Rng & rng, State & state, int process, int element,
    bool polar, double ki, double kf)
{
        if (process < 0) {
                if (state.weight > 0)
                        trk.step_event = EVENT_NONE;
                else
                        trk.step_event = EVENT_WEIGHT;
        } else {
                trk.step_event = (process == 0) ?
                    EVENT_VERTEX_BREMSSTRAHLUNG :
                    (process == 1) ?
                    EVENT_VERTEX_PAIR_CREATION :
                    (process == 2) ?
                    EVENT_VERTEX_PHOTONUCLEAR :
                    (process == 3) ? EVENT_VERTEX_DELTA_RAY : EVENT_VERTEX;
        }

        if ((cfg.scattering == MODE_MIXED) && polar) {
                if (kf > ki) {
                        const double tmp = kf;
                        kf = ki;
                        ki = tmp;
                }
                const double mu =
                    polar_sample(p, rng, process, p.element[element], ki, kf);
                step_rotate_direction(rng, state, mu);
        }
}

/* pumas.c:5706-5757 */
template <class Cfg, class Rng>
PB_FN void transport_do_del(const Physics & p, const Cfg & cfg,
    Track & trk, Rng & rng, State & state, int material)
{
        double ki, kf;
        int process = -1;
        int element = 0;
        bool polar;
        if (cfg.direction == MODE_FORWARD) {
                ki = state.energy;
                polar = del_randomise_forward(
                    p, cfg, rng, state, material, &process, &element);
                kf = state.energy;
        } else {
                kf = state.energy;
                polar = del_randomise_reverse(
                    p, cfg, rng, state, material, &process, &element);
                ki = state.energy;
        }
        transport_do_del_tail(
            p, cfg, trk, rng, state, process, element, polar, ki, kf);
}

/* Restricted (hard) cross-section of one component, pumas.c:5786-5813. */
PB_NOINLINE double coulomb_cs_hard(const Physics & p, int component,
    double kinetic, double mu0)
{
        const Element & el = p.element[p.comp_element[component]];
        CoulombData data;
        double kinetic0;
        coulomb_data_set(p, el, kinetic, &data, &kinetic0);
        CoulombPoles<CoulombReal> poles;
        coulomb_pole_reduction<CoulombReal>(
            data.n_parameters, data.amplitude, data.screening, &poles);
        const double normalisation = p.comp_fraction[component] /
            coulomb_normalisation(el.Z, el.A, p.mass, kinetic, kinetic0);
        return normalisation *
            coulomb_restricted_cs<CoulombReal>(
                mu0, data.fspin, data.n_parameters, data.screening, poles);
}

/* pumas.c:5768-5885 */
template <class Cfg, class Rng>
PB_FN void transport_do_ehs(const Physics & p, const Cfg & cfg, Rng & rng,
    State & state, int material, int hint)
{
        const double kinetic = state.energy;

        double mu0 = 0., invlb1 = 0.;
        table_get_msc_at(p, cfg.energy_loss, material,
            locate_near(p, p.K, kinetic, hint), &mu0, &invlb1);

        /* Randomise the hard scatterer element. For a single element material
         * the outcome is forced, whatever the cross-section; the variate is
         * still drawn (pumas.c:5816). */
        const int n_el = p.elements_in[material];
        const int ic0 = p.comp_offset[material];
        int ihard = n_el - 1;
        if (n_el == 1) {
                (void)rng.uniform();
                ihard = 0;
        } else {
                double cs_hard[MAX_COMPONENTS];
                double cs_tot = 0.;
                int i;
                for (i = 0; i < n_el; i++) {
                        cs_hard[i] = coulomb_cs_hard(p, ic0 + i, kinetic, mu0);
                        cs_tot += cs_hard[i];
                }
                const double cc0 = rng.uniform() * cs_tot;
                double cs = 0.;
                for (i = 0; i < n_el; i++) {
                        cs += cs_hard[i];
                        if (cs >= cc0) {
                                ihard = i;
                                break;
                        }
                }
        }

        /* Rejection sampling of the angular parameter under a Wentzel DCS. */
        CoulombData data;
        double kinetic0;
        coulomb_data_set(
            p, p.element[p.comp_element[ic0 + ihard]], kinetic, &data, &kinetic0);
        const int n = data.n_parameters - 1;
        double A = data.screening[0];
        int i;
        for (i = 1; i < n; i++) {
                const double Ai = data.screening[i];
                if (Ai < A) A = Ai;
        }
        double mu1;
        for (;;) {
                const double zeta = rng.uniform();
                const double tmp = 1. + A - zeta * (1 - mu0);
                if (tmp <= 0.) continue;
                mu1 = (A + mu0) * (A + 1.) / tmp - A;
                if (mu1 < mu0)
                        mu1 = mu0;
                else if (mu1 > 1.)
                        mu1 = 1;

                double ratio = 0.;
                for (i = 0; i < n; i++) {
                        ratio += data.amplitude[i] * (A + mu1) /
                            (data.screening[i] + mu1);
                }
                ratio *= coulomb_nuclear_form_factor(mu1, data.screening[n]);
                ratio *= ratio * (1. - data.fspin * mu1);

                if (rng.uniform() <= ratio) break;
        }

        /* CM to laboratory frame. */
        const double gamma = data.fCM[0];
        const double tau = data.fCM[1];
        double mu;
        if (mu1 > 1E-06) {
                const double a = gamma * (tau + 1. - 2. * mu1);
                const double ct_h = a / sqrt(4. * mu1 * (1. - mu1) + a * a);
                mu = 0.5 * (1. - ct_h);
        } else {
                const double d = gamma * (1. + tau);
                mu = mu1 / (d * d);
        }
        step_rotate_direction(rng, state, mu);
}

/* Alignment points of the step. A GPU scheduler that runs many warps through
 * step_transport together may re-align them at the `sync.point(k)` marks
 * (resident.cuh): the step is larger than the instruction cache, and warps that
 * stay close to each other share the fetched lines. No effect on results; the
 * default does nothing. The marks sit where every thread of a step passes
 * exactly once (no early return before the last one in a uniform medium). */
constexpr int N_SYNC_POINTS = 5;
struct NoSync {
        PB_FN void point(int) const {}
};

/* One Monte Carlo step in a single medium. pumas.c:6296-6925. Returns false
 * on error (trk.error set). */
template <class Cfg, class Geo, class Rng, class Sync = NoSync>
PB_FN bool step_transport(const Physics & p, const Cfg & cfg,
    const Geo & geo, Track & trk, Rng & rng, State & state, int straight,
    int medium, Locals & locals, double grammage_max, double step_max_medium,
    int step_max_type, double * step_max_locals, int * out_medium,
    const Sync & sync = Sync())
{
        trk.n_steps++;
        const int scheme = cfg.energy_loss;
        double * const direction = state.direction;
        double * const position = state.position;
        const int material = geo_material(geo, medium);
        const double density = locals.density;
        const double density_i = 1. / density;
        const double momentum = sqrt(state.energy * (state.energy + 2. * p.mass));

        /* The energies of this step are each located once in the grid. */
        KLocator kl;
        klocator_init(kl, trk.iK);
        const Loc l0 = klocate(p, kl, state.energy);
        const double * const Xcol = col_sm(p, p.X,
            (scheme == MODE_DISABLED) ? MODE_CSDA : scheme, material);

        const int tmp_scheme = (scheme == MODE_DISABLED) ? MODE_CSDA : scheme;
        const double Xtot = cel_grammage_at(p, tmp_scheme, material, l0);

        /* Local step length. */
        double step_loc, rLarmor0 = 0., uT0[3] = { 0., 0., 0. };
        double invlb1 = 0.;
        if (!straight) {
                double r = cfg.accuracy;
                const double k_threshold = 1E+09;
                if ((scheme == MODE_STRAGGLED) && (state.energy > k_threshold)) {
                        double f = k_threshold / state.energy;
                        if (f < 0.1) f = 0.1;
                        r *= f;
                }
                step_loc = r * density_i * Xtot;

                if (cfg.scattering == MODE_MIXED) {
                        if (trk.step_first != 0) {
                                double mu0;
                                table_get_msc_at(p, cfg.energy_loss, material,
                                    l0, &mu0, &invlb1);
                                invlb1 *= density;
                        } else
                                invlb1 = trk.step_invlb1;
                        const double stepT = cfg.accuracy / invlb1;
                        if (stepT < step_loc) step_loc = stepT;
                }

                if (!geo_has_locals(geo) || (locals.magnetized == 0)) {
                        rLarmor0 = 0.;
                } else {
                        const double * const B = locals.magnet;
                        uT0[0] = direction[1] * B[2] - direction[2] * B[1];
                        uT0[1] = direction[2] * B[0] - direction[0] * B[2];
                        uT0[2] = direction[0] * B[1] - direction[1] * B[0];
                        const double BT = sqrt(
                            uT0[0] * uT0[0] + uT0[1] * uT0[1] + uT0[2] * uT0[2]);
                        if (BT == 0.) {
                                rLarmor0 = 0.;
                        } else {
                                rLarmor0 = momentum / (BT * LARMOR_FACTOR);
                                const double pinv = 1. / momentum;
                                uT0[0] *= pinv;
                                uT0[1] *= pinv;
                                uT0[2] *= pinv;
                        }
                }

                if (rLarmor0 > 0.) {
                        const double stepM = cfg.accuracy * rLarmor0;
                        if (stepM < step_loc) step_loc = stepM;
                }
        } else {
                if (grammage_max <= 0.)
                        step_loc = 1E+09 * density_i;
                else
                        step_loc = grammage_max * density_i;
        }

        trk.step_first = 0;

        if (step_loc < STEP_MIN) step_loc = STEP_MIN;
        double step = (step_max_medium <= 0) ?
            step_loc :
            (step_max_medium < step_loc ? step_max_medium : step_loc);
        if ((*step_max_locals > 0.) && (step > *step_max_locals))
                step = *step_max_locals;

        *out_medium = medium;
        int event = EVENT_NONE;
        if (cfg.event & EVENT_LIMIT_DISTANCE) {
                const double d = cfg.limit_distance - state.distance;
                if (d <= step) {
                        step = d;
                        event = EVENT_LIMIT_DISTANCE;
                }
        }

        const double sgn = (cfg.direction == MODE_FORWARD) ? 1. : -1.;
        position[0] += step * sgn * direction[0];
        position[1] += step * sgn * direction[1];
        position[2] += step * sgn * direction[2];

        /* Change of medium? */
        int end_medium = -1;
        medium_locate(geo, cfg.direction, state, true, &end_medium, false, 0);
        double end_position[3] = { position[0], position[1], position[2] };
        if (geo_has_boundaries(geo) && (end_medium != medium)) {
                if (step_max_type == STEP_CHECK) {
                        const double step_min =
                            (step < STEP_MIN) ? step : STEP_MIN;
                        double pi[3] = { position[0], position[1],
                                position[2] };
                        double s1 = 0., s2 = -step_min;
                        position[0] = pi[0] + s2 * sgn * direction[0];
                        position[1] = pi[1] + s2 * sgn * direction[1];
                        position[2] = pi[2] + s2 * sgn * direction[2];
                        int tmp_medium = -1;
                        medium_locate(geo, cfg.direction, state, true,
                            &tmp_medium, false, 0);
                        if (tmp_medium != medium) {
                                if (tmp_medium != end_medium)
                                        end_medium = tmp_medium;
                                s1 = s2;
                                s2 = -step;
                                while (fabs(s1 - s2) > STEP_MIN) {
                                        const double s3 = 0.5 * (s1 + s2);
                                        position[0] =
                                            pi[0] + s3 * sgn * direction[0];
                                        position[1] =
                                            pi[1] + s3 * sgn * direction[1];
                                        position[2] =
                                            pi[2] + s3 * sgn * direction[2];
                                        tmp_medium = -1;
                                        medium_locate(geo, cfg.direction, state,
                                            true, &tmp_medium, false, 0);
                                        if (tmp_medium == medium) {
                                                s2 = s3;
                                        } else {
                                                s1 = s3;
                                                if (tmp_medium != end_medium)
                                                        end_medium = tmp_medium;
                                        }
                                }
                                position[0] = pi[0] + s2 * sgn * direction[0];
                                position[1] = pi[1] + s2 * sgn * direction[1];
                                position[2] = pi[2] + s2 * sgn * direction[2];
                                step += s1;
                                end_position[0] =
                                    pi[0] + s1 * sgn * direction[0];
                                end_position[1] =
                                    pi[1] + s1 * sgn * direction[1];
                                end_position[2] =
                                    pi[2] + s1 * sgn * direction[2];
                                /* (The reference calls the medium function
                                 * once more here; it has no side effect.) */
                        }
                }
                event = EVENT_MEDIUM;
                *out_medium = end_medium;
        }

        /* End step locals. */
        double Bi[3] = { 0., 0., 0. };
        if (geo_has_locals(geo) && (locals.magnetized != 0)) {
                Bi[0] = locals.magnet[0];
                Bi[1] = locals.magnet[1];
                Bi[2] = locals.magnet[2];
        }
        if (geo_has_locals(geo) && (*step_max_locals > 0.) &&
            ((step_max_type != STEP_RAW) || (event != EVENT_MEDIUM))) {
                *step_max_locals =
                    set_locals(p, geo, cfg.accuracy, medium, state, &locals);
                if (locals.density <= 0.) {
                        trk.error = RETURN_DENSITY_ERROR;
                        return false;
                }
        }

        if (event & EVENT_MEDIUM) {
                position[0] = end_position[0];
                position[1] = end_position[1];
                position[2] = end_position[2];
        }

        sync.point(0);
        /* End step kinetic energy. */
        double k1 = state.energy, dk = 0.;
        const double dX = 0.5 * step * (density + locals.density);
        if ((scheme >= MODE_CSDA) && (scheme <= MODE_MIXED)) {
                const double X = Xtot - sgn * dX;
                if ((cfg.direction == MODE_FORWARD) && (X <= trk.step_X_limit)) {
                        k1 = (cfg.event & EVENT_LIMIT_ENERGY) ?
                            cfg.limit_energy :
                            0.;
                        const double grammage =
                            state.grammage + Xtot - trk.step_X_limit;
                        if ((grammage_max <= 0.) || (grammage < grammage_max)) {
                                grammage_max = grammage;
                                event = EVENT_LIMIT_ENERGY;
                        }
                } else if ((cfg.direction == MODE_BACKWARD) &&
                    (trk.step_X_limit > 0.) && (X > trk.step_X_limit)) {
                        k1 = (cfg.event & EVENT_LIMIT_ENERGY) ?
                            cfg.limit_energy :
                            0.;
                        const double grammage =
                            state.grammage + trk.step_X_limit - Xtot;
                        if ((grammage_max <= 0.) || (grammage < grammage_max)) {
                                grammage_max = grammage;
                                event = EVENT_LIMIT_ENERGY;
                        }
                } else if (dX < EPSILON_X * Xtot) {
                        k1 -= sgn *
                            cel_energy_loss_at(p, scheme, material, l0) * dX;
                        if ((cfg.event & EVENT_LIMIT_ENERGY) &&
                            (cfg.limit_energy > 0.)) {
                                if (((cfg.direction == MODE_FORWARD) &&
                                        (k1 < cfg.limit_energy)) ||
                                    ((cfg.direction == MODE_BACKWARD) &&
                                        (k1 > cfg.limit_energy))) {
                                        k1 = cfg.limit_energy;
                                        event = EVENT_LIMIT_ENERGY;
                                }
                        } else if (k1 < 0.) {
                                k1 = 0.;
                        }
                } else {
                        k1 = cel_kinetic_energy_at(p, scheme, material,
                            locate_near(p, Xcol, X, l0.i));
                }
        } else if (scheme == MODE_STRAGGLED) {
                double ratio;
                step_fluctuate(p, cfg, trk, rng, state, material, kl, l0, Xtot,
                    dX, &k1, &ratio);
                dk = fabs(state.energy - k1);

                double kinetic_limit = -1.;
                if (cfg.direction == MODE_FORWARD) {
                        const double kinetic_min =
                            (cfg.event & EVENT_LIMIT_ENERGY) ?
                            cfg.limit_energy :
                            0.;
                        if (k1 <= kinetic_min) kinetic_limit = kinetic_min;
                } else {
                        if ((cfg.event & EVENT_LIMIT_ENERGY) &&
                            (k1 >= cfg.limit_energy))
                                kinetic_limit = cfg.limit_energy;
                }
                if (kinetic_limit >= 0.) {
                        const double grammage = state.grammage +
                            fabs(state.energy - kinetic_limit) / dk * dX;
                        k1 = kinetic_limit;
                        if ((grammage_max <= 0.) || (grammage < grammage_max)) {
                                grammage_max = grammage;
                                event = EVENT_LIMIT_ENERGY;
                        }
                }

                sync.point(1);
                /* Discrete events by thinning. */
                double Xmax = dX;
                if (grammage_max > 0.) {
                        const double dX1 = grammage_max - state.grammage;
                        if (Xmax > dX1) Xmax = dX1;
                }
                double k_x = -1.;

                do {
                        double xs_del = del_cross_section_at(p, material, l0);
                        const double tmp_del = del_cross_section_at(
                            p, material, klocate(p, kl, k1));
                        if (tmp_del > xs_del) xs_del = tmp_del;
                        if (xs_del <= 0.) break;
                        const double X_del = -xlog(rng.uniform()) / xs_del;
                        if (X_del >= Xmax) break;
                        const double k_h = cel_kinetic_energy_at(p, scheme,
                            material,
                            locate_near(p, Xcol, Xtot - sgn * X_del, l0.i));
                        const double k_del = state.energy -
                            sgn * ratio * fabs(state.energy - k_h);
                        if (k_del <= 0.) break;
                        const double r = del_cross_section_at(p, material,
                                             klocate(p, kl, k_del)) /
                            xs_del;
                        if (rng.uniform() > r) break;
                        k_x = k_del;
                        Xmax = X_del;
                        event = EVENT_VERTEX_DEL;
                } while (0);

                if (cfg.scattering == MODE_MIXED) {
                        do {
                                double kmin, kmax;
                                if (k1 <= state.energy)
                                        kmin = k1, kmax = state.energy;
                                else
                                        kmax = k1, kmin = state.energy;
                                if (kmin < p.K[1]) {
                                        const double tmp = kmin;
                                        kmin = kmax, kmax = tmp;
                                }
                                (void)kmax;
                                const double lb_ehs = coulomb_ehs_length_at(
                                    p, material, klocate(p, kl, kmin));
                                if (lb_ehs <= 0.) break;
                                const double X_ehs =
                                    -xlog(rng.uniform()) * lb_ehs;
                                if (X_ehs >= Xmax) break;
                                const double k_h = cel_kinetic_energy_at(p,
                                    scheme, material,
                                    locate_near(
                                        p, Xcol, Xtot - sgn * X_ehs, l0.i));
                                const double k_ehs = state.energy -
                                    sgn * ratio * fabs(state.energy - k_h);
                                if (k_ehs <= 0.) break;
                                const double r =
                                    coulomb_ehs_length_at(p, material,
                                        klocate(p, kl, k_ehs)) /
                                    lb_ehs;
                                if ((r <= 0.) || (rng.uniform() > 1. / r))
                                        break;
                                k_x = k_ehs;
                                Xmax = X_ehs;
                                event = EVENT_VERTEX_COULOMB;
                        } while (0);
                }

                if (k_x >= 0.) {
                        k1 = k_x;
                        grammage_max = state.grammage + Xmax;
                }
        }

        if (scheme != MODE_STRAGGLED) sync.point(1);
        sync.point(2);
        /* Grammage. */
        const double Xi = state.grammage;
        state.grammage += dX;
        const double sf0 = step;
        double h_int = 0.;
        if ((grammage_max > 0.) && (state.grammage >= grammage_max)) {
                const double dX_ = grammage_max - Xi;
                if ((fabs(density - locals.density) <= (double)FLT_EPSILON) ||
                    (sf0 <= (double)FLT_EPSILON)) {
                        step = dX_ * density_i;
                        h_int = dX_ / (state.grammage - Xi);
                } else {
                        const double drho = locals.density - density;
                        double tmp = density * density + 2. * dX_ * drho / sf0;
                        tmp = (tmp > (double)FLT_EPSILON) ? sqrt(tmp) : 0;
                        h_int = (tmp - density) / drho;
                        step *= h_int;
                }
                state.grammage = grammage_max;
                if (!(event & EVENT_LIMIT_ENERGY)) {
                        if (scheme <= MODE_MIXED) {
                                if (scheme != MODE_DISABLED)
                                        k1 = cel_kinetic_energy_at(p, scheme,
                                            material,
                                            locate_near(p, Xcol,
                                                Xtot -
                                                    sgn *
                                                        (state.grammage - Xi),
                                                l0.i));
                                event = trk.step_foreseen;
                        } else if (!(event &
                                       (EVENT_VERTEX_COULOMB |
                                           EVENT_VERTEX_DEL))) {
                                k1 = state.energy -
                                    sgn * (state.grammage - Xi) * dk / dX;
                                if (k1 < 0.) k1 = 0.;
                                event = EVENT_LIMIT_GRAMMAGE;
                        }
                }

                const double ds_ = step - sf0;
                position[0] += ds_ * sgn * direction[0];
                position[1] += ds_ * sgn * direction[1];
                position[2] += ds_ * sgn * direction[2];
        }

        /* Proper time. */
        int decayed = 0;
        double time_max = (cfg.event & EVENT_LIMIT_TIME) ? cfg.limit_time : 0.;
        if (cfg.decay == MODE_RANDOMISED) {
                if ((time_max <= 0.) || (trk.lifetime < time_max)) {
                        time_max = trk.lifetime;
                        decayed = 1;
                }
        }

        const double sf1 = step;
        if (straight && (scheme != MODE_DISABLED)) {
                const double Ti =
                    cel_proper_time(p, scheme, material, state.energy);
                if (time_max > 0.) {
                        const double Tf =
                            Ti - sgn * (time_max - state.time) * density;
                        if (Tf > 0.) {
                                const double dxT = fabs(Xtot -
                                    cel_grammage_as_time(
                                        p, scheme, material, Tf));
                                if (Xi + dxT < state.grammage) {
                                        event = EVENT_LIMIT_TIME;
                                        state.time = time_max;
                                        state.decayed = decayed;
                                        step = dxT * density_i;
                                        if (step > sf1) step = sf1;
                                        state.grammage = Xi + dxT;
                                        const double xf = Xtot - sgn * dxT;
                                        k1 = cel_kinetic_energy(
                                            p, scheme, material, xf);
                                }
                        }
                }

                if (event != EVENT_LIMIT_TIME) {
                        const double Tf =
                            cel_proper_time(p, scheme, material, k1);
                        state.time += fabs(Tf - Ti) * density_i;
                }
        } else {
                const double p_f =
                    (k1 <= 0) ? momentum : sqrt(k1 * (k1 + 2. * p.mass));
                const double ti = state.time;
                state.time += 0.5 * step * p.mass * (1. / momentum + 1. / p_f);
                if ((time_max > 0.) && (state.time >= time_max)) {
                        event = EVENT_LIMIT_TIME;
                        state.time = time_max;
                        state.decayed = decayed;

                        const double a = (momentum / p_f - 1.) / sf1;
                        const double c =
                            -2. * (time_max - ti) * momentum / p.mass;
                        if (a != 0.) {
                                double delta = 1. - a * c;
                                if (delta < 0.) delta = 0.;
                                step = 1. / a * (sqrt(delta) - 1.);
                        } else {
                                step = -0.5 * c;
                        }
                        if (step < 0.) step = 0.;

                        if (scheme != MODE_DISABLED) {
                                const double p1 = momentum / (1 + a * step);
                                k1 = sqrt(p1 * p1 - p.mass * p.mass) - p.mass;
                                if (k1 < 0.) k1 = 0.;
                        }

                        const double Xf = Xi + dX;
                        if (density == locals.density)
                                state.grammage = Xi + step * density;
                        else
                                state.grammage = Xi +
                                    step *
                                        (density +
                                            0.5 * step / sf0 *
                                                (locals.density - density));
                        if (state.grammage > Xf) state.grammage = Xf;
                }
        }

        if (event == EVENT_LIMIT_TIME) {
                const double ds_ = step - sf1;
                position[0] += ds_ * sgn * direction[0];
                position[1] += ds_ * sgn * direction[1];
                position[2] += ds_ * sgn * direction[2];
        }

        trk.step_event = event;
        state.energy = k1;
        state.distance += step;

        sync.point(3);
        /* End of step multiple scattering path length. */
        if (cfg.scattering == MODE_MIXED) {
                double mu0, invlb1_;
                if (state.energy <= 0.) {
                        trk.step_invlb1 = invlb1;
                } else {
                        table_get_msc_at(p, cfg.energy_loss, material,
                            klocate(p, kl, state.energy), &mu0, &invlb1_);
                        trk.step_invlb1 = locals.density * invlb1_;
                }
        }
        trk.iK = kl.hint;

        /* End of step magnetic field. */
        double rLarmor1 = 0., uT1[3] = { 0., 0., 0. };
        if (geo_has_locals(geo) && locals.magnetized) {
                double B[3] = { locals.magnet[0], locals.magnet[1],
                        locals.magnet[2] };
                int magnetized = locals.magnetized;
                if (h_int > 0.) {
                        B[0] = Bi[0] + h_int * (B[0] - Bi[0]);
                        B[1] = Bi[1] + h_int * (B[1] - Bi[1]);
                        B[2] = Bi[2] + h_int * (B[2] - Bi[2]);
                        /* Sic: the reference inverts the flag here
                         * (pumas.c:6843-6844); kept for parity. */
                        magnetized =
                            (B[0] * B[0] + B[1] * B[1] + B[2] * B[2] == 0.);
                }

                if (magnetized == 0) {
                        rLarmor1 = 0.;
                        uT1[0] = uT1[1] = uT1[2] = 0.;
                } else {
                        uT1[0] = direction[1] * B[2] - direction[2] * B[1];
                        uT1[1] = direction[2] * B[0] - direction[0] * B[2];
                        uT1[2] = direction[0] * B[1] - direction[1] * B[0];
                        const double BT = sqrt(
                            uT1[0] * uT1[0] + uT1[1] * uT1[1] + uT1[2] * uT1[2]);
                        if (BT == 0.) {
                                rLarmor1 = 0.;
                        } else {
                                const double pp = (state.energy <= 0) ?
                                    momentum :
                                    sqrt(state.energy *
                                        (state.energy + 2. * p.mass));
                                rLarmor1 = pp / (BT * LARMOR_FACTOR);
                                const double pinv = 1. / pp;
                                uT1[0] *= pinv;
                                uT1[1] *= pinv;
                                uT1[2] *= pinv;
                        }
                }
        }

        /* Magnetic deflection. */
        if (geo_has_locals(geo) && ((rLarmor0 > 0.) || (rLarmor1 > 0.))) {
                const double u[3] = { direction[0], direction[1],
                        direction[2] };
                const double theta0 =
                    (rLarmor0 > 0.) ? state.charge * step / rLarmor0 : 0.;
                const double theta1 =
                    (rLarmor1 > 0.) ? state.charge * step / rLarmor1 : 0.;
                double theta = 0.5 * (theta0 + theta1);
                if (cfg.direction == MODE_BACKWARD) theta = -theta;
                const SinCos sc = xsincos(theta);
                double c = sc.c;
                double s = sc.s;
                double v[3];
                int i;
                for (i = 0; i < 3; i++) v[i] = 0.5 * (uT0[i] + uT1[i]);
                double nrmv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
                if (nrmv > 0.) {
                        nrmv = 1. / sqrt(nrmv);
                        for (i = 0; i < 3; i++) v[i] *= nrmv;
                } else {
                        c = 1.;
                        s = 0.;
                }
                for (i = 0; i < 3; i++) direction[i] = c * u[i] + s * v[i];

                const double norm = 1. /
                    sqrt(direction[0] * direction[0] +
                        direction[1] * direction[1] +
                        direction[2] * direction[2]);
                for (i = 0; i < 3; i++) direction[i] *= norm;
        }

        sync.point(4);
        /* Soft multiple scattering. */
        if ((invlb1 > 0.) || (trk.step_invlb1 > 0.)) {
                double ilb1 = 0.25 * step * (invlb1 + trk.step_invlb1);
                if (ilb1 > 1.) ilb1 = 1.;
                double mu;
                do {
                        mu = -ilb1 * xlog(rng.uniform());
                } while (mu > 1.);
                step_rotate_direction(rng, state, mu);
        }

        return true;
}

template <class Cfg>
PB_FN int straight_flag(
    const Cfg & cfg, double step_max_locals, const Locals & locals)
{
        return ((cfg.scattering == MODE_DISABLED) &&
                   (cfg.energy_loss <= MODE_MIXED) && (step_max_locals <= 0.) &&
                   !locals.magnetized) ?
            1 :
            0;
}

/*
 * transport_with_stepping (pumas.c:5194-5514) as a resumable state machine.
 *
 * The reference runs one track to completion inside nested loops. Here the
 * loop is cut at the points where a track needs a different kind of work, so
 * that a GPU scheduler can regroup tracks by pending work (kernels.cu):
 *   stepping_begin   the prologue up to the first transport_limit (:5210-5273)
 *   stepping_step    one loop iteration up to the event dispatch (:5277-5335,
 *                    medium changes :5417-5467 handled in place)
 *   stepping_del     a hard inelastic event, transport_do_del (:5340-5392)
 *   stepping_ehs     a hard elastic event, transport_do_ehs (:5393-5402)
 *   stepping_rearm   re-arm after an event (:5404-5416, :5476-5488)
 *   stepping_finish  final protections (:5492-5513)
 * `Loop` holds the locals of the reference function that live across steps.
 * The CPU build and the simple kernel just chain the pieces (transport_track).
 */
enum Phase {
        PHASE_STEP = 0, /* needs a Monte Carlo step */
        PHASE_EHS = 1,  /* pending hard elastic event */
        PHASE_DEL = 2,  /* pending hard inelastic event */
        PHASE_DONE = 3  /* transport over, trk.step_event is the end event */
};

struct Loop {
        int medium;
        int material;
        int straight;
        int step_max_type;
        int first; /* step_index == 1 */
        double ti, wi, Xi, Xf, dei;
        double grammage_max;
        double step_max_medium;
        double step_max_locals;
        Locals locals;
};

/* The per-track values that cannot change in a given configuration, re-stated
 * from the configuration itself. In a single uniform medium without locals the
 * density, the (null) field, the medium, its material and the step limits are
 * the same for every track and every step: a GPU scheduler calls this before a
 * step so that the compiler sees them as constants of the kernel instead of
 * keeping 20 registers per track for them (-23 % register spills, +5 %
 * tracks/s). Values are the ones the track already holds (set_locals,
 * medium_locate and straight_flag above give nothing else). */
template <class Cfg, class Geo>
PB_FN void loop_canonical(
    const Physics &, const Cfg & cfg, const Geo &, Track & trk, Loop &)
{
        /* (the lifetime is only drawn, and read, for randomised decays) */
        if (cfg.decay != MODE_RANDOMISED) trk.lifetime = 0.;
}

template <class Cfg>
PB_FN void loop_canonical(const Physics & p, const Cfg & cfg,
    const UniformGeometry & g, Track & trk, Loop & lp)
{
        if (cfg.decay != MODE_RANDOMISED) trk.lifetime = 0.;
        lp.medium = 0;
        lp.material = g.material;
        lp.step_max_type = g.step_type;
        lp.step_max_medium = 0.;
        lp.step_max_locals = 0.;
        lp.locals.density = uniform_density(p, g);
        lp.locals.magnet[0] = lp.locals.magnet[1] = lp.locals.magnet[2] = 0.;
        lp.locals.magnetized = 0;
        lp.straight = straight_flag(cfg, 0., lp.locals);
}

/* pumas.c:5492-5513 */
template <class Cfg>
PB_FN void stepping_finish(const Cfg & cfg, State & state)
{
        if (cfg.direction == MODE_FORWARD) {
                const double kinetic_min =
                    (cfg.event & EVENT_LIMIT_ENERGY) ? cfg.limit_energy : 0.;
                if (fabs(state.energy - kinetic_min) < (double)FLT_EPSILON)
                        state.energy = kinetic_min;
        } else {
                if ((cfg.event & EVENT_LIMIT_ENERGY) &&
                    (fabs(state.energy - cfg.limit_energy) <
                        (double)FLT_EPSILON))
                        state.energy = cfg.limit_energy;
        }
}

/* pumas.c:5210-5273. Returns PHASE_STEP, or PHASE_DONE when the track stops
 * before its first step (then `*final` tells whether the final protections
 * apply: the reference returns early, skipping them, at :5250-5273). */
template <class Cfg, class Geo, class Rng, class Rec = NoRecorder>
PB_FN int stepping_begin(const Physics & p, const Cfg & cfg,
    const Geo & geo, Track & trk, Rng & rng, State & state, Loop & lp,
    const Rec & rec = Rec())
{
        lp.material = geo_material(geo, lp.medium);
        const int scheme = cfg.energy_loss;
        lp.straight = straight_flag(cfg, lp.step_max_locals, lp.locals);
        lp.first = 1;

        trk.step_event = EVENT_NONE;
        /* the start of the track, pumas.c:5229-5231 */
        if (Rec::enabled)
                rec.frame(lp.medium, trk.step_event | EVENT_START, state);
        lp.ti = state.time;
        lp.wi = state.weight;
        lp.Xi = state.grammage;
        if (scheme > MODE_DISABLED) {
                const Loc li = locate_near(p, p.K, state.energy, trk.iK);
                lp.Xf = cel_grammage_at(p, scheme, lp.material, li);
                lp.dei = 1. / cel_energy_loss_at(p, scheme, lp.material, li);
        } else {
                lp.Xf = lp.dei = 0.;
        }
        lp.grammage_max = 0.;

        int early = EVENT_NONE;
        if ((cfg.event & EVENT_LIMIT_DISTANCE) &&
            (state.distance >= cfg.limit_distance))
                early = EVENT_LIMIT_DISTANCE;
        else if ((cfg.event & EVENT_LIMIT_GRAMMAGE) &&
            (state.grammage >= cfg.limit_grammage))
                early = EVENT_LIMIT_GRAMMAGE;
        else if ((cfg.event & EVENT_LIMIT_TIME) &&
            (state.time >= cfg.limit_time))
                early = EVENT_LIMIT_TIME;
        else if (state.weight <= 0.)
                early = EVENT_WEIGHT;
        if (early != EVENT_NONE) {
                trk.step_event = early;
                return PHASE_DONE;
        }

        trk.step_event = EVENT_NONE;
        trk.step_first = 1;
        trk.step_X_limit = (cfg.event & EVENT_LIMIT_ENERGY) ?
            cel_grammage(p, (scheme > MODE_DISABLED) ? scheme : MODE_CSDA,
                lp.material, cfg.limit_energy) :
            0.;
        trk.step_invlb1 = 0;
        transport_limit(p, cfg, trk, rng, state, lp.material, lp.Xi, lp.Xf,
            &lp.grammage_max);
        /* (an early return of the reference, pumas.c:5273: no STOP frame) */
        if (trk.step_event) return PHASE_DONE;
        return PHASE_STEP;
}

/* pumas.c:5404-5416 and 5476-5488: refresh the locals after a vertex, update
 * the initial conditions and draw the next limits. */
/* The end of the stepping loop: final protections then the STOP frame,
 * pumas.c:5492-5510. */
template <class Cfg, class Rec>
PB_FN int stepping_stop(const Cfg & cfg, const Track & trk, State & state,
    const Loop & lp, const Rec & rec)
{
        stepping_finish(cfg, state);
        if (Rec::enabled)
                rec.frame(lp.medium, trk.step_event | EVENT_STOP, state);
        return PHASE_DONE;
}

template <class Cfg, class Geo, class Rng, class Rec = NoRecorder>
PB_FN int stepping_rearm(const Physics & p, const Cfg & cfg,
    const Geo & geo, Track & trk, Rng & rng, State & state, Loop & lp,
    bool after_vertex, const Rec & rec = Rec())
{
        if (geo_has_locals(geo) && after_vertex && (lp.step_max_locals > 0.)) {
                lp.step_max_locals = set_locals(
                    p, geo, cfg.accuracy, lp.medium, state, &lp.locals);
                if (lp.locals.density <= 0.) {
                        trk.error = RETURN_DENSITY_ERROR;
                        return PHASE_DONE;
                }
        }
        const int scheme = cfg.energy_loss;
        lp.ti = state.time;
        lp.wi = state.weight;
        lp.Xi = state.grammage;
        if (cfg.energy_loss >= MODE_CSDA) {
                const Loc li = locate_near(p, p.K, state.energy, trk.iK);
                lp.Xf = cel_grammage_at(p, scheme, lp.material, li);
                lp.dei = 1. / cel_energy_loss_at(p, scheme, lp.material, li);
        }
        transport_limit(p, cfg, trk, rng, state, lp.material, lp.Xi, lp.Xf,
            &lp.grammage_max);
        if (trk.step_event) return stepping_stop(cfg, trk, state, lp, rec);
        return PHASE_STEP;
}

/* One iteration of the stepping loop, pumas.c:5277-5335 and, for a change of
 * medium, 5417-5467 + re-arm. */
template <class Cfg, class Geo, class Rng, class Sync = NoSync,
    class Rec = NoRecorder>
PB_FN int stepping_step(const Physics & p, const Cfg & cfg,
    const Geo & geo, Track & trk, Rng & rng, State & state, Loop & lp,
    const Sync & sync = Sync(), const Rec & rec = Rec())
{
        if (!lp.first) {
                lp.step_max_type = medium_locate(geo, cfg.direction, state,
                    false, 0, true, &lp.step_max_medium);
                if ((lp.step_max_medium > 0.) &&
                    (lp.step_max_type == STEP_CHECK))
                        lp.step_max_medium += 0.5 * STEP_MIN;
        }

        int new_medium = -1;
        if (!step_transport(p, cfg, geo, trk, rng, state, lp.straight,
                lp.medium, lp.locals, lp.grammage_max, lp.step_max_medium,
                lp.step_max_type, &lp.step_max_locals, &new_medium, sync))
                return PHASE_DONE; /* error: no final protection (:5293) */
        lp.first = 0;

        /* sampled steps, pumas.c:5296-5304 (step_index = steps so far + 1) */
        bool record_step = false;
        if (Rec::enabled && (rec.period() > 0))
                record_step =
                    ((trk.n_steps + 1u) % (unsigned int)rec.period()) == 0u;
        if (!trk.step_event && !record_step) return PHASE_STEP;

        /* Weight update at a boundary or hard loss, pumas.c:5309-5321. */
        const int scheme = cfg.energy_loss;
        if ((cfg.direction == MODE_BACKWARD) &&
            (cfg.energy_loss >= MODE_CSDA)) {
                const double wt = (cfg.decay == MODE_WEIGHTED) ?
                    xexp(-fabs(state.time - lp.ti) / p.ctau) :
                    1.;
                state.weight = lp.wi * wt * lp.dei *
                    cel_energy_loss_at(p, scheme, lp.material,
                        locate_near(p, p.K, state.energy, trk.iK));
        } else if (cfg.decay == MODE_WEIGHTED) {
                state.weight = lp.wi * xexp(-fabs(state.time - lp.ti) / p.ctau);
        }

        if (!trk.step_event) {
                /* a sampled step, pumas.c:5324-5327 */
                if (Rec::enabled) rec.frame(lp.medium, trk.step_event, state);
                return PHASE_STEP;
        }
        if (trk.step_event &
            (EVENT_LIMIT | EVENT_WEIGHT | EVENT_VERTEX_DECAY)) {
                return stepping_stop(cfg, trk, state, lp, rec);
        } else if (trk.step_event & EVENT_VERTEX_DEL) {
                return PHASE_DEL;
        } else if (trk.step_event & EVENT_VERTEX_COULOMB) {
                return PHASE_EHS;
        }

        /* EVENT_MEDIUM */
        if (!geo_has_boundaries(geo)) {
                stepping_finish(cfg, state); /* not reached */
                return PHASE_DONE;
        }
        lp.medium = new_medium;
        if ((lp.medium < 0) || (cfg.event & EVENT_MEDIUM))
                return stepping_stop(cfg, trk, state, lp, rec);
        lp.material = geo_material(geo, lp.medium);
        lp.locals.density = 0.;
        lp.locals.magnet[0] = lp.locals.magnet[1] = lp.locals.magnet[2] = 0.;
        lp.locals.magnetized = 0;
        lp.step_max_locals =
            set_locals(p, geo, cfg.accuracy, lp.medium, state, &lp.locals);
        if (lp.locals.density <= 0.) {
                trk.error = RETURN_DENSITY_ERROR;
                return PHASE_DONE;
        }
        lp.straight = straight_flag(cfg, lp.step_max_locals, lp.locals);

        const int tmp_scheme = scheme > MODE_DISABLED ? scheme : MODE_CSDA;
        trk.step_X_limit = (cfg.event & EVENT_LIMIT_ENERGY) ?
            cel_grammage(p, tmp_scheme, lp.material, cfg.limit_energy) :
            0.;
        trk.step_first = 1;
        trk.step_invlb1 = 0;
        /* the change of medium, pumas.c:5464-5467 */
        if (Rec::enabled && !(cfg.event & EVENT_MEDIUM))
                rec.frame(lp.medium, trk.step_event, state);
        return stepping_rearm(p, cfg, geo, trk, rng, state, lp, false, rec);
}

/* A hard inelastic event, pumas.c:5340-5392 then re-arm. */
template <class Cfg, class Geo, class Rng, class Rec = NoRecorder>
PB_FN int stepping_del(const Physics & p, const Cfg & cfg,
    const Geo & geo, Track & trk, Rng & rng, State & state, Loop & lp,
    const Rec & rec = Rec())
{
        /* the pre-vertex point, pumas.c:5343-5375 */
        const bool record = Rec::enabled && (rec.period() > 0);
        const double ki = state.energy;
        const double ui[3] = { state.direction[0], state.direction[1],
                state.direction[2] };
        transport_do_del(p, cfg, trk, rng, state, lp.material);
        if (record) {
                State pre = state;
                pre.energy = ki;
                pre.direction[0] = ui[0];
                pre.direction[1] = ui[1];
                pre.direction[2] = ui[2];
                rec.frame(lp.medium, trk.step_event, pre);
        }
        if ((cfg.event & trk.step_event) || trk.step_event == EVENT_WEIGHT)
                return stepping_stop(cfg, trk, state, lp, rec);
        /* the post-vertex point, pumas.c:5383-5386 */
        if (record) rec.frame(lp.medium, EVENT_NONE, state);
        trk.step_first = 1;
        trk.step_invlb1 = 0;
        return stepping_rearm(p, cfg, geo, trk, rng, state, lp, true, rec);
}

/* A hard elastic event, pumas.c:5393-5402 then re-arm. */
template <class Cfg, class Geo, class Rng, class Rec = NoRecorder>
PB_FN int stepping_ehs(const Physics & p, const Cfg & cfg,
    const Geo & geo, Track & trk, Rng & rng, State & state, Loop & lp,
    const Rec & rec = Rec())
{
        transport_do_ehs(p, cfg, rng, state, lp.material, trk.iK);
        if (cfg.event & trk.step_event)
                return stepping_stop(cfg, trk, state, lp, rec);
        return stepping_rearm(p, cfg, geo, trk, rng, state, lp, true, rec);
}

/* Total magnetic transport within the CSDA, uniform medium: the Larmor
 * polynomials of the phase on the tabulated moments Li. pumas.c:5095-5168.
 * Returns false when the rotation exceeds the range of the polynomials
 * (PUMAS_RETURN_VALUE_ERROR, pumas.c:5133-5136). */
PB_NOINLINE bool csda_magnetic_transport(const Physics & p, int material,
    double density, double magnet, double charge, double kinetic, double phase,
    double * x, double * y, double * z)
{
        const int imax = p.n_energies - 1;
        double k0 = kinetic;
        int i1, i2;
        if (kinetic <= p.K[0]) {
                k0 = p.K[0];
                i1 = 0;
                i2 = 1;
        } else if (kinetic >= p.K[imax]) {
                /* deflection neglected at very high energy */
                *y = 0.0;
                *x = *z = -(cel_grammage(p, MODE_CSDA, material, kinetic) -
                    col_sm(p, p.X, MODE_CSDA, material)[imax]);
                return true;
        } else {
                i1 = table_index(p, p.K, k0);
                i2 = i1 + 1;
        }

        const double max_phi = 2.0 * PI;
        const double poly_x[N_LARMOR_ORDERS] = { 1.000000000e+000,
                0.000000000e+000, -5.000000000e-001, -1.048969541e-002,
                5.597396438e-002, -6.401353612e-003, -4.495849930e-004,
                6.583532135e-005 };
        const double poly_y[N_LARMOR_ORDERS] = { 0.000000000e+000,
                1.000000000e+000, -1.400755143e-002, -1.350173383e-001,
                -2.838778336e-002, 2.123237056e-002, -3.094290091e-003,
                1.409012754e-004 };
        if (fabs(phase) > max_phi) return false;

        /* table_get_Li, pumas.c:4594-4599 */
        const double * const Li =
            p.Li + (size_t)material * N_LARMOR_ORDERS * p.n_energies;
        double x1, x2, y1, y2, z1, z2;
        {
                const double a = -magnet * charge / density;
                x1 = x2 = y1 = y2 = 0.;
                double u = 1.;
                for (int j = 0; j < N_LARMOR_ORDERS; j++) {
                        const double l1 = Li[j * p.n_energies + i1];
                        const double l2 = Li[j * p.n_energies + i2];
                        x1 += u * poly_x[j] * l1;
                        y1 += u * poly_y[j] * l1;
                        x2 += u * poly_x[j] * l2;
                        y2 += u * poly_y[j] * l2;
                        u *= a;
                }
                z1 = Li[i1];
                z2 = Li[i2];
        }
        {
                const double h = (k0 - p.K[i1]) / (p.K[i2] - p.K[i1]);
                *x = (x1 + h * (x2 - x1)) / density;
                *y = (y1 + h * (y2 - y1)) / density;
                *z = (z1 + h * (z2 - z1)) / density;
        }
        return true;
}

/* Magnetic deflection of the closed form CSDA in a uniform field.
 * pumas.c:4968-5075. `state.energy` is the final kinetic energy. Returns a
 * pumas_return code. */
PB_NOINLINE int transport_csda_deflect(const Physics & p, int direction_mode,
    State & state, int material, const Locals & locals, double ki,
    double distance)
{
        (void)direction_mode;
        const double charge = state.charge;
        const double kf = state.energy;
        double * const position = state.position;
        double * const direction = state.direction;
        const double density = locals.density;
        const double * const magnet = locals.magnet;

        double b0, ex[3], ey[3], ez[3];
        {
                b0 = sqrt(magnet[0] * magnet[0] + magnet[1] * magnet[1] +
                    magnet[2] * magnet[2]);
                const double b0_i = 1.0 / b0;
                ey[0] = (direction[2] * magnet[1] - direction[1] * magnet[2]) *
                    b0_i;
                ey[1] = (direction[0] * magnet[2] - direction[2] * magnet[0]) *
                    b0_i;
                ey[2] = (direction[1] * magnet[0] - direction[0] * magnet[1]) *
                    b0_i;
                {
                        const double d = b0_i * b0_i *
                            (magnet[0] * direction[0] +
                                magnet[1] * direction[1] +
                                magnet[2] * direction[2]);
                        ez[0] = magnet[0] * d;
                        ez[1] = magnet[1] * d;
                        ez[2] = magnet[2] * d;
                }
                ex[0] = direction[0] - ez[0];
                ex[1] = direction[1] - ez[1];
                ex[2] = direction[2] - ez[2];
        }

        double dx, dy, dz, dp;
        {
                const double a = -b0 * charge / density;
                const double pi = a * cel_magnetic_rotation(p, material, ki);
                const double pf = a * cel_magnetic_rotation(p, material, kf);
                dp = pf - pi;

                const double ps = fabs(pi + pf);
                if (ps == 0.) return RETURN_SUCCESS;
                const double dp0 = fabs(dp);
                if ((dp0 < 1E-02) && (dp0 / ps < 1E-02)) {
                        if (ki >= kf)
                                dz = distance;
                        else
                                dz = -distance;
                        const double sixth = 1 / 6.0;
                        dx = dz * (1.0 - sixth * dp * dp);
                        dy = 0.5 * dz * dp;
                } else {
                        double xi, yi, zi, xf, yf, zf;
                        if (!csda_magnetic_transport(p, material, density, b0,
                                charge, ki, pi, &xi, &yi, &zi))
                                return RETURN_VALUE_ERROR;
                        if (!csda_magnetic_transport(p, material, density, b0,
                                charge, kf, pf, &xf, &yf, &zf))
                                return RETURN_VALUE_ERROR;
                        {
                                const double si = xsin(pi);
                                const double ci = xcos(pi);
                                xf -= xi;
                                yf -= yi;
                                dx = ci * xf + si * yf;
                                dy = -si * xf + ci * yf;
                        }
                        dz = zf - zi;
                }
        }
        position[0] += dx * ex[0] + dy * ey[0] + dz * ez[0];
        position[1] += dx * ex[1] + dy * ey[1] + dz * ez[1];
        position[2] += dx * ex[2] + dy * ey[2] + dz * ez[2];

        const double s = xsin(dp);
        const double c = xcos(dp);
        direction[0] = c * ex[0] + s * ey[0] + ez[0];
        direction[1] = c * ex[1] + s * ey[1] + ez[1];
        direction[2] = c * ex[2] + s * ey[2] + ez[2];
        return RETURN_SUCCESS;
}

/* Closed form CSDA in a uniform medium. pumas.c:4777-4949. `*rc` receives the
 * return code of the magnetic deflection (pumas.c:4929-4933). */
template <class Cfg, class Geo, class Rec = NoRecorder>
PB_FN int transport_with_csda(const Physics & p, const Cfg & cfg,
    const Geo & geo, const Track & trk, State & state, int medium,
    const Locals & locals, int * rc, const Rec & rec = Rec())
{
        const int material = geo_material(geo, medium);
        const double density = locals.density;
        const double ki = state.energy;
        const double di = state.distance;
        const double ti = state.time;
        const double xi = cel_grammage(p, MODE_CSDA, material, ki);
        const double Ti = cel_proper_time(p, MODE_CSDA, material, ki);

        int event = EVENT_NONE;
        /* the start of the track, pumas.c:4797-4798 */
        if (Rec::enabled) rec.frame(medium, event | EVENT_START, state);
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
        int decayed = 0;
        double time_max = (cfg.event & EVENT_LIMIT_TIME) ? cfg.limit_time : 0.;
        if (cfg.decay == MODE_RANDOMISED) {
                if ((time_max <= 0.) || (trk.lifetime < time_max)) {
                        time_max = trk.lifetime;
                        decayed = 1;
                }
        }
        if (time_max > 0.) {
                const double dt = time_max - ti;
                if (dt <= 0.) {
                        xB = 0.;
                        event = EVENT_LIMIT_TIME;
                } else {
                        const double sgn =
                            (cfg.direction == MODE_FORWARD) ? 1. : -1.;
                        const double Tf = Ti - sgn * dt * density;
                        if (Tf > 0.) {
                                const double xT = fabs(xi -
                                    cel_grammage_as_time(
                                        p, MODE_CSDA, material, Tf));
                                if (xT < xB) {
                                        xB = xT;
                                        event = EVENT_LIMIT_TIME;
                                }
                        }
                }
        }
        if (xB <= 0.) return event;

        double xf, kf;
        if (cfg.direction == MODE_FORWARD) {
                if (xi > xB) {
                        xf = xi - xB;
                        if (xB < EPSILON_X * xi) {
                                kf = ki -
                                    cel_energy_loss(p, MODE_CSDA, material, ki) *
                                        xB;
                                if (kf < 0.) kf = 0.;
                        } else {
                                kf = cel_kinetic_energy(
                                    p, MODE_CSDA, material, xf);
                        }
                } else {
                        xf = kf = 0.;
                        event = EVENT_LIMIT_ENERGY;
                }

                if (cfg.event & EVENT_LIMIT_ENERGY) {
                        if (ki <= cfg.limit_energy) return EVENT_LIMIT_ENERGY;
                        if (kf < cfg.limit_energy) {
                                kf = cfg.limit_energy;
                                xf = cel_grammage(p, MODE_CSDA, material, kf);
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
                                kf = ki +
                                    cel_energy_loss(p, MODE_CSDA, material, ki) *
                                        xB;
                        } else {
                                kf = cel_kinetic_energy(
                                    p, MODE_CSDA, material, xf);
                        }
                }
                if (cfg.event & EVENT_LIMIT_ENERGY) {
                        if (ki >= cfg.limit_energy) return EVENT_LIMIT_ENERGY;
                        if (kf > cfg.limit_energy) {
                                kf = cfg.limit_energy;
                                xf = cel_grammage(p, MODE_CSDA, material, kf);
                                event = EVENT_LIMIT_ENERGY;
                        }
                }
        }

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
        if (event == EVENT_LIMIT_TIME) {
                state.time = time_max;
                state.decayed = decayed;
                if (decayed) event = EVENT_VERTEX_DECAY;
        } else
                state.time +=
                    fabs(Ti - cel_proper_time(p, MODE_CSDA, material, kf)) /
                    density;
        if (cfg.direction == MODE_BACKWARD)
                state.weight *= cel_energy_loss(p, MODE_CSDA, material, kf) /
                    cel_energy_loss(p, MODE_CSDA, material, ki);
        if (cfg.decay == MODE_WEIGHTED)
                state.weight *= xexp(-fabs(ti - state.time) / p.ctau);

        if (geo_has_locals(geo) && (locals.magnetized != 0)) {
                *rc = transport_csda_deflect(
                    p, cfg.direction, state, material, locals, ki, distance);
        } else {
                double path;
                if (cfg.direction == MODE_FORWARD)
                        path = distance;
                else
                        path = -distance;
                state.position[0] += path * state.direction[0];
                state.position[1] += path * state.direction[1];
                state.position[2] += path * state.direction[2];
        }

        /* the end of the track, pumas.c:4945-4946 */
        if (Rec::enabled) rec.frame(medium, event | EVENT_STOP, state);
        return event;
}

/* Result of one track: the reference's return code, event and media[2]. */
struct Outcome {
        int rc;
        int event;
        int medium[2];
};

/* Batch level validation, pumas.c:3152-3177 (done once on the host). */
template <class Cfg>
PB_FN int settings_check(const Physics & p, const Cfg & cfg)
{
        if ((p.particle == 1) && (cfg.direction == MODE_FORWARD) &&
            (cfg.decay == MODE_WEIGHTED))
                return 3; /* PUMAS_RETURN_DECAY_ERROR */
        if ((cfg.accuracy <= 0) || (cfg.accuracy > 1))
                return 1; /* PUMAS_RETURN_ACCURACY_ERROR */
        if ((cfg.direction == MODE_BACKWARD) &&
            (cfg.energy_loss > MODE_CSDA) && (p.cutoff < 1E-02))
                return 1;
        return 0;
}

/* Start of pumas_context_transport for one track, pumas.c:3114-3235: checks,
 * first medium and locals, lifetime draw, then either the closed form CSDA
 * (the track is then over) or the prologue of the stepping. Returns the phase
 * of the track; `out` holds rc, the first medium and, when PHASE_DONE is
 * returned, the end event and last medium. */
template <class Cfg, class Geo, class Rng, class Rec = NoRecorder>
PB_FN int track_begin(const Physics & p, const Cfg & cfg,
    const Geo & geo, Rng & rng, State & state, Track & trk, Loop & lp,
    Outcome & out, const Rec & rec = Rec())
{
        out.rc = RETURN_SUCCESS;
        out.event = EVENT_NONE;
        out.medium[0] = out.medium[1] = -1;

        trk.lifetime = 0.;
        trk.step_first = 1;
        trk.step_event = EVENT_NONE;
        trk.step_foreseen = EVENT_NONE;
        trk.step_X_limit = 0.;
        trk.step_invlb1 = 0.;
        trk.randn_done = 0;
        trk.randn_next = 0.;
        trk.error = RETURN_SUCCESS;
        trk.n_steps = 0;
        trk.iK = p.n_energies / 2;

        if (state.decayed) {
                out.event = EVENT_VERTEX_DECAY;
                return PHASE_DONE;
        }

        {
                const double * const u = state.direction;
                const double norm2 = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
                if (fabs(norm2 - 1) > (double)FLT_EPSILON) {
                        out.rc = 5; /* PUMAS_RETURN_DIRECTION_ERROR */
                        return PHASE_DONE;
                }
        }

        int medium = -1;
        double step_max_medium = 0.;
        const int step_max_type = medium_locate(
            geo, cfg.direction, state, true, &medium, true, &step_max_medium);
        out.medium[0] = medium;
        out.medium[1] = -1;
        if (medium < 0) {
                out.event = EVENT_MEDIUM;
                /* pumas.c:3190-3194 */
                if (Rec::enabled)
                        rec.frame(-1,
                            EVENT_MEDIUM | EVENT_START | EVENT_STOP, state);
                return PHASE_DONE;
        } else if ((step_max_medium > 0.) && (step_max_type == STEP_CHECK))
                step_max_medium += 0.5 * STEP_MIN;
        Locals & locals = lp.locals;
        locals.density = 0.;
        locals.magnet[0] = locals.magnet[1] = locals.magnet[2] = 0.;
        locals.magnetized = 0;
        const double step_max_locals =
            set_locals(p, geo, cfg.accuracy, medium, state, &locals);
        if ((step_max_locals > 0.) && (step_max_locals < step_max_medium))
                step_max_medium = step_max_locals;
        if (locals.density <= 0.) {
                out.rc = RETURN_DENSITY_ERROR;
                return PHASE_DONE;
        }

        if (cfg.decay == MODE_RANDOMISED) {
                const double u = rng.uniform();
                trk.lifetime = state.time - p.ctau * xlog(u);
        }

        int do_stepping = 1;
        if ((step_max_medium <= 0.) && (step_max_locals <= 0.) &&
            (cfg.energy_loss <= MODE_CSDA)) {
                if ((cfg.energy_loss == MODE_DISABLED) &&
                    ((cfg.event & EVENT_LIMIT) == 0)) {
                        out.rc = 15; /* PUMAS_RETURN_MISSING_LIMIT */
                        return PHASE_DONE;
                } else if ((cfg.scattering == MODE_DISABLED) &&
                    (cfg.energy_loss == MODE_CSDA)) {
                        do_stepping = 0;
                }
        }

        lp.medium = medium;
        if (!do_stepping) {
                out.event = transport_with_csda(
                    p, cfg, geo, trk, state, medium, locals, &out.rc, rec);
                out.medium[1] = medium;
                return PHASE_DONE;
        }

        lp.step_max_medium = step_max_medium;
        lp.step_max_type = step_max_type;
        lp.step_max_locals = step_max_locals;
        const int phase =
            stepping_begin(p, cfg, geo, trk, rng, state, lp, rec);
        if (phase == PHASE_DONE) {
                out.event = trk.step_event;
                out.medium[1] = lp.medium;
                out.rc = trk.error;
        }
        return phase;
}

/* End of a track that went through the stepping, pumas.c:3248-3250. */
PB_FN void track_end(const Track & trk, const Loop & lp, Outcome & out)
{
        out.event = trk.step_event;
        out.medium[1] = lp.medium;
        out.rc = trk.error;
}

/* pumas_context_transport for one track run to completion (CPU build and the
 * reference kernel): the phases chained in sequence. */
template <class Cfg, class Geo, class Rng, class Rec = NoRecorder>
PB_FN Outcome transport_track(const Physics & p, const Cfg & cfg,
    const Geo & geo, Rng & rng, State & state, unsigned int * n_steps,
    const Rec & rec = Rec())
{
        Outcome out;
        Track trk;
        Loop lp;
        int phase = track_begin(p, cfg, geo, rng, state, trk, lp, out, rec);
        if (phase != PHASE_DONE) {
                while (phase != PHASE_DONE) {
                        if (phase == PHASE_STEP)
                                phase = stepping_step(p, cfg, geo, trk, rng,
                                    state, lp, NoSync(), rec);
                        else if (phase == PHASE_EHS)
                                phase = stepping_ehs(
                                    p, cfg, geo, trk, rng, state, lp, rec);
                        else
                                phase = stepping_del(
                                    p, cfg, geo, trk, rng, state, lp, rec);
                }
                track_end(trk, lp, out);
        }
        if (n_steps) *n_steps = trk.n_steps;
        return out;
}

} /* namespace pb */

#endif
