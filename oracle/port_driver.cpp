/*
 * port_driver.cpp -- CPU debugging build of the single-source transport core.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/oracle_driver.c for the rules). This
 * compiles the very headers the CUDA kernels are made of (pumas_b200/csrc/
 * pb_*.h) with g++ so that the restated algorithm can be checked against the
 * real reference (oracle/_ref) on a machine without a GPU, track by track and
 * function by function. It is NOT independent of the product (same source), so
 * it is not evidence of parity by itself: the evidence is `port == reference`
 * here (tests/test_port_vs_reference.py) and `GPU == reference` on the GPU box
 * (tests/test_gpu_parity.py).
 * Each restated function cites the reference lines it follows, in the headers.
 */
#include <pthread.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "pb_dump.hpp"
#include "pb_csda.h"
#include "pb_hot.h"
#include "pb_transport.h"
#include "philox.h"
#include "pumas_b200.h"

namespace {

struct PhiloxRng {
        pb_philox g;
        double uniform() { return pb_philox_uniform(&g); }
        /* (the interface of the lean step, pb_hot.h) */
        double uniform_fast() { return uniform(); }
        double uniform_keep() { return uniform(); }
        double log_uniform_fast() { return pb::xlog(uniform()); }
        double log_uniform_keep() { return pb::xlog(uniform()); }
};

pb::Settings settings_from(const struct pumas_context * c)
{
        pb::Settings s;
        s.energy_loss = (int)c->mode.energy_loss;
        s.decay = (int)c->mode.decay;
        s.direction = (int)c->mode.direction;
        s.scattering = (int)c->mode.scattering;
        s.event = (int)c->event;
        s.limit_energy = c->limit.energy;
        s.limit_distance = c->limit.distance;
        s.limit_grammage = c->limit.grammage;
        s.limit_time = c->limit.time;
        s.accuracy = c->accuracy;
        return s;
}

struct Job {
        const pb::HostPhysics * hp;
        pb::Settings cfg;
        const pumas_b200_geometry * geo;
        uint64_t seed, offset;
        long begin, end;
        pb::State * states;
        int * events;
        int * media;
        unsigned int * n_random;
        unsigned int * n_steps;
        int rc;
};

/* 1: use the compile time specialised instantiations the GPU library uses
 * (pumas_b200/csrc/kernels_spec*.cu) when the configuration matches; 2: the
 * same through the lean step of pb_hot.h (what the GPU runs for them); 0: the
 * general instantiation only. */
int g_specialise = 1;

template <class Cfg, class Geo>
void run_tracks(Job * job, const Cfg & cfg, const Geo & geo)
{
        const pb::Physics & p = job->hp->view;
        for (long i = job->begin; i < job->end; i++) {
                PhiloxRng rng;
                pb_philox_init(&rng.g, job->seed, job->offset + (uint64_t)i);
                unsigned int steps = 0;
                const pb::Outcome out = pb::transport_track(
                    p, cfg, geo, rng, job->states[i], &steps);
                if ((out.rc != 0) && (job->rc == 0)) job->rc = out.rc;
                if (job->events) job->events[i] = out.event;
                if (job->media) {
                        job->media[2 * i] = out.medium[0];
                        job->media[2 * i + 1] = out.medium[1];
                }
                if (job->n_random) job->n_random[i] = rng.g.count;
                if (job->n_steps) job->n_steps[i] = steps;
        }
}

/* The lean step of the GPU kernels (pb_hot.h) chained with the general event
 * handlers, through the slot packing the kernel uses. */
template <class Cfg>
void run_tracks_hot(Job * job, const Cfg & cfg, const pb::UniformGeometry & geo)
{
        using namespace pb;
        const Physics & p = job->hp->view;
        /* (the packed tables the dump loader built) */
        const HotCtx h = hot_context(
            p, p, p.hotK, p.hotX, cfg, geo.material, geo.density);
        for (long i = job->begin; i < job->end; i++) {
                PhiloxRng rng;
                pb_philox_init(&rng.g, job->seed, job->offset + (uint64_t)i);
                State & state = job->states[i];
                const double charge = state.charge;
                Outcome out;
                Track trk;
                Loop lp;
                int phase = track_begin(p, cfg, geo, rng, state, trk, lp, out);
                if (phase != PHASE_DONE) {
                        PlainSlot s;
                        hot_pack(s, state, trk, lp, phase, false, 0.);
                        while (phase != PHASE_DONE) {
                                if (phase == PHASE_STEP) {
                                        phase = hot_step(p, cfg, geo, h, s, rng);
                                        continue;
                                }
                                if (phase == PHASE_EHS) {
                                        phase = hot_ehs(p, cfg, geo, h, s, rng);
                                        continue;
                                }
                                hot_unpack(s, p, cfg, geo, state, trk, lp);
                                trk.step_X_limit = h.step_X_limit;
                                phase = stepping_del(
                                    p, cfg, geo, trk, rng, state, lp);
                                hot_pack(s, state, trk, lp, phase, false, 0.);
                        }
                        hot_unpack(s, p, cfg, geo, state, trk, lp);
                        state.charge = charge;
                        track_end(trk, lp, out);
                }
                if ((out.rc != 0) && (job->rc == 0)) job->rc = out.rc;
                if (job->events) job->events[i] = out.event;
                if (job->media) {
                        job->media[2 * i] = out.medium[0];
                        job->media[2 * i + 1] = out.medium[1];
                }
                if (job->n_random) job->n_random[i] = rng.g.count;
                if (job->n_steps) job->n_steps[i] = trk.n_steps;
        }
}

/* The lean closed form CSDA of the GPU kernel (pb_csda.h), with the general
 * transport_track behind it. */
template <class Cfg>
void run_tracks_csda(Job * job, const Cfg & cfg, const pb::UniformGeometry & geo)
{
        using namespace pb;
        const Physics & p = job->hp->view;
        const CsdaCtx c = csda_context(p, p, cfg, geo.material, geo.density);
        for (long i = job->begin; i < job->end; i++) {
                PhiloxRng rng;
                pb_philox_init(&rng.g, job->seed, job->offset + (uint64_t)i);
                unsigned int steps = 0;
                Outcome out;
                if (!csda_track(cfg, c, job->states[i], out))
                        out = transport_track(
                            p, cfg, geo, rng, job->states[i], &steps);
                if ((out.rc != 0) && (job->rc == 0)) job->rc = out.rc;
                if (job->events) job->events[i] = out.event;
                if (job->media) {
                        job->media[2 * i] = out.medium[0];
                        job->media[2 * i + 1] = out.medium[1];
                }
                if (job->n_random) job->n_random[i] = rng.g.count;
                if (job->n_steps) job->n_steps[i] = steps;
        }
}

template <int DI> bool try_csda(Job * job)
{
        typedef pb::StaticSettings<pb::MODE_CSDA, pb::MODE_DISABLED, DI,
            pb::MODE_WEIGHTED>
            Cfg;
        const pumas_b200_geometry & geo = *job->geo;
        pb::UniformGeometry ug;
        if (!Cfg::matches(job->cfg) || !pb::uniform_geometry_of(geo, &ug))
                return false;
        run_tracks_csda(job, Cfg(job->cfg), ug);
        return true;
}

template <int EL, int SC, int DI, int DE> bool try_specialised(Job * job)
{
        typedef pb::StaticSettings<EL, SC, DI, DE> Cfg;
        const pumas_b200_geometry & geo = *job->geo;
        pb::UniformGeometry ug;
        if (!Cfg::matches(job->cfg) || !pb::uniform_geometry_of(geo, &ug))
                return false;
        if (g_specialise == 2)
                run_tracks_hot(job, Cfg(job->cfg), ug);
        else
                run_tracks(job, Cfg(job->cfg), ug);
        return true;
}

void * worker(void * arg)
{
        Job * job = static_cast<Job *>(arg);
        using namespace pb;
        if ((g_specialise == 2) &&
            (try_csda<MODE_FORWARD>(job) || try_csda<MODE_BACKWARD>(job)))
                return NULL;
        if (g_specialise &&
            (try_specialised<MODE_STRAGGLED, MODE_MIXED, MODE_FORWARD,
                 MODE_WEIGHTED>(job) ||
                try_specialised<MODE_STRAGGLED, MODE_MIXED, MODE_BACKWARD,
                    MODE_WEIGHTED>(job) ||
                try_specialised<MODE_STRAGGLED, MODE_DISABLED, MODE_FORWARD,
                    MODE_WEIGHTED>(job) ||
                try_specialised<MODE_STRAGGLED, MODE_DISABLED, MODE_BACKWARD,
                    MODE_WEIGHTED>(job)))
                return NULL;
        run_tracks(job, job->cfg, *job->geo);
        return NULL;
}

} /* namespace */

extern "C" {

void port_set_specialise(int enable) { g_specialise = enable; }

void * port_physics_load(const char * path)
{
        FILE * f = fopen(path, "rb");
        if (f == NULL) return NULL;
        pb::HostPhysics * hp = new pb::HostPhysics;
        std::string err;
        const int rc = pb::load_physics_stream(f, *hp, err);
        fclose(f);
        if (rc != 0) {
                fprintf(stderr, "port_physics_load: %s\n", err.c_str());
                delete hp;
                return NULL;
        }
        return hp;
}

void port_physics_destroy(void * physics)
{
        delete static_cast<pb::HostPhysics *>(physics);
}

int port_transport_n(const void * physics, const struct pumas_context * settings,
    const struct pumas_b200_geometry * geometry, uint64_t seed, uint64_t offset,
    long n, struct pumas_state * states, int * events, int * media,
    unsigned int * n_random, unsigned int * n_steps, int n_threads)
{
        const pb::HostPhysics * hp = static_cast<const pb::HostPhysics *>(physics);
        const pb::Settings cfg = settings_from(settings);
        const int rc0 = pb::settings_check(hp->view, cfg);
        if (rc0 != 0) return rc0;
        if (n_threads < 1) n_threads = 1;
        if ((long)n_threads > n) n_threads = (n > 0) ? (int)n : 1;
        std::vector<Job> jobs(n_threads);
        std::vector<pthread_t> threads(n_threads);
        for (int t = 0; t < n_threads; t++) {
                Job & j = jobs[t];
                j.hp = hp;
                j.cfg = cfg;
                j.geo = geometry;
                j.seed = seed;
                j.offset = offset;
                j.begin = n * t / n_threads;
                j.end = n * (t + 1) / n_threads;
                j.states = reinterpret_cast<pb::State *>(states);
                j.events = events;
                j.media = media;
                j.n_random = n_random;
                j.n_steps = n_steps;
                j.rc = 0;
                if (n_threads == 1)
                        worker(&j);
                else
                        pthread_create(&threads[t], NULL, &worker, &j);
        }
        int rc = 0;
        for (int t = 0; t < n_threads; t++) {
                if (n_threads > 1) pthread_join(threads[t], NULL);
                if ((rc == 0) && (jobs[t].rc != 0)) rc = jobs[t].rc;
        }
        return rc;
}

/* ---- the callback interface on the CPU build --------------------------------
 * CallbackGeometry (pb_media.h) with HOST functions: the user media of
 * examples/user_media.h, the same header the device build and the reference's
 * host callbacks (oracle_driver.c) compile. tests/test_port_vs_reference.py
 * compares this path with the reference bit for bit. */
#include "../examples/user_media.h"

static int port_user_medium_fn(const void * user_data,
    const struct pumas_state * state, int backward, int * medium, double * step)
{
        return user_media_locate(static_cast<const struct user_media *>(user_data),
            state->position, state->direction, backward, medium, step);
}

static double port_user_locals_fn(const void * user_data, int medium,
    const struct pumas_state * state, double * density, double * magnet)
{
        magnet[0] = magnet[1] = magnet[2] = 0.;
        return user_media_density(
            static_cast<const struct user_media *>(user_data), medium,
            state->position, density);
}

int port_transport_user_media(const void * physics,
    const struct pumas_context * settings, const struct user_media * shells,
    const int * materials, uint64_t seed, uint64_t offset, long n,
    struct pumas_state * states, int * events, int * media,
    unsigned int * n_random, unsigned int * n_steps)
{
        const pb::HostPhysics * hp = static_cast<const pb::HostPhysics *>(physics);
        const pb::Settings cfg = settings_from(settings);
        const int rc0 = pb::settings_check(hp->view, cfg);
        if (rc0 != 0) return rc0;
        pb::CallbackGeometry geo;
        memset(&geo, 0x0, sizeof geo);
        geo.cb.medium = &port_user_medium_fn;
        geo.cb.locals = &port_user_locals_fn;
        geo.cb.user_data = shells;
        geo.cb.n_media = shells->n_shells;
        for (int i = 0; i < shells->n_shells; i++) {
                geo.cb.material[i] = materials[i];
                geo.cb.has_locals[i] = 1;
        }
        Job job;
        job.hp = hp;
        job.cfg = cfg;
        job.geo = NULL;
        job.seed = seed;
        job.offset = offset;
        job.begin = 0;
        job.end = n;
        job.states = reinterpret_cast<pb::State *>(states);
        job.events = events;
        job.media = media;
        job.n_random = n_random;
        job.n_steps = n_steps;
        job.rc = 0;
        run_tracks(&job, cfg, geo);
        return job.rc;
}

/* ---- a user DCS model on the CPU build (examples/user_dcs.h) -------------- */
#include "../examples/user_dcs.h"

static double port_user_dcs(double Z, double A, double m, double K, double q)
{
        return user_dcs_bremsstrahlung(Z, A, m, K, q);
}

/* Resolve the user models of a loaded physics (what pumas_physics_load of the
 * library does against its registry). Returns the number resolved. */
int port_register_user_dcs(void * physics)
{
        pb::HostPhysics * hp = static_cast<pb::HostPhysics *>(physics);
        int n = 0;
        if ((hp->view.model[0] == pb::MODEL_USER) &&
            (hp->model[0] == USER_DCS_MODEL)) {
                hp->view.user_dcs[0] = &port_user_dcs;
                n++;
        }
        return n;
}

/* Function level hooks, mirroring oracle_driver.c. */
double port_dcs_evaluate(
    const void * physics, int process, int element, double K, double q)
{
        const pb::HostPhysics * hp = static_cast<const pb::HostPhysics *>(physics);
        return pb::dcs_evaluate(hp->view, process, element, K, q);
}

double port_dcs_model(
    const void * physics, int process, int element, double K, double q)
{
        const pb::HostPhysics * hp = static_cast<const pb::HostPhysics *>(physics);
        return pb::dcs_model(
            hp->view, process, hp->view.element[element], K, q);
}

double port_coulomb_ehs_length(const void * physics, int material, double K)
{
        const pb::HostPhysics * hp = static_cast<const pb::HostPhysics *>(physics);
        return pb::coulomb_ehs_length(hp->view, material, K);
}

void port_table_get_msc(const void * physics, int scheme, int material,
    double K, double * mu0, double * invlb1)
{
        const pb::HostPhysics * hp = static_cast<const pb::HostPhysics *>(physics);
        pb::table_get_msc(hp->view, scheme, material, K, mu0, invlb1);
}

int port_coulomb_cs_hard(const void * physics, int material, int scheme,
    double kinetic, double * cs_hard)
{
        const pb::HostPhysics * hp = static_cast<const pb::HostPhysics *>(physics);
        const pb::Physics & p = hp->view;
        double mu0 = 0., invlb1 = 0.;
        pb::table_get_msc(p, scheme, material, kinetic, &mu0, &invlb1);
        const int n = p.elements_in[material];
        for (int i = 0; i < n; i++)
                cs_hard[i] = pb::coulomb_cs_hard(
                    p, p.comp_offset[material] + i, kinetic, mu0);
        return n;
}

/* Philox stream hooks (tests/test_philox.py). */
void port_philox_block(const uint32_t * ctr, const uint32_t * key, uint32_t * out)
{
        pb_philox4x32_10(ctr, key, out);
}

void port_philox_uniforms(uint64_t seed, uint64_t track, int n, double * out)
{
        pb_philox g;
        pb_philox_init(&g, seed, track);
        for (int i = 0; i < n; i++) out[i] = pb_philox_uniform(&g);
}

/* The recorder hooks of the engine (pb_transport.h) with a host recorder:
 * same interface as oracle_transport_recorded (tests/test_recorder.py). */
namespace {
struct HostRecorder {
        static constexpr bool enabled = true;
        int period_;
        std::vector<pb::State> * states;
        std::vector<int> * media;
        std::vector<int> * events;
        int period() const { return period_; }
        void frame(int medium, int event, const pb::State & state) const
        {
                states->push_back(state);
                media->push_back(medium);
                events->push_back(event);
        }
};
} /* namespace */

int port_transport_recorded(const void * physics,
    const struct pumas_context * settings,
    const struct pumas_b200_geometry * geometry, uint64_t seed, uint64_t offset,
    long n, struct pumas_state * states, int * events, int period,
    long max_frames, struct pumas_state * frame_states, int * frame_media,
    int * frame_events, int * frame_counts)
{
        const pb::HostPhysics * hp = static_cast<const pb::HostPhysics *>(physics);
        const pb::Settings cfg = settings_from(settings);
        std::vector<pb::State> fs;
        std::vector<int> fm, fe;
        HostRecorder rec = { period, &fs, &fm, &fe };
        int rc = 0;
        long total = 0;
        for (long i = 0; i < n; i++) {
                PhiloxRng rng;
                pb_philox_init(&rng.g, seed, offset + (uint64_t)i);
                fs.clear();
                fm.clear();
                fe.clear();
                unsigned int steps = 0;
                pb::State * s = reinterpret_cast<pb::State *>(states) + i;
                const pb::Outcome out = pb::transport_track(
                    hp->view, cfg, *geometry, rng, *s, &steps, rec);
                if ((out.rc != 0) && (rc == 0)) rc = out.rc;
                if (events) events[i] = out.event;
                frame_counts[i] = (int)fs.size();
                for (size_t k = 0; (k < fs.size()) && (total < max_frames);
                     k++, total++) {
                        memcpy(&frame_states[total], &fs[k], sizeof fs[k]);
                        frame_media[total] = fm[k];
                        frame_events[total] = fe[k];
                }
        }
        return rc;
}

/* The device logarithm of a uniform variate (tests/test_hot_math.py). */
double port_hot_log_uniform(uint32_t w) { return pb::hot_log_uniform(w); }

void port_hot_log_uniform_n(const uint32_t * w, long n, double * out)
{
        for (long i = 0; i < n; i++) out[i] = pb::hot_log_uniform(w[i]);
}

/* Single precision libm restatements (tests/test_libm_float.py). */
float port_logf(float x) { return pb::logf_libm(x); }
float port_expf(float x) { return pb::expf_libm(x); }

/* Raw table access, same names as oracle_table(). */
const double * port_table(const void * physics, const char * name, long * length)
{
        const pb::HostPhysics * hp = static_cast<const pb::HostPhysics *>(physics);
        const pb::Physics & p = hp->view;
        const long ne = p.n_energies, nm = p.n_materials;
        const std::string s(name);
#define T(tag, ptr, len)                                                       \
        if (s == tag) {                                                        \
                *length = (len);                                               \
                return (ptr);                                                  \
        }
        T("K", p.K, ne)
        T("K_dX", p.K_dX, pb::N_SCHEMES * nm * ne)
        T("X", p.X, pb::N_SCHEMES * nm * ne)
        T("X_dK", p.X_dK, pb::N_SCHEMES * nm * ne)
        T("T", p.T, pb::N_SCHEMES * nm * ne)
        T("dE", p.dE, pb::N_SCHEMES * nm * ne)
        T("Omega", p.Omega, nm * ne)
        T("CS", p.CS, nm * ne)
        T("NI_in", p.NI_in, nm * ne)
        T("NI_el", p.NI_el, pb::N_SCHEMES * nm * ne)
        T("Mu0", p.Mu0, nm * ne)
        T("Lb", p.Lb, nm * ne)
        T("Ms1", p.Ms1, (pb::N_SCHEMES + 1) * nm * ne)
        T("Kt", p.Kt, nm)
        T("a_max", p.a_max, nm)
        T("b_max", p.b_max, pb::N_SCHEMES * nm)
#undef T
        *length = 0;
        return NULL;
}

} /* extern "C" */
