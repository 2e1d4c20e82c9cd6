/*
 * oracle_driver.c -- CPU oracle for the transport hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (libpumas_b200.so, the
 * pumas_b200 python package) may include, link or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * do, and only as the checker or as the timed CPU baseline.
 *
 * This translation unit `#include`s the UNMODIFIED reference implementation
 * (src/pumas.c, from where it lies under /root/reference; the path comes from
 * the Makefile) and adds, outside of it:
 *   - a `context->random` callback (pumas.h:551, documented extension point
 *     pumas.h:623-624) backed by the Philox4x32-10 stream the GPU uses;
 *   - host `pumas_medium_cb` / `pumas_locals_cb` callbacks interpreting the
 *     same POD `struct pumas_b200_geometry` the GPU interprets;
 *   - a batch driver looping `pumas_context_transport` (pumas.c:3114) over n
 *     states, one `pumas_context` per thread (pumas.h:545-549);
 *   - a reset, before every track, of the per-context Box-Muller cache
 *     (`randn_done`, `randn_next`, pumas.c:455-462, toggled at pumas.c:7020),
 *     which otherwise leaks from one track to the next and is not reachable
 *     through the API. This is why the TU includes pumas.c rather than linking
 *     against libpumas.so.
 * Parity status: this IS the reference, so parity is pinned by construction;
 * tests/test_oracle_golden.py additionally checks it against the known-answer
 * values of the reference's own test suite (tests/test-pumas.c:908-986 etc.).
 */
#ifndef PUMAS_SOURCE
#error "PUMAS_SOURCE must point to the reference src/pumas.c"
#endif
#include PUMAS_SOURCE

#include <pthread.h>
#include <stdint.h>
#include <time.h>

#include "pumas_b200_geometry.h"
#include "philox.h"

/* ---- random callback ---------------------------------------------------- */
struct oracle_user_data {
        struct pb_philox rng;
};

static double oracle_random(struct pumas_context * context)
{
        struct oracle_user_data * u = context->user_data;
        return pb_philox_uniform(&u->rng);
}

/* ---- media callbacks ---------------------------------------------------- */
/* Geometry of the current batch. Read-only while worker threads run. */
static struct pumas_b200_geometry oracle_geometry;

struct oracle_medium {
        struct pumas_medium api; /* must be first: sub-classing, pumas.h:360 */
        int layer;
};
static struct oracle_medium oracle_media[PUMAS_B200_MAX_LAYERS];

static double oracle_locals(struct pumas_medium * medium,
    struct pumas_state * state, struct pumas_locals * locals)
{
        const struct oracle_medium * m = (struct oracle_medium *)medium;
        const struct pumas_b200_layer * l = &oracle_geometry.layer[m->layer];
        locals->magnet[0] = l->magnet[0];
        locals->magnet[1] = l->magnet[1];
        locals->magnet[2] = l->magnet[2];
        if (l->density_model == PUMAS_B200_DENSITY_EXPONENTIAL) {
                locals->density = l->density * exp(-state->position[2] / l->scale);
                const double uz = fabs(state->direction[2]);
                return l->scale / ((uz <= l->eps) ? l->eps : uz);
        } else {
                locals->density = l->density;
                return 0.;
        }
}

static enum pumas_step oracle_medium_cb(struct pumas_context * context,
    struct pumas_state * state, struct pumas_medium ** medium_ptr,
    double * step_ptr)
{
        const struct pumas_b200_geometry * g = &oracle_geometry;
        if (g->kind == PUMAS_B200_GEOMETRY_UNIFORM) {
                if (medium_ptr != NULL) *medium_ptr = &oracle_media[0].api;
                if (step_ptr != NULL) *step_ptr = 0.;
                return (enum pumas_step)g->step_type;
        }

        if ((medium_ptr == NULL) && (step_ptr == NULL))
                return (enum pumas_step)g->step_type;

        const double z = state->position[2];
        const double sgn =
            (context->mode.direction == PUMAS_MODE_FORWARD) ? 1 : -1;
        const double uz = state->direction[2] * sgn;
        double step = g->step_default;
        const int n = g->n_layers;
        if ((z < g->z[0]) || !(z < g->z[n])) {
                if (medium_ptr != NULL) *medium_ptr = NULL;
                step = -1.;
        } else {
                int i = 0;
                while ((i < n - 1) && !(z < g->z[i + 1])) i++;
                if (medium_ptr != NULL) *medium_ptr = &oracle_media[i].api;
                if (uz > FLT_EPSILON)
                        step = (g->z[i + 1] - z) / uz;
                else if (uz < -FLT_EPSILON)
                        step = (g->z[i] - z) / uz;
        }
        if (step_ptr != NULL) {
                if (step > 0) step += g->step_epsilon;
                *step_ptr = step;
        }
        return (enum pumas_step)g->step_type;
}

static void oracle_silent_handler(
    enum pumas_return rc, pumas_function_t * caller, const char * message);

/* ---- user media (examples/user_media.h) through host callbacks -----------
 * The same source the device build of examples/user_media.cu compiles, here as
 * pumas_medium_cb / pumas_locals_cb of the reference (tests/test_user_media.py). */
#include "../examples/user_media.h"

static struct user_media oracle_user_media;
static int oracle_user_media_on = 0;

static double oracle_user_locals(struct pumas_medium * medium,
    struct pumas_state * state, struct pumas_locals * locals)
{
        const struct oracle_medium * m = (struct oracle_medium *)medium;
        locals->magnet[0] = locals->magnet[1] = locals->magnet[2] = 0.;
        return user_media_density(
            &oracle_user_media, m->layer, state->position, &locals->density);
}

static enum pumas_step oracle_user_medium_cb(struct pumas_context * context,
    struct pumas_state * state, struct pumas_medium ** medium_ptr,
    double * step_ptr)
{
        int medium = -1;
        const int backward =
            (context->mode.direction == PUMAS_MODE_FORWARD) ? 0 : 1;
        const int type = user_media_locate(&oracle_user_media, state->position,
            state->direction, backward, (medium_ptr != NULL) ? &medium : NULL,
            step_ptr);
        if (medium_ptr != NULL)
                *medium_ptr = (medium < 0) ? NULL : &oracle_media[medium].api;
        return (enum pumas_step)type;
}

/* Switch the next oracle_transport_n calls to the user media (`shells` != NULL;
 * materials[i]: material of shell i) or back to the descriptor (NULL). */
void oracle_set_user_media(const struct user_media * shells, const int * materials)
{
        if (shells == NULL) {
                oracle_user_media_on = 0;
                return;
        }
        memcpy(&oracle_user_media, shells, sizeof oracle_user_media);
        int i;
        for (i = 0; i < shells->n_shells; i++) {
                oracle_media[i].api.material = materials[i];
                oracle_media[i].api.locals = &oracle_user_locals;
                oracle_media[i].layer = i;
        }
        oracle_user_media_on = 1;
}

/* ---- user DCS model (examples/user_dcs.h), registered with the reference -- */
#include "../examples/user_dcs.h"

static double oracle_user_dcs(double Z, double A, double m, double K, double q)
{
        return user_dcs_bremsstrahlung(Z, A, m, K, q);
}

/* pumas_dcs_register (pumas.c:15422) of the model; idempotent. */
int oracle_register_user_dcs(void)
{
        static int done = 0;
        if (done) return PUMAS_RETURN_SUCCESS;
        pumas_handler_cb * previous = pumas_error_handler_get();
        pumas_error_handler_set(&oracle_silent_handler);
        const int rc = pumas_dcs_register(
            PUMAS_PROCESS_BREMSSTRAHLUNG, USER_DCS_MODEL, &oracle_user_dcs);
        pumas_error_handler_set(previous);
        if (rc == PUMAS_RETURN_SUCCESS) done = 1;
        return rc;
}

static void oracle_set_geometry(const struct pumas_b200_geometry * geometry)
{
        if (oracle_user_media_on) return;
        memcpy(&oracle_geometry, geometry, sizeof oracle_geometry);
        int i;
        for (i = 0; i < PUMAS_B200_MAX_LAYERS; i++) {
                oracle_media[i].api.material = geometry->layer[i].material;
                oracle_media[i].api.locals =
                    geometry->layer[i].has_locals ? &oracle_locals : NULL;
                oracle_media[i].layer = i;
        }
}

/* ---- batch driver ------------------------------------------------------- */
struct oracle_job {
        const struct pumas_physics * physics;
        const struct pumas_context * settings;
        uint64_t seed;
        uint64_t offset;
        long begin, end;
        struct pumas_state * states;
        int * events;
        int * media;
        unsigned int * n_random;
        int rc;
};

static void oracle_silent_handler(
    enum pumas_return rc, pumas_function_t * caller, const char * message)
{
        (void)rc;
        (void)caller;
        (void)message;
}

static void * oracle_worker(void * arg)
{
        struct oracle_job * job = arg;
        struct pumas_context * context = NULL;
        if (pumas_context_create(&context, job->physics,
                sizeof(struct oracle_user_data)) != PUMAS_RETURN_SUCCESS) {
                job->rc = PUMAS_RETURN_MEMORY_ERROR;
                return NULL;
        }
        context->mode = job->settings->mode;
        context->event = job->settings->event;
        context->limit = job->settings->limit;
        context->accuracy = job->settings->accuracy;
        context->medium =
            oracle_user_media_on ? &oracle_user_medium_cb : &oracle_medium_cb;
        context->random = &oracle_random;
        struct oracle_user_data * u = context->user_data;
        struct simulation_context * hidden =
            (struct simulation_context *)context;

        long i;
        for (i = job->begin; i < job->end; i++) {
                pb_philox_init(&u->rng, job->seed, job->offset + (uint64_t)i);
                hidden->randn_done = 0;
                hidden->randn_next = 0.;

                enum pumas_event event = PUMAS_EVENT_NONE;
                struct pumas_medium * media[2] = { NULL, NULL };
                const enum pumas_return rc = pumas_context_transport(
                    context, job->states + i, &event, media);
                if ((rc != PUMAS_RETURN_SUCCESS) && (job->rc == 0))
                        job->rc = rc;
                if (job->events != NULL) job->events[i] = (int)event;
                if (job->media != NULL) {
                        int k;
                        for (k = 0; k < 2; k++)
                                job->media[2 * i + k] = (media[k] == NULL) ?
                                    -1 :
                                    ((struct oracle_medium *)media[k])->layer;
                }
                if (job->n_random != NULL) job->n_random[i] = u->rng.count;
        }
        pumas_context_destroy(&context);
        return NULL;
}

/*
 * Transport n states through `geometry` with the reference engine.
 * `settings` is a `struct pumas_context` of which only mode, event, limit and
 * accuracy are read. Track i uses the Philox stream (seed, offset + i).
 * Returns the first non-success pumas_return seen (0 if none) and, when
 * `seconds` is not NULL, the wall-clock time of the transport loop only.
 */
int oracle_transport_n(const struct pumas_physics * physics,
    const struct pumas_context * settings,
    const struct pumas_b200_geometry * geometry, uint64_t seed,
    uint64_t offset, long n, struct pumas_state * states, int * events,
    int * media, unsigned int * n_random, int n_threads, double * seconds)
{
        if ((physics == NULL) || (settings == NULL) || (geometry == NULL) ||
            (states == NULL) || (n < 0))
                return PUMAS_RETURN_VALUE_ERROR;
        if (n_threads < 1) n_threads = 1;
        if (n_threads > 1024) n_threads = 1024;
        if ((long)n_threads > n) n_threads = (n > 0) ? (int)n : 1;

        oracle_set_geometry(geometry);
        pumas_handler_cb * previous = pumas_error_handler_get();
        pumas_error_handler_set(&oracle_silent_handler);

        struct oracle_job * jobs = calloc(n_threads, sizeof *jobs);
        pthread_t * threads = calloc(n_threads, sizeof *threads);
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        int t;
        for (t = 0; t < n_threads; t++) {
                struct oracle_job * j = jobs + t;
                j->physics = physics;
                j->settings = settings;
                j->seed = seed;
                j->offset = offset;
                j->begin = n * t / n_threads;
                j->end = n * (t + 1) / n_threads;
                j->states = states;
                j->events = events;
                j->media = media;
                j->n_random = n_random;
                j->rc = 0;
                if (n_threads == 1)
                        oracle_worker(j);
                else
                        pthread_create(threads + t, NULL, &oracle_worker, j);
        }
        int rc = 0;
        for (t = 0; t < n_threads; t++) {
                if (n_threads > 1) pthread_join(threads[t], NULL);
                if ((rc == 0) && (jobs[t].rc != 0)) rc = jobs[t].rc;
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (seconds != NULL)
                *seconds = (double)(t1.tv_sec - t0.tv_sec) +
                    1E-09 * (double)(t1.tv_nsec - t0.tv_nsec);

        free(jobs);
        free(threads);
        pumas_error_handler_set(previous);
        return rc;
}

/*
 * The same with a `pumas_recorder` of the given period attached to the context
 * (single thread): the frames of track i are copied, in order, to
 * frame_states / frame_media / frame_events starting at
 * sum(frame_counts[0 .. i)) (at most max_frames in all; the counts are always
 * complete). Media as layer indices, -1 for NULL.
 */
int oracle_transport_recorded(const struct pumas_physics * physics,
    const struct pumas_context * settings,
    const struct pumas_b200_geometry * geometry, uint64_t seed,
    uint64_t offset, long n, struct pumas_state * states, int * events,
    int period, long max_frames, struct pumas_state * frame_states,
    int * frame_media, int * frame_events, int * frame_counts)
{
        if ((physics == NULL) || (settings == NULL) || (geometry == NULL) ||
            (states == NULL) || (n < 0))
                return PUMAS_RETURN_VALUE_ERROR;
        oracle_set_geometry(geometry);
        pumas_handler_cb * previous = pumas_error_handler_get();
        pumas_error_handler_set(&oracle_silent_handler);

        struct pumas_context * context = NULL;
        struct pumas_recorder * recorder = NULL;
        if ((pumas_context_create(&context, physics,
                 sizeof(struct oracle_user_data)) != PUMAS_RETURN_SUCCESS) ||
            (pumas_recorder_create(&recorder, 0) != PUMAS_RETURN_SUCCESS))
                return PUMAS_RETURN_MEMORY_ERROR;
        recorder->period = period;
        context->recorder = recorder;
        context->mode = settings->mode;
        context->event = settings->event;
        context->limit = settings->limit;
        context->accuracy = settings->accuracy;
        context->medium = &oracle_medium_cb;
        context->random = &oracle_random;
        struct oracle_user_data * u = context->user_data;
        struct simulation_context * hidden =
            (struct simulation_context *)context;

        int rc = 0;
        long i, total = 0;
        for (i = 0; i < n; i++) {
                pb_philox_init(&u->rng, seed, offset + (uint64_t)i);
                hidden->randn_done = 0;
                hidden->randn_next = 0.;
                pumas_recorder_clear(recorder);
                enum pumas_event event = PUMAS_EVENT_NONE;
                const enum pumas_return r =
                    pumas_context_transport(context, states + i, &event, NULL);
                if ((r != PUMAS_RETURN_SUCCESS) && (rc == 0)) rc = r;
                if (events != NULL) events[i] = (int)event;
                const struct pumas_frame * frame;
                int count = 0;
                for (frame = recorder->first; frame != NULL;
                     frame = frame->next, count++) {
                        if (total >= max_frames) continue;
                        frame_states[total] = frame->state;
                        frame_media[total] = (frame->medium == NULL) ?
                            -1 :
                            ((struct oracle_medium *)frame->medium)->layer;
                        frame_events[total] = (int)frame->event;
                        total++;
                }
                frame_counts[i] = count;
        }
        pumas_recorder_destroy(&recorder);
        pumas_context_destroy(&context);
        pumas_error_handler_set(previous);
        return rc;
}

/* The Philox stream as the reference's random callback sees it. */
void oracle_philox_uniforms(uint64_t seed, uint64_t track, int n, double * out)
{
        struct pumas_context context;
        struct oracle_user_data u;
        memset(&context, 0x0, sizeof context);
        context.user_data = &u;
        pb_philox_init(&u.rng, seed, track);
        int i;
        for (i = 0; i < n; i++) out[i] = oracle_random(&context);
}

/* Load a physics dump by path (ctypes has no FILE *). */
int oracle_physics_load(struct pumas_physics ** physics, const char * path)
{
        FILE * stream = fopen(path, "rb");
        if (stream == NULL) return PUMAS_RETURN_PATH_ERROR;
        pumas_error_handler_set(&oracle_silent_handler);
        const int rc = pumas_physics_load(physics, stream);
        fclose(stream);
        return rc;
}

/* Create a physics from an MDF and dump it (build-time tool entry). */
int oracle_physics_create_and_dump(int particle, const char * mdf_path,
    const char * dedx_path, const char * dump_path)
{
        struct pumas_physics * physics = NULL;
        int rc = pumas_physics_create(&physics, (enum pumas_particle)particle,
            mdf_path, dedx_path, NULL);
        if (rc != PUMAS_RETURN_SUCCESS) return rc;
        FILE * stream = fopen(dump_path, "wb");
        if (stream == NULL) {
                pumas_physics_destroy(&physics);
                return PUMAS_RETURN_PATH_ERROR;
        }
        rc = pumas_physics_dump(physics, stream);
        fclose(stream);
        pumas_physics_destroy(&physics);
        return rc;
}

/* Same, with the DCS models named (NULL = default). pumas.h:742-772. */
int oracle_physics_create_and_dump_models(int particle, const char * mdf_path,
    const char * dump_path, const char * bremsstrahlung,
    const char * pair_production, const char * photonuclear)
{
        struct pumas_physics_settings settings;
        memset(&settings, 0x0, sizeof settings);
        settings.bremsstrahlung = bremsstrahlung;
        settings.pair_production = pair_production;
        settings.photonuclear = photonuclear;
        struct pumas_physics * physics = NULL;
        int rc = pumas_physics_create(&physics, (enum pumas_particle)particle,
            mdf_path, NULL, &settings);
        if (rc != PUMAS_RETURN_SUCCESS) return rc;
        FILE * stream = fopen(dump_path, "wb");
        if (stream == NULL) {
                pumas_physics_destroy(&physics);
                return PUMAS_RETURN_PATH_ERROR;
        }
        rc = pumas_physics_dump(physics, stream);
        fclose(stream);
        pumas_physics_destroy(&physics);
        return rc;
}

/* Raw table access for loader parity tests: returns the address and length
 * (in doubles) of the named table of the reference's physics object. */
const double * oracle_table(
    const struct pumas_physics * p, const char * name, long * length)
{
        const long ne = p->n_energies, nm = p->n_materials;
#define T(tag, ptr, len)                                                       \
        if (strcmp(name, tag) == 0) {                                          \
                *length = (len);                                               \
                return (ptr);                                                  \
        }
        T("K", p->table_K, ne)
        T("K_dX", p->table_K_dX, N_SCHEMES * nm * ne)
        T("X", p->table_X, N_SCHEMES * nm * ne)
        T("X_dK", p->table_X_dK, N_SCHEMES * nm * ne)
        T("T", p->table_T, N_SCHEMES * nm * ne)
        T("dE", p->table_dE, N_SCHEMES * nm * ne)
        T("Omega", p->table_Omega, nm * ne)
        T("CS", p->table_CS, nm * ne)
        T("NI_in", p->table_NI_in, nm * ne)
        T("NI_el", p->table_NI_el, N_SCHEMES * nm * ne)
        T("Mu0", p->table_Mu0, nm * ne)
        T("Lb", p->table_Lb, nm * ne)
        T("Ms1", p->table_Ms1, (N_SCHEMES + 1) * nm * ne)
        T("Kt", p->table_Kt, nm)
        T("a_max", p->table_a_max, nm)
        T("b_max", p->table_b_max, N_SCHEMES * nm)
#undef T
        *length = 0;
        return NULL;
}

/* Exposed static helpers of the reference, for unit level parity checks of
 * the device functions (same arguments as the originals). */
double oracle_dcs_evaluate(const struct pumas_physics * physics, int process,
    int element, double K, double q)
{
        struct pumas_context * context = NULL;
        pumas_context_create(&context, physics, 0);
        const double d = dcs_evaluate(physics, context, dcs_get(process),
            physics->element[element], K, q);
        pumas_context_destroy(&context);
        return d;
}

double oracle_dcs_model(const struct pumas_physics * physics, int process,
    int element, double K, double q)
{
        return dcs_get(process)(physics, physics->element[element], K, q);
}

double oracle_coulomb_ehs_length(
    const struct pumas_physics * physics, int material, double K)
{
        return coulomb_ehs_length(physics, NULL, material, K);
}

void oracle_table_get_msc(const struct pumas_physics * physics, int scheme,
    int material, double K, double * mu0, double * invlb1)
{
        struct pumas_context * context = NULL;
        pumas_context_create(&context, physics, 0);
        context->mode.energy_loss = (enum pumas_mode)scheme;
        table_get_msc(physics, context, material, K, mu0, invlb1);
        pumas_context_destroy(&context);
}

/* Hard Coulomb cross-sections per element of a material at kinetic energy K,
 * as computed inside transport_do_ehs (pumas.c:5786-5813). */
int oracle_coulomb_cs_hard(const struct pumas_physics * physics, int material,
    int scheme, double kinetic, double * cs_hard)
{
        struct pumas_context * context = NULL;
        pumas_context_create(&context, physics, 0);
        context->mode.energy_loss = (enum pumas_mode)scheme;
        double mu0 = 0., invlb1 = 0.;
        table_get_msc(physics, context, material, kinetic, &mu0, &invlb1);
        int i;
        for (i = 0; i < physics->elements_in[material]; i++) {
                struct coulomb_data data_, * data = &data_;
                const struct material_component * const component =
                    &physics->composition[material][i];
                const struct atomic_element * const element =
                    physics->element[component->element];
                double kinetic0;
                coulomb_frame_parameters(element->Z, element->A, physics->mass,
                    kinetic, &kinetic0, data->fCM);
                data->fspin = coulomb_spin_factor(physics->mass, kinetic);
                coulomb_screening_parameters(element->Z, element->A,
                    physics->mass, kinetic, kinetic0, &data->n_parameters,
                    data->amplitude, data->screening);
                coulomb_pole_reduction(data->n_parameters, data->amplitude,
                    data->screening, data->a, data->b, data->c);
                data->normalisation = component->fraction /
                    coulomb_normalisation(element->Z, element->A,
                        physics->mass, kinetic, kinetic0);
                cs_hard[i] = data->normalisation *
                    coulomb_restricted_cs(mu0, data->fspin,
                        data->n_parameters, data->screening, data->a, data->b,
                        data->c);
        }
        pumas_context_destroy(&context);
        return physics->elements_in[material];
}
