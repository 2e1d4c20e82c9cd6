/*
 * pumas_b200.h -- C ABI of libpumas_b200.so
 *
 * The drop-in boundary for the per-track transport engine of PUMAS rebuilt as
 * sm_100a CUDA kernels. Every type below has the exact binary layout of its
 * namesake in the reference header (include/pumas.h, v1.2.3), so user code
 * compiled against the reference header links against this library unchanged
 * for the entry points listed here. Citations `pumas.h:N` / `pumas.c:N` are
 * relative to the reference tree (include/pumas.h, src/pumas.c).
 *
 * What differs from the reference:
 *   - `pumas_context_transport_n` (new): batched transport, one GPU thread per
 *     track, whole step loop on the device.
 *   - media are described by a POD `struct pumas_b200_geometry` (host function
 *     pointers cannot be called from a kernel); `pumas_context::medium` is
 *     ignored by the batched entry point.
 *   - randomness is a counter based Philox4x32-10 stream keyed by
 *     (context seed, track id); `pumas_context::random` is ignored on device.
 *   - there is NO CPU fallback: transport entry points return
 *     PUMAS_RETURN_INTERNAL_ERROR ("no CUDA device") when no GPU is usable.
 */
#ifndef PUMAS_B200_H
#define PUMAS_B200_H

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PUMAS_API
#define PUMAS_API
#endif

#define PUMAS_VERSION_MAJOR 1
#define PUMAS_VERSION_MINOR 2
#define PUMAS_VERSION_PATCH 3

/* pumas.h:50-56 */
enum pumas_particle { PUMAS_PARTICLE_MUON = 0, PUMAS_PARTICLE_TAU };

/* pumas.h:59-87 */
enum pumas_property {
        PUMAS_PROPERTY_CROSS_SECTION = 0,
        PUMAS_PROPERTY_ELASTIC_CUTOFF_ANGLE,
        PUMAS_PROPERTY_ELASTIC_PATH,
        PUMAS_PROPERTY_STOPPING_POWER,
        PUMAS_PROPERTY_RANGE,
        PUMAS_PROPERTY_KINETIC_ENERGY,
        PUMAS_PROPERTY_MAGNETIC_ROTATION,
        PUMAS_PROPERTY_TRANSPORT_PATH,
        PUMAS_PROPERTY_PROPER_TIME
};

/* pumas.h:90-140 */
enum pumas_mode {
        PUMAS_MODE_DISABLED = -1,
        PUMAS_MODE_CSDA = 0,
        PUMAS_MODE_MIXED = 1,
        PUMAS_MODE_STRAGGLED = 2,
        PUMAS_MODE_WEIGHTED = 0,
        PUMAS_MODE_RANDOMISED = 1,
        PUMAS_MODE_FORWARD = 0,
        PUMAS_MODE_BACKWARD = 1
};

/* pumas.h:143-196 */
enum pumas_return {
        PUMAS_RETURN_SUCCESS = 0,
        PUMAS_RETURN_ACCURACY_ERROR,
        PUMAS_RETURN_END_OF_FILE,
        PUMAS_RETURN_DECAY_ERROR,
        PUMAS_RETURN_DENSITY_ERROR,
        PUMAS_RETURN_DIRECTION_ERROR,
        PUMAS_RETURN_INCOMPLETE_FILE,
        PUMAS_RETURN_INDEX_ERROR,
        PUMAS_RETURN_PHYSICS_ERROR,
        PUMAS_RETURN_INTERNAL_ERROR,
        PUMAS_RETURN_IO_ERROR,
        PUMAS_RETURN_FORMAT_ERROR,
        PUMAS_RETURN_MEDIUM_ERROR,
        PUMAS_RETURN_MEMORY_ERROR,
        PUMAS_RETURN_MODEL_ERROR,
        PUMAS_RETURN_MISSING_LIMIT,
        PUMAS_RETURN_MISSING_RANDOM,
        PUMAS_RETURN_PATH_ERROR,
        PUMAS_RETURN_RAISE_ERROR,
        PUMAS_RETURN_TOO_LONG,
        PUMAS_RETURN_UNDEFINED_MDF,
        PUMAS_RETURN_UNKNOWN_ELEMENT,
        PUMAS_RETURN_UNKNOWN_MATERIAL,
        PUMAS_RETURN_UNKNOWN_PARTICLE,
        PUMAS_RETURN_VALUE_ERROR,
        PUMAS_N_RETURNS
};

/* pumas.h:199-236 */
enum pumas_event {
        PUMAS_EVENT_NONE = 0,
        PUMAS_EVENT_LIMIT_ENERGY = 1,
        PUMAS_EVENT_LIMIT_DISTANCE = 2,
        PUMAS_EVENT_LIMIT_GRAMMAGE = 4,
        PUMAS_EVENT_LIMIT_TIME = 8,
        PUMAS_EVENT_LIMIT = 15,
        PUMAS_EVENT_MEDIUM = 16,
        PUMAS_EVENT_VERTEX_BREMSSTRAHLUNG = 32,
        PUMAS_EVENT_VERTEX_PAIR_CREATION = 64,
        PUMAS_EVENT_VERTEX_PHOTONUCLEAR = 128,
        PUMAS_EVENT_VERTEX_DELTA_RAY = 256,
        PUMAS_EVENT_VERTEX_DEL = 480,
        PUMAS_EVENT_VERTEX_COULOMB = 512,
        PUMAS_EVENT_VERTEX_DECAY = 1024,
        PUMAS_EVENT_VERTEX = 2016,
        PUMAS_EVENT_WEIGHT = 2048,
        PUMAS_EVENT_START = 4096,
        PUMAS_EVENT_STOP = 8192
};

/* pumas.h:239-246 */
enum pumas_process {
        PUMAS_PROCESS_BREMSSTRAHLUNG = 0,
        PUMAS_PROCESS_PAIR_PRODUCTION,
        PUMAS_PROCESS_PHOTONUCLEAR
};

/* pumas.h:249-278 */
enum pumas_constant {
        PUMAS_CONSTANT_ALPHA_EM = 0,
        PUMAS_CONSTANT_AVOGADRO_NUMBER,
        PUMAS_CONSTANT_BOHR_RADIUS,
        PUMAS_CONSTANT_ELECTRON_MASS,
        PUMAS_CONSTANT_ELECTRON_RADIUS,
        PUMAS_CONSTANT_HBAR_C,
        PUMAS_CONSTANT_MUON_C_TAU,
        PUMAS_CONSTANT_MUON_MASS,
        PUMAS_CONSTANT_NEUTRON_MASS,
        PUMAS_CONSTANT_PION_MASS,
        PUMAS_CONSTANT_PROTON_MASS,
        PUMAS_CONSTANT_TAU_C_TAU,
        PUMAS_CONSTANT_TAU_MASS,
        PUMAS_N_CONSTANTS
};

/* pumas.h:289-309 -- 12 doubles + int, 104 bytes with tail padding. */
struct pumas_state {
        double charge;
        double energy;       /* kinetic, GeV */
        double distance;     /* m */
        double grammage;     /* kg/m^2 */
        double time;         /* proper time, m/c */
        double weight;
        double position[3];  /* m */
        double direction[3];
        int decayed;
};

/* pumas.h:314-319 */
struct pumas_locals {
        double density;      /* kg/m^3 */
        double magnet[3];    /* T */
};

struct pumas_medium;
/* pumas.h:348-349 */
typedef double pumas_locals_cb(struct pumas_medium * medium,
    struct pumas_state * state, struct pumas_locals * locals);

/* pumas.h:363-376 */
struct pumas_medium {
        int material;
        pumas_locals_cb * locals;
};

/* pumas.h:380-392 */
struct pumas_frame {
        struct pumas_state state;
        struct pumas_medium * medium;
        enum pumas_event event;
        struct pumas_frame * next;
};

struct pumas_context;
/* pumas.h:409-411 */
typedef void pumas_recorder_cb(struct pumas_context * context,
    struct pumas_state * state, struct pumas_medium * medium,
    enum pumas_event event);

/* pumas.h:432-458 */
struct pumas_recorder {
        struct pumas_frame * first;
        int length;
        int period;
        pumas_recorder_cb * record;
        void * user_data;
};

/* pumas.h:461-466 */
enum pumas_step { PUMAS_STEP_CHECK = 0, PUMAS_STEP_RAW };

/* pumas.h:509-511 */
typedef enum pumas_step pumas_medium_cb(struct pumas_context * context,
    struct pumas_state * state, struct pumas_medium ** medium, double * step);

typedef void pumas_function_t(void);
/* pumas.h:529-530 */
typedef void pumas_handler_cb(
    enum pumas_return rc, pumas_function_t * caller, const char * message);
/* pumas.h:551 */
typedef double pumas_random_cb(struct pumas_context * context);

/* pumas.h:554-617 */
struct pumas_context_mode {
        enum pumas_mode energy_loss;
        enum pumas_mode decay;
        enum pumas_mode direction;
        enum pumas_mode scattering;
};

struct pumas_context_limit {
        double energy;
        double distance;
        double grammage;
        double time;
};

/* pumas.h:636-672 */
struct pumas_context {
        pumas_medium_cb * medium;
        pumas_random_cb * random;
        struct pumas_recorder * recorder;
        void * user_data;
        struct pumas_context_mode mode;
        enum pumas_event event;
        struct pumas_context_limit limit;
        double accuracy;
};

struct pumas_physics;

/* pumas.h:721-815 (only accepted as NULL by this library) */
struct pumas_physics_settings {
        double cutoff;
        double elastic_ratio;
        const char * bremsstrahlung;
        const char * pair_production;
        const char * photonuclear;
        int n_energies;
        double * energy;
        int update;
        int dry;
};

#include "pumas_b200_geometry.h"

/* ------------------------------------------------------------------------
 * Physics (pumas.h:898-1000, 1106, 1368-1640, 1640-1960)
 * ------------------------------------------------------------------------ */
/* Replaces pumas_physics_load (pumas.c:2049): parses the tag-13 binary dump,
 * keeps a host mirror and uploads the transport tables to the current CUDA
 * device (lazily, on the first transport call). */
PUMAS_API enum pumas_return pumas_physics_load(
    struct pumas_physics ** physics, FILE * stream);
/* Replaces pumas_physics_dump (pumas.c:2167): writes back the bytes loaded. */
PUMAS_API enum pumas_return pumas_physics_dump(
    const struct pumas_physics * physics, FILE * stream);
/* Replaces pumas_physics_create (pumas.c:1916). Table generation is out of the
 * hot-path scope: returns PUMAS_RETURN_MODEL_ERROR. */
PUMAS_API enum pumas_return pumas_physics_create(
    struct pumas_physics ** physics, enum pumas_particle particle,
    const char * mdf_path, const char * dedx_path,
    const struct pumas_physics_settings * settings);
/* pumas.c:2198 */
PUMAS_API void pumas_physics_destroy(struct pumas_physics ** physics);
/* pumas.c:2213, 2218 */
PUMAS_API double pumas_physics_cutoff(const struct pumas_physics * physics);
PUMAS_API double pumas_physics_elastic_ratio(
    const struct pumas_physics * physics);
/* pumas.c:3254 */
PUMAS_API enum pumas_return pumas_physics_particle(
    const struct pumas_physics * physics, enum pumas_particle * particle,
    double * lifetime, double * mass);

/* Property getters, pumas.c:2854-3111. Host evaluations of the same
 * interpolation code the kernels run. */
PUMAS_API enum pumas_return pumas_physics_property_range(
    const struct pumas_physics * physics, enum pumas_mode mode, int material,
    double energy, double * range);
PUMAS_API enum pumas_return pumas_physics_property_proper_time(
    const struct pumas_physics * physics, enum pumas_mode mode, int material,
    double energy, double * time);
PUMAS_API enum pumas_return pumas_physics_property_magnetic_rotation(
    const struct pumas_physics * physics, int material, double energy,
    double * angle);
PUMAS_API enum pumas_return pumas_physics_property_kinetic_energy(
    const struct pumas_physics * physics, enum pumas_mode mode, int material,
    double range, double * energy);
PUMAS_API enum pumas_return pumas_physics_property_stopping_power(
    const struct pumas_physics * physics, enum pumas_mode mode, int material,
    double energy, double * dedx);
PUMAS_API enum pumas_return pumas_physics_property_energy_straggling(
    const struct pumas_physics * physics, int material, double energy,
    double * straggling);
PUMAS_API enum pumas_return pumas_physics_property_elastic_cutoff_angle(
    const struct pumas_physics * physics, int material, double energy,
    double * angle);
PUMAS_API enum pumas_return pumas_physics_property_elastic_path(
    const struct pumas_physics * physics, int material, double energy,
    double * length);
PUMAS_API enum pumas_return pumas_physics_property_transport_path(
    const struct pumas_physics * physics, enum pumas_mode mode, int material,
    double energy, double * path);
PUMAS_API enum pumas_return pumas_physics_property_cross_section(
    const struct pumas_physics * physics, int material, double energy,
    double * cross_section);

/* pumas.c:3270-3400 */
PUMAS_API int pumas_physics_element_length(const struct pumas_physics * physics);
PUMAS_API enum pumas_return pumas_physics_element_properties(
    const struct pumas_physics * physics, int index, double * Z, double * A,
    double * I);
PUMAS_API enum pumas_return pumas_physics_element_name(
    const struct pumas_physics * physics, int index, const char ** element);
PUMAS_API enum pumas_return pumas_physics_element_index(
    const struct pumas_physics * physics, const char * element, int * index);
PUMAS_API enum pumas_return pumas_physics_material_name(
    const struct pumas_physics * physics, int index, const char ** material);
PUMAS_API enum pumas_return pumas_physics_material_index(
    const struct pumas_physics * physics, const char * material, int * index);
PUMAS_API int pumas_physics_material_length(
    const struct pumas_physics * physics);
PUMAS_API int pumas_physics_composite_length(
    const struct pumas_physics * physics);
PUMAS_API enum pumas_return pumas_physics_material_properties(
    const struct pumas_physics * physics, int index, int * length,
    double * density, double * I, int * components, double * fractions);

/* pumas.c:3446-3475: base materials and mass fractions of a composite (the
 * last pumas_physics_composite_length materials). */
PUMAS_API enum pumas_return pumas_physics_composite_properties(
    const struct pumas_physics * physics, int material, int * length,
    int * components, double * fractions);

/* pumas.c:3486-3625 */
PUMAS_API enum pumas_return pumas_physics_table_value(
    const struct pumas_physics * physics, enum pumas_property property,
    enum pumas_mode mode, int material, int row, double * value);
PUMAS_API int pumas_physics_table_length(const struct pumas_physics * physics);
PUMAS_API enum pumas_return pumas_physics_table_index(
    const struct pumas_physics * physics, enum pumas_property property,
    enum pumas_mode mode, int material, double value, int * index);

/* ------------------------------------------------------------------------
 * Differential cross-section models (pumas.h:717, 2110-2284), host side.
 * The models the device code implements (pb_dcs.h, pb_dcs_models.h) are
 * exposed with the reference's signature: atomic DCS in m^2 / GeV for a target
 * (Z, A), a projectile of mass m and kinetic energy K [GeV], energy loss q.
 * ------------------------------------------------------------------------ */
/* pumas.h:717 */
typedef double pumas_dcs_t(double Z, double A, double m, double K, double q);

/* pumas.c:13673-13699: the model (name, function) a physics was built with. */
PUMAS_API enum pumas_return pumas_physics_dcs(
    const struct pumas_physics * physics, enum pumas_process process,
    const char ** model, pumas_dcs_t ** dcs);

/* pumas.c:15469-15505: look up a model by name (NULL: the default one) among
 * the built-in ones (bremsstrahlung ABB, KKP, SSR; pair production KKP, SSR;
 * photonuclear BBKS, BM, DRSS). */
PUMAS_API enum pumas_return pumas_dcs_get(
    enum pumas_process process, const char * model, pumas_dcs_t ** dcs);

/* pumas.c:15422-15466. Not supported: a host function cannot be evaluated by
 * the transport kernels; returns PUMAS_RETURN_MODEL_ERROR. */
PUMAS_API enum pumas_return pumas_dcs_register(
    enum pumas_process process, const char * model, pumas_dcs_t * dcs);

/* pumas.c:15532-15542 ("SSR", "SSR", "DRSS"); NULL for a bad process. */
PUMAS_API const char * pumas_dcs_default(enum pumas_process process);

/* pumas.c:15508-15529: kinematic range of the energy loss. */
PUMAS_API enum pumas_return pumas_dcs_range(enum pumas_process process,
    double Z, double mass, double energy, double * min, double * max);

/* pumas.c:15545-15581: elastic (Coulomb) DCS of the hard scattering sampled
 * by the transport (screened nucleus, spin and nuclear form factors), in
 * m^2 / rad, at the laboratory angle theta. */
PUMAS_API double pumas_elastic_dcs(
    double Z, double A, double m, double K, double theta);

/* pumas.c:15584-15609: elastic path length (order 0) or first transport path
 * length (order 1), in kg/m^2, for a target (Z, A). Evaluated on the host in
 * `long double` like the reference. */
PUMAS_API double pumas_elastic_path(
    int order, double Z, double A, double mass, double kinetic);

/* pumas.c:11082-11101: close collisions with atomic electrons. */
PUMAS_API double pumas_electronic_dcs(
    double Z, double I, double m, double K, double q);

/* ------------------------------------------------------------------------
 * Context (pumas.h:1214-1340)
 * ------------------------------------------------------------------------ */
/* pumas.c:2545. Same defaults (pumas.c:2592-2605). */
PUMAS_API enum pumas_return pumas_context_create(
    struct pumas_context ** context, const struct pumas_physics * physics,
    int extra_memory);
/* pumas.c:2617 */
PUMAS_API void pumas_context_destroy(struct pumas_context ** context);
/* pumas.c:2629 */
PUMAS_API const struct pumas_physics * pumas_context_physics_get(
    const struct pumas_context * context);
/* pumas.c:2362 / 2393. The seed is the Philox key of the device streams. */
PUMAS_API enum pumas_return pumas_context_random_seed_set(
    struct pumas_context * context, const unsigned long * seed);
PUMAS_API enum pumas_return pumas_context_random_seed_get(
    struct pumas_context * context, unsigned long * seed);

/* pumas.c:2415-2492. The random state of a context is (seed, id of its next
 * first track): the streams are counter based, so that is all a later run
 * needs to resume with the same variates. The format differs from the
 * reference's (a Mersenne Twister state). */
PUMAS_API enum pumas_return pumas_context_random_dump(
    struct pumas_context * context, FILE * stream);
PUMAS_API enum pumas_return pumas_context_random_load(
    struct pumas_context * context, FILE * stream);

/* New. Attach a built-in geometry (copied) to the context; replaces setting
 * `context->medium` (pumas.h:641). */
PUMAS_API enum pumas_return pumas_b200_context_geometry_set(
    struct pumas_context * context,
    const struct pumas_b200_geometry * geometry);
/* New. `__device__` medium / locals functions (struct pumas_b200_callbacks,
 * pumas_b200_geometry.h): the device form of `context->medium` (pumas.h:641,
 * 509-511) and of `pumas_medium::locals` (pumas.h:348-349, 466-480) for
 * geometries the descriptor does not cover. The table is copied; it takes
 * precedence over a descriptor; NULL removes it. The functions must belong to
 * the device link unit of the kernels: the built-in planar client below, or
 * user code built in with `python -m pumas_b200.build --user file.cu`
 * (INTEGRATION.md section 6). */
PUMAS_API enum pumas_return pumas_b200_context_callbacks_set(
    struct pumas_context * context,
    const struct pumas_b200_callbacks * callbacks);
/* New. The planar media of `geometry` behind the callback interface (the
 * library's own client of it): fills `callbacks`, whose user_data is a device
 * copy of the descriptor, released by pumas_b200_planar_callbacks_free. */
PUMAS_API enum pumas_return pumas_b200_planar_callbacks(
    const struct pumas_b200_geometry * geometry,
    struct pumas_b200_callbacks * callbacks);
PUMAS_API void pumas_b200_planar_callbacks_free(
    struct pumas_b200_callbacks * callbacks);
/* New. The `__device__` counterpart of a DCS model registered with
 * pumas_dcs_register (pumas.h:2150-2167): same signature, same value, as a
 * device function pointer of the device link unit of the kernels (user code
 * built in with `python -m pumas_b200.build --user file.cu`). A physics whose
 * tables were made with a user model (the dump names it) loads once the model
 * is registered and transports once its device function is. */
PUMAS_API enum pumas_return pumas_b200_dcs_register_device(
    enum pumas_process process, const char * model, pumas_dcs_t * device_dcs);
/* New. Id of the first track of the next batch: track i of a batch draws from
 * the Philox stream (seed, offset + i). Lets ranks/batches own disjoint,
 * reproducible id ranges. */
PUMAS_API enum pumas_return pumas_b200_context_track_offset_set(
    struct pumas_context * context, unsigned long long offset);
/* New. Run the kernels of this context on the given cudaStream_t (as void*);
 * NULL is the legacy default stream. */
PUMAS_API enum pumas_return pumas_b200_context_stream_set(
    struct pumas_context * context, void * cuda_stream);

/* New. Select the device kernel: 0 (default) the persistent kernels that
 * regroup tracks by pending event: tracks resident in shared memory for the
 * straggled mode in a uniform medium (hot_kernel.cuh), in registers otherwise
 * (resident.cuh); 1 the monolithic kernel, one thread per track for its whole
 * life (kept for A/B parity and profiling); 2 the register resident kernel
 * without compile time specialisation; 3 the register resident kernel with
 * its specialisations (round 1 product path, A/B). All give the same
 * per-track results. */
PUMAS_API enum pumas_return pumas_b200_context_kernel_set(
    struct pumas_context * context, int kernel);

/* ------------------------------------------------------------------------
 * Transport
 * ------------------------------------------------------------------------ */
/* Replaces pumas_context_transport (pumas.c:3114) for a built-in geometry:
 * one track, n = 1 batch. `media`, when not NULL, receives NULL pointers (the
 * device media are indices, see pumas_context_transport_n). */
PUMAS_API enum pumas_return pumas_context_transport(
    struct pumas_context * context, struct pumas_state * state,
    enum pumas_event * event, struct pumas_medium * media[2]);

/* New: batched transport (BASELINE.json north_star). Same per-track semantics
 * as pumas_context_transport (pumas.c:3114-3251).
 *   states  n x struct pumas_state, host or device memory, updated in place.
 *   events  n x int (enum pumas_event of the end condition), host or device,
 *           may be NULL.
 *   media   n x int[2]: layer index of the first and last medium, -1 for
 *           outside (the reference's media[2] pointers), may be NULL.
 * Host buffers are staged through device memory inside the call. */
PUMAS_API enum pumas_return pumas_context_transport_n(
    struct pumas_context * context, long n, struct pumas_state * states,
    int * events, int * media);

/* New: per-track diagnostics of the last batch of this context, host copy.
 * n_random[i] = uniform variates consumed by track i, n_steps[i] = calls of
 * the step function. Either pointer may be NULL. */
PUMAS_API enum pumas_return pumas_b200_context_last_counters(
    struct pumas_context * context, long n, unsigned int * n_random,
    unsigned int * n_steps);

/* New: seconds of GPU time (CUDA events on the context stream) spent in the
 * transport kernel(s) of the last batch, and the number of kernels launched. */
PUMAS_API enum pumas_return pumas_b200_context_last_kernel_time(
    struct pumas_context * context, double * milliseconds, int * launches);

/* ------------------------------------------------------------------------
 * Errors and misc (pumas.h:1120-1184, 1090, 2189)
 * ------------------------------------------------------------------------ */
PUMAS_API const char * pumas_error_function(pumas_function_t * function);
PUMAS_API void pumas_error_handler_set(pumas_handler_cb * handler);
PUMAS_API pumas_handler_cb * pumas_error_handler_get(void);
PUMAS_API void pumas_error_catch(int enable);
PUMAS_API enum pumas_return pumas_error_raise(void);
/* pumas.h:2020-2079: user supplied allocation of the host objects the library
 * hands out (physics, contexts and their user_data). NULL restores malloc /
 * realloc / free. Not thread safe, like the reference. */
typedef void * pumas_allocate_cb(size_t size);
typedef void * pumas_reallocate_cb(void * ptr, size_t size);
typedef void pumas_deallocate_cb(void * ptr);
PUMAS_API void pumas_memory_allocator(pumas_allocate_cb * allocator);
PUMAS_API void pumas_memory_reallocator(pumas_reallocate_cb * reallocator);
PUMAS_API void pumas_memory_deallocator(pumas_deallocate_cb * deallocator);

PUMAS_API void pumas_version(int * major, int * minor, int * patch);
PUMAS_API enum pumas_return pumas_constant(
    enum pumas_constant index, double * value);

/* New: tallies of a transported batch, on the device (what the user loop of
 * examples/pumas/straight.c:175-178 accumulates per track). `states` and
 * `events` are device pointers. `out` (device, 4 + n_bins doubles) receives
 * [sum w, sum w^2, n, n selected, counts per log-spaced energy bin ...] over the
 * tracks whose end event intersects `event_mask` (0 = all); reduction order is
 * fixed, so results are reproducible. These are the values the ranks of a
 * multi-GPU run all-reduce. */
PUMAS_API long pumas_b200_tally_scratch_bytes(int n_bins);
PUMAS_API enum pumas_return pumas_b200_tally(const struct pumas_state * states,
    const int * events, long n, int event_mask, double energy_min,
    double energy_max, int n_bins, double * out, void * scratch,
    void * cuda_stream);

/* pumas.c:2797-2851. A recorder attached to `context->recorder` receives the
 * frames of the tracks of pumas_context_transport / _transport_n (pumas.h:
 * 432-458: every `period`-th Monte Carlo step, the events, START and STOP),
 * in track order, as the reference's linked list of `pumas_frame` or through
 * the `record` callback (called on the host after the batch). `frame->medium`
 * points to a `pumas_medium` of the context (material of the layer), NULL
 * outside. Recorded batches run the one-thread-per-track kernel twice (count,
 * then store): a debugging tool, as in the reference. */
PUMAS_API enum pumas_return pumas_recorder_create(
    struct pumas_recorder ** recorder, int extra_memory);
PUMAS_API void pumas_recorder_clear(struct pumas_recorder * recorder);
PUMAS_API void pumas_recorder_destroy(struct pumas_recorder ** recorder);
/* New: the frames as flat arrays, media as layer indices (bindings). Returns
 * the number of frames of the recorder; at most `max` are copied. */
PUMAS_API int pumas_b200_recorder_export(const struct pumas_context * context,
    const struct pumas_recorder * recorder, long max,
    struct pumas_state * states, int * media, int * events);

/* New. pumas_context_transport_n without the wait: the batch is enqueued on
 * the context's stream and the call returns. `states`, `events`, `media` must
 * be device pointers. pumas_b200_context_synchronize waits for every batch
 * enqueued since the last wait and raises the first error one of their tracks
 * met (the return code pumas_context_transport_n would have given). */
PUMAS_API enum pumas_return pumas_b200_context_transport_async(
    struct pumas_context * context, long n, struct pumas_state * states,
    int * events, int * media);
PUMAS_API enum pumas_return pumas_b200_context_synchronize(
    struct pumas_context * context);

/* New: multi-GPU runs (one rank per GPU, tracks sharded by id range with
 * pumas_b200_context_track_offset_set). The tallies of
 * the batches of a rank are accumulated on the device
 * (pumas_b200_tally_accumulate: total[k] += batch[k], in launch order) and
 * summed over the ranks ONCE, at the end, by pumas_b200_comm_allreduce
 * (ncclAllReduce of doubles, in place, on `cuda_stream`). The communicator is
 * created from a unique id made on one rank and handed to the others by the
 * caller (MPI, a file, torch.distributed ...), or wraps the caller's own
 * ncclComm_t. NCCL is loaded at run time (libnccl.so.2, or the library named
 * by the environment variable PUMAS_B200_NCCL); without it these calls return
 * PUMAS_RETURN_VALUE_ERROR and pumas_b200_comm_available() is 0. */
struct pumas_b200_comm;
#define PUMAS_B200_UNIQUE_ID_BYTES 128
PUMAS_API int pumas_b200_comm_available(void);
PUMAS_API enum pumas_return pumas_b200_comm_unique_id(void * id, int bytes);
PUMAS_API enum pumas_return pumas_b200_comm_create(
    struct pumas_b200_comm ** comm, int rank, int world, const void * id,
    int bytes);
PUMAS_API enum pumas_return pumas_b200_comm_wrap(
    struct pumas_b200_comm ** comm, void * nccl_comm, int rank, int world);
PUMAS_API enum pumas_return pumas_b200_comm_allreduce(
    struct pumas_b200_comm * comm, double * values, long count,
    void * cuda_stream);
PUMAS_API void pumas_b200_comm_destroy(struct pumas_b200_comm ** comm);
PUMAS_API enum pumas_return pumas_b200_tally_accumulate(
    double * total, const double * batch, long count, void * cuda_stream);

/* New: tally over a grid of bins (examples/pumas/geometry.c extended to a
 * zenith x azimuth grid, BASELINE.json configs[3]): the batch holds `per_bin`
 * consecutive tracks per bin; out[3 b .. 3 b + 2] (device) receives the sum of
 * the weights, of their squares and the number of the tracks of bin b whose
 * end event intersects `event_mask` (0 = all). Fixed reduction order. Ranks
 * shard the grid by bin and by track id and all-reduce `out`. */
PUMAS_API enum pumas_return pumas_b200_tally_segments(
    const struct pumas_state * states, const int * events, long n_bins,
    long per_bin, int event_mask, double * out, void * cuda_stream);

/* New: measured FP64 FMA throughput of the current device, in TFLOP/s (the
 * denominator of the roofline fraction reported by bench.py). */
PUMAS_API enum pumas_return pumas_b200_fp64_peak(double * tflops);
/* New. Self test of the division / square root of the shared memory resident
 * kernel (pb_hot_math.h) against the IEEE operations on n random operand
 * pairs with exponents over +-spread binades (<= 480): the numbers of results
 * whose bits differ, [0] divisions, [1] square roots. */
PUMAS_API enum pumas_return pumas_b200_hot_math_selftest(
    long n, unsigned long seed, int spread, unsigned long long * mismatches);

/* New: 1 when a CUDA device with compute capability 10.x is usable. */
PUMAS_API int pumas_b200_device_available(void);

#ifdef __cplusplus
}
#endif
#endif
