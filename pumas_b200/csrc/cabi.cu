/*
 * cabi.cu -- the C ABI of libpumas_b200.so (include/pumas_b200.h).
 *
 * Host side of the drop-in boundary: physics load / dump, contexts, property
 * getters, the error machinery of the reference (src/pumas.c:203-270,
 * 2299-2327, 7814-7859) and the transport entry points, which stage the
 * states through HBM and launch the kernels of kernels.cu. There is no CPU
 * transport path in this library: without a usable CUDA device the transport
 * entry points fail with PUMAS_RETURN_INTERNAL_ERROR.
 */
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"
#include "pb_coulomb.h"
#include "pb_dcs.h"
#include "pb_dcs_models.h"
#include "pb_dump.hpp"
#include "pb_tables.h"
#include "pumas_b200.h"

/* ---- opaque objects ------------------------------------------------------ */
/* A copy of the tables on one device. */
struct DeviceImage {
        int device;
        void * d_image;
        pb::Physics d_view;
};

struct pumas_physics {
        pb::HostPhysics host;
        /* Device copies of host.image, one per device, created on first use
         * (a host process may drive several GPUs with one physics object). */
        std::mutex mutex;
        std::vector<DeviceImage> images;
};

namespace {

/* ---- user registered DCS models (dcs_stack, pumas.c:15373-15460) -----------
 * A model is a name and a host function, as in the reference, plus -- for the
 * transport kernels -- its `__device__` counterpart
 * (pumas_b200_dcs_register_device). 64 entries in all, 8 of them built in. */
struct UserModel {
        int process;
        std::string name;
        pumas_dcs_t * host;
        pumas_dcs_t * device; /* a device function pointer, or NULL */
};
constexpr size_t USER_MODEL_STACK = 64 - 8;
const char * const PROCESS_NAME[3] = { "bremsstrahlung", "pair production",
        "photonuclear" };
std::vector<UserModel> g_user_models;
std::mutex g_user_models_mutex;

const UserModel * user_model_find(int process, const char * name)
{
        for (size_t i = 0; i < g_user_models.size(); i++)
                if ((g_user_models[i].process == process) &&
                    (g_user_models[i].name == name))
                        return &g_user_models[i];
        return nullptr;
}

/* struct simulation_context (pumas.c:426-477): the public struct comes first
 * so that `struct pumas_context *` is a pointer to this object. */
struct ContextImpl {
        struct pumas_context api;
        const pumas_physics * physics;
        unsigned long seed;
        int seed_set;
        unsigned long long track_offset;
        int has_geometry;
        struct pumas_b200_geometry geometry;
        /* `__device__` medium / locals callbacks (kernels_callback.cu); take
         * precedence over the descriptor */
        int has_callbacks;
        struct pumas_b200_callbacks callbacks;
        /* the descriptor behind the callback interface, for a physics with
         * user DCS models (those run on the device-linked kernel only) */
        int has_planar;
        struct pumas_b200_callbacks planar;
        cudaStream_t stream;
        /* device staging buffers (grown on demand) */
        long capacity;
        pb::State * d_states;
        int * d_events;
        int * d_media;
        unsigned int * d_random;
        unsigned int * d_steps;
        int * d_status;
        void * d_workspace;
        int kernel;
        long last_n;
        double last_ms;
        int last_launches;
        cudaEvent_t ev0, ev1;
        int events_ready;
        /* what `frame->medium` of a recorded frame points to: one object per
         * layer of the geometry (material set, no locals callback) */
        struct pumas_medium media_objects[PUMAS_B200_MAX_LAYERS];
        int device;  /* the device the buffers live on, -1: none yet */
        int pending; /* batches enqueued by transport_async, not waited for */
        int extra_memory;
        /* user memory follows */
        double data[1];
};

/* ---- host memory hooks, pumas.c:1368-1398 -----------------------------------
 * The library objects handed to the user (struct pumas_physics, struct
 * pumas_context and its user_data) are allocated with these; the C++
 * containers inside a physics object and the device memory are not. */
pumas_allocate_cb * g_allocate = &malloc;
pumas_reallocate_cb * g_reallocate = &realloc;
pumas_deallocate_cb * g_deallocate = &free;

/* ---- error machinery, pumas.c:203-270, 2299-2327, 7814-7859 --------------- */
void default_handler(
    enum pumas_return rc, pumas_function_t * caller, const char * message)
{
        (void)rc;
        (void)caller;
        fputs("pumas: library error. See details below\n", stderr);
        fprintf(stderr, "error: %s\n", message);
        exit(EXIT_FAILURE);
}

struct {
        pumas_handler_cb * handler = &default_handler;
        int catch_ = 0;
        int caught_code = PUMAS_RETURN_SUCCESS;
        pumas_function_t * caught_function = nullptr;
        char caught_message[1024];
} g_error;

#define FN(f) ((pumas_function_t *)(f))

enum pumas_return raise(pumas_function_t * caller, enum pumas_return rc,
    const char * file, int line, const char * fmt, ...)
{
        if (rc == PUMAS_RETURN_SUCCESS) return rc;
        char text[768];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(text, sizeof text, fmt, ap);
        va_end(ap);
        char message[1024];
        const char * name = pumas_error_function(caller);
        snprintf(message, sizeof message, "{ %s [#%d], %s:%d } %s",
            (name != nullptr) ? name : "?", (int)rc, file, line, text);

        if (g_error.catch_) {
                if (g_error.caught_code == PUMAS_RETURN_SUCCESS) {
                        g_error.caught_code = rc;
                        g_error.caught_function = caller;
                        snprintf(g_error.caught_message,
                            sizeof g_error.caught_message, "%s", message);
                }
                return rc;
        }
        if (g_error.handler != nullptr) g_error.handler(rc, caller, message);
        return rc;
}

#define RAISE(f, rc, ...) raise(FN(f), rc, __FILE__, __LINE__, __VA_ARGS__)

/* ---- device helpers -------------------------------------------------------- */
int device_usable(std::string * why)
{
        /* (cudaGetDeviceProperties costs milliseconds: asked once per device) */
        static int usable[64] = { 0 };
        {
                int dev = -1;
                if ((cudaGetDevice(&dev) == cudaSuccess) && (dev >= 0) &&
                    (dev < 64) && usable[dev])
                        return 1;
        }
        int count = 0;
        const cudaError_t e = cudaGetDeviceCount(&count);
        if ((e != cudaSuccess) || (count <= 0)) {
                cudaGetLastError();
                if (why)
                        *why = std::string("no CUDA device (") +
                            cudaGetErrorString(e) + ")";
                return 0;
        }
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
                cudaGetLastError();
                if (why) *why = "cannot query the CUDA device";
                return 0;
        }
        if (prop.major != 10) {
                if (why)
                        *why = std::string("device `") + prop.name +
                            "' is not an sm_100 (B200) part; this library "
                            "carries sm_100a code only";
                return 0;
        }
        if ((dev >= 0) && (dev < 64)) usable[dev] = 1;
        return 1;
}

enum pumas_return physics_upload(pumas_function_t * caller,
    pumas_physics * physics, pb::Physics * d_view)
{
        std::lock_guard<std::mutex> lock(physics->mutex);
        int dev = 0;
        cudaGetDevice(&dev);
        for (size_t i = 0; i < physics->images.size(); i++) {
                if (physics->images[i].device == dev) {
                        *d_view = physics->images[i].d_view;
                        return PUMAS_RETURN_SUCCESS;
                }
        }
        DeviceImage image;
        image.device = dev;
        image.d_image = nullptr;
        const size_t size = physics->host.image.size();
        cudaError_t e = cudaMalloc(&image.d_image, size);
        if (e == cudaSuccess)
                e = cudaMemcpy(image.d_image, physics->host.image.data(), size,
                    cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
                if (image.d_image != nullptr) cudaFree(image.d_image);
                return raise(caller, PUMAS_RETURN_MEMORY_ERROR, __FILE__,
                    __LINE__, "could not upload the physics tables (%s)",
                    cudaGetErrorString(e));
        }
        image.d_view = physics->host.view_at(image.d_image);
        physics->images.push_back(image);
        *d_view = image.d_view;
        return PUMAS_RETURN_SUCCESS;
}

void context_free_buffers(ContextImpl * c)
{
        if (c->d_states) cudaFree(c->d_states);
        if (c->d_events) cudaFree(c->d_events);
        if (c->d_media) cudaFree(c->d_media);
        if (c->d_random) cudaFree(c->d_random);
        if (c->d_steps) cudaFree(c->d_steps);
        c->d_states = nullptr;
        c->d_events = nullptr;
        c->d_media = nullptr;
        c->d_random = nullptr;
        c->d_steps = nullptr;
        c->capacity = 0;
}

cudaError_t context_reserve(ContextImpl * c, long n)
{
        cudaError_t e = cudaSuccess;
        if (c->d_status == nullptr) {
                e = cudaMalloc(&c->d_status, sizeof(int));
                if (e != cudaSuccess) return e;
        }
        if (c->d_workspace == nullptr) {
                size_t bytes = 0;
                e = pb::transport_workspace_bytes(&bytes);
                if (e != cudaSuccess) return e;
                e = cudaMalloc(&c->d_workspace, bytes);
                if (e != cudaSuccess) return e;
        }
        if (!c->events_ready) {
                e = cudaEventCreate(&c->ev0);
                if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
                if (e != cudaSuccess) return e;
                c->events_ready = 1;
        }
        if (n <= c->capacity) return cudaSuccess;
        context_free_buffers(c);
        e = cudaMalloc(&c->d_states, n * sizeof(pb::State));
        if (e == cudaSuccess) e = cudaMalloc(&c->d_events, n * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&c->d_media, 2 * n * sizeof(int));
        if (e == cudaSuccess)
                e = cudaMalloc(&c->d_random, n * sizeof(unsigned int));
        if (e == cudaSuccess)
                e = cudaMalloc(&c->d_steps, n * sizeof(unsigned int));
        if (e != cudaSuccess) {
                context_free_buffers(c);
                return e;
        }
        c->capacity = n;
        return cudaSuccess;
}

bool is_device_pointer(const void * ptr)
{
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
                cudaGetLastError();
                return false;
        }
        return (attr.type == cudaMemoryTypeDevice) ||
            (attr.type == cudaMemoryTypeManaged);
}

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

const char * return_text(int rc)
{
        switch (rc) {
        case PUMAS_RETURN_DENSITY_ERROR:
                return "negative density";
        case PUMAS_RETURN_DIRECTION_ERROR:
                return "bad norm for state direction";
        case PUMAS_RETURN_MISSING_LIMIT:
                return "infinite medium without external limit(s)";
        case PUMAS_RETURN_VALUE_ERROR: /* pumas.c:5133 */
                return "magnetic rotation is too strong";
        default:
                return "transport error";
        }
}

} /* namespace */

extern "C" {

/* ---- errors ---------------------------------------------------------------- */
void pumas_error_handler_set(pumas_handler_cb * handler)
{
        g_error.handler = handler;
}

pumas_handler_cb * pumas_error_handler_get(void) { return g_error.handler; }

void pumas_error_catch(int enable)
{
        if (enable) {
                g_error.catch_ = 1;
                g_error.caught_code = PUMAS_RETURN_SUCCESS;
                g_error.caught_function = nullptr;
        } else {
                g_error.catch_ = 0;
        }
}

enum pumas_return pumas_error_raise(void)
{
        if (g_error.catch_ == 0) {
                return RAISE(pumas_error_raise, PUMAS_RETURN_RAISE_ERROR,
                    "`pumas_error_raise' called with no catch");
        }
        g_error.catch_ = 0;
        const enum pumas_return rc = (enum pumas_return)g_error.caught_code;
        if ((rc != PUMAS_RETURN_SUCCESS) && (g_error.handler != nullptr))
                g_error.handler(
                    rc, g_error.caught_function, g_error.caught_message);
        g_error.caught_code = PUMAS_RETURN_SUCCESS;
        return rc;
}

const char * pumas_error_function(pumas_function_t * caller)
{
#define NAME(f)                                                                \
        if (caller == FN(f)) return #f;
        NAME(pumas_physics_load)
        NAME(pumas_physics_dump)
        NAME(pumas_physics_create)
        NAME(pumas_physics_particle)
        NAME(pumas_physics_property_range)
        NAME(pumas_physics_property_proper_time)
        NAME(pumas_physics_property_magnetic_rotation)
        NAME(pumas_physics_property_kinetic_energy)
        NAME(pumas_physics_property_stopping_power)
        NAME(pumas_physics_property_energy_straggling)
        NAME(pumas_physics_property_elastic_cutoff_angle)
        NAME(pumas_physics_property_elastic_path)
        NAME(pumas_physics_property_transport_path)
        NAME(pumas_physics_property_cross_section)
        NAME(pumas_physics_element_properties)
        NAME(pumas_physics_element_name)
        NAME(pumas_physics_element_index)
        NAME(pumas_physics_material_name)
        NAME(pumas_physics_material_index)
        NAME(pumas_physics_material_properties)
        NAME(pumas_physics_composite_properties)
        NAME(pumas_physics_table_value)
        NAME(pumas_physics_table_index)
        NAME(pumas_context_create)
        NAME(pumas_context_random_seed_set)
        NAME(pumas_context_random_seed_get)
        NAME(pumas_context_random_dump)
        NAME(pumas_context_random_load)
        NAME(pumas_context_transport)
        NAME(pumas_context_transport_n)
        NAME(pumas_recorder_create)
        NAME(pumas_recorder_clear)
        NAME(pumas_recorder_destroy)
        NAME(pumas_b200_recorder_export)
        NAME(pumas_b200_context_transport_async)
        NAME(pumas_b200_context_synchronize)
        NAME(pumas_b200_context_geometry_set)
        NAME(pumas_b200_context_callbacks_set)
        NAME(pumas_b200_planar_callbacks)
        NAME(pumas_b200_context_track_offset_set)
        NAME(pumas_b200_context_stream_set)
        NAME(pumas_b200_context_kernel_set)
        NAME(pumas_b200_context_last_counters)
        NAME(pumas_b200_context_last_kernel_time)
        NAME(pumas_error_raise)
        NAME(pumas_constant)
        NAME(pumas_physics_dcs)
        NAME(pumas_dcs_get)
        NAME(pumas_dcs_register)
        NAME(pumas_b200_dcs_register_device)
        NAME(pumas_dcs_range)
#undef NAME
        return nullptr;
}

/* pumas.c:1379-1398: NULL restores the C library function. */
void pumas_memory_allocator(pumas_allocate_cb * allocator)
{
        g_allocate = (allocator == nullptr) ? &malloc : allocator;
}

void pumas_memory_reallocator(pumas_reallocate_cb * reallocator)
{
        g_reallocate = (reallocator == nullptr) ? &realloc : reallocator;
}

void pumas_memory_deallocator(pumas_deallocate_cb * deallocator)
{
        g_deallocate = (deallocator == nullptr) ? &free : deallocator;
}

void pumas_version(int * major, int * minor, int * patch)
{
        if (major != nullptr) *major = PUMAS_VERSION_MAJOR;
        if (minor != nullptr) *minor = PUMAS_VERSION_MINOR;
        if (patch != nullptr) *patch = PUMAS_VERSION_PATCH;
}

/* pumas.c:13673 */
enum pumas_return pumas_constant(enum pumas_constant index, double * value)
{
        static const double values[PUMAS_N_CONSTANTS] = { pb::ALPHA_EM,
                pb::AVOGADRO_NUMBER, pb::BOHR_RADIUS, pb::ELECTRON_MASS,
                pb::ELECTRON_RADIUS, pb::HBAR_C, pb::MUON_C_TAU, pb::MUON_MASS,
                pb::NEUTRON_MASS, pb::PION_MASS, pb::PROTON_MASS,
                pb::TAU_C_TAU, pb::TAU_MASS };
        if (value == nullptr)
                return RAISE(pumas_constant, PUMAS_RETURN_VALUE_ERROR,
                    "a NULL pointer was provided");
        if (((int)index < 0) || (index >= PUMAS_N_CONSTANTS))
                return RAISE(pumas_constant, PUMAS_RETURN_INDEX_ERROR,
                    "invalid `constant' index [%d]", (int)index);
        *value = values[index];
        return PUMAS_RETURN_SUCCESS;
}

int pumas_b200_device_available(void) { return device_usable(nullptr); }

/* ---- physics --------------------------------------------------------------- */
} /* extern "C" */
namespace {
void physics_delete(pumas_physics * physics)
{
        physics->~pumas_physics();
        g_deallocate(physics);
}
} /* namespace */
extern "C" {

enum pumas_return pumas_physics_load(
    struct pumas_physics ** physics_ptr, FILE * stream)
{
        if (physics_ptr == nullptr)
                return RAISE(pumas_physics_load, PUMAS_RETURN_PHYSICS_ERROR,
                    "a NULL physics pointer was provided");
        *physics_ptr = nullptr;
        if (stream == nullptr)
                return RAISE(pumas_physics_load, PUMAS_RETURN_PATH_ERROR,
                    "invalid input stream (null)");
        void * const memory = g_allocate(sizeof(pumas_physics));
        pumas_physics * physics =
            (memory != nullptr) ? new (memory) pumas_physics : nullptr;
        if (physics == nullptr)
                return RAISE(pumas_physics_load, PUMAS_RETURN_MEMORY_ERROR,
                    "could not allocate memory");
        std::string err;
        const int rc = pb::load_physics_stream(stream, physics->host, err);
        if (rc != 0) {
                physics_delete(physics);
                return RAISE(pumas_physics_load, (enum pumas_return)rc, "%s",
                    err.c_str());
        }
        /* models the dump names that are not built in: the user's, which must
         * be registered by now (dcs_get, pumas.c:10954; the reference fails
         * the same way) */
        for (int i = 0; i < 3; i++) {
                if (physics->host.view.model[i] != pb::MODEL_USER) continue;
                std::lock_guard<std::mutex> lock(g_user_models_mutex);
                const UserModel * m =
                    user_model_find(i, physics->host.model[i].c_str());
                if (m == nullptr) {
                        const std::string name = physics->host.model[i];
                        physics_delete(physics);
                        return RAISE(pumas_physics_load,
                            PUMAS_RETURN_MODEL_ERROR,
                            "cannot find %s model for %s process",
                            name.c_str(), PROCESS_NAME[i]);
                }
                physics->host.view.user_dcs[i] = m->host;
        }
        if (physics->host.view.max_components > pb::MAX_COMPONENTS) {
                physics_delete(physics);
                return RAISE(pumas_physics_load, PUMAS_RETURN_MEMORY_ERROR,
                    "materials with more than %d elements are not supported",
                    pb::MAX_COMPONENTS);
        }
        *physics_ptr = physics;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_dump(
    const struct pumas_physics * physics, FILE * stream)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_dump, PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        if (stream == nullptr)
                return RAISE(pumas_physics_dump, PUMAS_RETURN_PATH_ERROR,
                    "invalid output stream (null)");
        const std::vector<unsigned char> & raw = physics->host.raw;
        if (fwrite(raw.data(), raw.size(), 1, stream) != 1)
                return RAISE(pumas_physics_dump, PUMAS_RETURN_IO_ERROR,
                    "could not write to dump file");
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_create(struct pumas_physics ** physics,
    enum pumas_particle particle, const char * mdf_path,
    const char * dedx_path, const struct pumas_physics_settings * settings)
{
        (void)particle;
        (void)mdf_path;
        (void)dedx_path;
        (void)settings;
        if (physics != nullptr) *physics = nullptr;
        return RAISE(pumas_physics_create, PUMAS_RETURN_MODEL_ERROR,
            "table generation from a MDF is not part of this library: create "
            "the physics with the reference and load its dump with "
            "pumas_physics_load (see INTEGRATION.md)");
}

void pumas_physics_destroy(struct pumas_physics ** physics_ptr)
{
        if ((physics_ptr == nullptr) || (*physics_ptr == nullptr)) return;
        pumas_physics * physics = *physics_ptr;
        for (size_t i = 0; i < physics->images.size(); i++) {
                /* (freed on the device they were allocated on) */
                int dev = 0;
                cudaGetDevice(&dev);
                cudaSetDevice(physics->images[i].device);
                cudaFree(physics->images[i].d_image);
                cudaSetDevice(dev);
        }
        physics_delete(physics);
        *physics_ptr = nullptr;
}

double pumas_physics_cutoff(const struct pumas_physics * physics)
{
        return (physics == nullptr) ? -1 : physics->host.view.cutoff;
}

double pumas_physics_elastic_ratio(const struct pumas_physics * physics)
{
        return (physics == nullptr) ? -1 : physics->host.view.elastic_ratio;
}

enum pumas_return pumas_physics_particle(const struct pumas_physics * physics,
    enum pumas_particle * particle, double * lifetime, double * mass)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_particle,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        const pb::Physics & p = physics->host.view;
        if (particle != nullptr) *particle = (enum pumas_particle)p.particle;
        if (lifetime != nullptr) *lifetime = p.ctau;
        if (mass != nullptr) *mass = p.mass;
        return PUMAS_RETURN_SUCCESS;
}

/* Property getters, pumas.c:2854-3111. */
#define PROPERTY_PROLOGUE(f, with_scheme)                                      \
        *value = 0.;                                                           \
        if (physics == nullptr)                                                \
                return RAISE(f, PUMAS_RETURN_PHYSICS_ERROR,                    \
                    "the Physics has not been initialised");                   \
        const pb::Physics & p = physics->host.view;                            \
        if (with_scheme &&                                                     \
            (((int)scheme <= PUMAS_MODE_DISABLED) ||                           \
                ((int)scheme >= PUMAS_MODE_STRAGGLED)))                        \
                return RAISE(f, PUMAS_RETURN_INDEX_ERROR,                      \
                    "invalid energy loss scheme [%d]", (int)scheme);           \
        if ((material < 0) || (material >= p.n_materials))                     \
                return RAISE(f, PUMAS_RETURN_INDEX_ERROR,                      \
                    "invalid material index [%d]", material);

enum pumas_return pumas_physics_property_range(
    const struct pumas_physics * physics, enum pumas_mode scheme, int material,
    double kinetic, double * value)
{
        PROPERTY_PROLOGUE(pumas_physics_property_range, 1)
        *value = pb::cel_grammage(p, (int)scheme, material, kinetic);
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_property_proper_time(
    const struct pumas_physics * physics, enum pumas_mode scheme, int material,
    double kinetic, double * value)
{
        PROPERTY_PROLOGUE(pumas_physics_property_proper_time, 1)
        *value = pb::cel_proper_time(p, (int)scheme, material, kinetic);
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_property_magnetic_rotation(
    const struct pumas_physics * physics, int material, double kinetic,
    double * value)
{
        const int scheme = 0;
        PROPERTY_PROLOGUE(pumas_physics_property_magnetic_rotation, 0)
        (void)scheme;
        *value = pb::cel_magnetic_rotation(p, material, kinetic);
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_property_kinetic_energy(
    const struct pumas_physics * physics, enum pumas_mode scheme, int material,
    double grammage, double * value)
{
        PROPERTY_PROLOGUE(pumas_physics_property_kinetic_energy, 1)
        *value = pb::cel_kinetic_energy(p, (int)scheme, material, grammage);
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_property_stopping_power(
    const struct pumas_physics * physics, enum pumas_mode scheme, int material,
    double kinetic, double * value)
{
        PROPERTY_PROLOGUE(pumas_physics_property_stopping_power, 1)
        *value = pb::cel_energy_loss(p, (int)scheme, material, kinetic);
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_property_energy_straggling(
    const struct pumas_physics * physics, int material, double kinetic,
    double * value)
{
        const int scheme = 0;
        PROPERTY_PROLOGUE(pumas_physics_property_energy_straggling, 0)
        (void)scheme;
        *value = pb::cel_straggling(p, material, kinetic);
        return PUMAS_RETURN_SUCCESS;
}

/* pumas.c:2967-3009: cut-off angle in radians. */
enum pumas_return pumas_physics_property_elastic_cutoff_angle(
    const struct pumas_physics * physics, int material, double kinetic,
    double * value)
{
        const int scheme = 0;
        PROPERTY_PROLOGUE(pumas_physics_property_elastic_cutoff_angle, 0)
        (void)scheme;
        double mu0 = 0., invlb1 = 0.;
        pb::table_get_msc(p, PUMAS_MODE_STRAGGLED, material, kinetic, &mu0,
            &invlb1);
        if (mu0 < 1E-08) {
                *value = 2 * sqrt(mu0);
        } else {
                *value = acos(1. - 2 * mu0);
        }
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_property_elastic_path(
    const struct pumas_physics * physics, int material, double kinetic,
    double * value)
{
        const int scheme = 0;
        PROPERTY_PROLOGUE(pumas_physics_property_elastic_path, 0)
        (void)scheme;
        *value = pb::coulomb_ehs_length(p, material, kinetic);
        return PUMAS_RETURN_SUCCESS;
}

/* pumas.c:3050-3082: the scheme may be DISABLED here (pure elastic part). */
enum pumas_return pumas_physics_property_transport_path(
    const struct pumas_physics * physics, enum pumas_mode scheme, int material,
    double kinetic, double * value)
{
        *value = 0.;
        if (physics == nullptr)
                return RAISE(pumas_physics_property_transport_path,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        const pb::Physics & p = physics->host.view;
        if (((int)scheme < PUMAS_MODE_DISABLED) ||
            ((int)scheme >= PUMAS_MODE_STRAGGLED))
                return RAISE(pumas_physics_property_transport_path,
                    PUMAS_RETURN_INDEX_ERROR,
                    "invalid energy loss scheme [%d]", (int)scheme);
        if ((material < 0) || (material >= p.n_materials))
                return RAISE(pumas_physics_property_transport_path,
                    PUMAS_RETURN_INDEX_ERROR, "invalid material index [%d]",
                    material);
        double mu0 = 0., invlb1 = 0.;
        pb::table_get_msc(p, (int)scheme, material, kinetic, &mu0, &invlb1);
        *value = (invlb1 > 0.) ? 1. / invlb1 : DBL_MAX;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_property_cross_section(
    const struct pumas_physics * physics, int material, double kinetic,
    double * value)
{
        const int scheme = 0;
        PROPERTY_PROLOGUE(pumas_physics_property_cross_section, 0)
        (void)scheme;
        *value = (kinetic < p.Kt[material]) ?
            0. :
            pb::del_cross_section(p, material, kinetic);
        return PUMAS_RETURN_SUCCESS;
}

/* Elements and materials, pumas.c:3270-3400. */
int pumas_physics_element_length(const struct pumas_physics * physics)
{
        return (physics == nullptr) ? 0 : physics->host.view.n_elements;
}

enum pumas_return pumas_physics_element_properties(
    const struct pumas_physics * physics, int index, double * Z, double * A,
    double * I)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_element_properties,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        const pb::Physics & p = physics->host.view;
        if ((index < 0) || (index >= p.n_elements))
                return RAISE(pumas_physics_element_properties,
                    PUMAS_RETURN_INDEX_ERROR, "invalid element index [%d]",
                    index);
        if (Z != nullptr) *Z = p.element[index].Z;
        if (A != nullptr) *A = p.element[index].A;
        if (I != nullptr) *I = p.element[index].I;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_element_name(
    const struct pumas_physics * physics, int index, const char ** element)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_element_name,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        if ((index < 0) || (index >= physics->host.view.n_elements))
                return RAISE(pumas_physics_element_name,
                    PUMAS_RETURN_INDEX_ERROR, "invalid element index [%d]",
                    index);
        *element = physics->host.element_name[index].c_str();
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_element_index(
    const struct pumas_physics * physics, const char * element, int * index)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_element_index,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        for (size_t i = 0; i < physics->host.element_name.size(); i++)
                if (physics->host.element_name[i] == element) {
                        *index = (int)i;
                        return PUMAS_RETURN_SUCCESS;
                }
        return RAISE(pumas_physics_element_index, PUMAS_RETURN_UNKNOWN_ELEMENT,
            "unknown element `%s'", element);
}

enum pumas_return pumas_physics_material_name(
    const struct pumas_physics * physics, int index, const char ** material)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_material_name,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        if ((index < 0) || (index >= physics->host.view.n_materials))
                return RAISE(pumas_physics_material_name,
                    PUMAS_RETURN_INDEX_ERROR, "invalid material index [%d]",
                    index);
        *material = physics->host.material_name[index].c_str();
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_material_index(
    const struct pumas_physics * physics, const char * material, int * index)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_material_index,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        for (size_t i = 0; i < physics->host.material_name.size(); i++)
                if (physics->host.material_name[i] == material) {
                        *index = (int)i;
                        return PUMAS_RETURN_SUCCESS;
                }
        return RAISE(pumas_physics_material_index,
            PUMAS_RETURN_UNKNOWN_MATERIAL, "unknown material `%s'", material);
}

int pumas_physics_material_length(const struct pumas_physics * physics)
{
        return (physics == nullptr) ? 0 : physics->host.view.n_materials;
}

int pumas_physics_composite_length(const struct pumas_physics * physics)
{
        return (physics == nullptr) ? 0 : physics->host.view.n_composites;
}

/* pumas.c:3446-3475 */
enum pumas_return pumas_physics_composite_properties(
    const struct pumas_physics * physics, int material, int * length,
    int * components, double * fractions)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_composite_properties,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        const pb::Physics & p = physics->host.view;
        const int i0 = p.n_materials - p.n_composites;
        if ((material < i0) || (material > p.n_materials - 1) ||
            ((size_t)(material - i0) >= physics->host.composite.size()))
                return RAISE(pumas_physics_composite_properties,
                    PUMAS_RETURN_INDEX_ERROR, "invalid material index [%d]",
                    material);
        const std::vector<std::pair<int, double> > & parts =
            physics->host.composite[material - i0];
        if (length != nullptr) *length = (int)parts.size();
        for (size_t i = 0; i < parts.size(); i++) {
                if (components != nullptr) components[i] = parts[i].first;
                if (fractions != nullptr) fractions[i] = parts[i].second;
        }
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_physics_material_properties(
    const struct pumas_physics * physics, int index, int * length,
    double * density, double * I, int * components, double * fractions)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_material_properties,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        const pb::Physics & p = physics->host.view;
        if ((index < 0) || (index >= p.n_materials))
                return RAISE(pumas_physics_material_properties,
                    PUMAS_RETURN_INDEX_ERROR, "invalid material index [%d]",
                    index);
        if (length != nullptr) *length = p.elements_in[index];
        if (density != nullptr) *density = p.material_density[index];
        if (I != nullptr) *I = physics->host.material_I[index];
        for (int i = 0; i < p.elements_in[index]; i++) {
                const int k = p.comp_offset[index] + i;
                if (components != nullptr) components[i] = p.comp_element[k];
                if (fractions != nullptr) fractions[i] = p.comp_fraction[k];
        }
        return PUMAS_RETURN_SUCCESS;
}

/* Table rows, pumas.c:3486-3625. */
int pumas_physics_table_length(const struct pumas_physics * physics)
{
        return (physics == nullptr) ? 0 : physics->host.view.n_energies;
}

enum pumas_return pumas_physics_table_value(
    const struct pumas_physics * physics, enum pumas_property property,
    enum pumas_mode scheme, int material, int row, double * value)
{
        *value = 0.;
        if (physics == nullptr)
                return RAISE(pumas_physics_table_value,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        const pb::Physics & p = physics->host.view;
        if ((material < 0) || (material >= p.n_materials))
                return RAISE(pumas_physics_table_value,
                    PUMAS_RETURN_INDEX_ERROR, "invalid material index [%d]",
                    material);
        if ((row < 0) || (row >= p.n_energies))
                return RAISE(pumas_physics_table_value,
                    PUMAS_RETURN_INDEX_ERROR, "invalid `row' index [%d]", row);
        const int s = (int)scheme;
        const bool scheme_ok =
            (s > PUMAS_MODE_DISABLED) && (s < PUMAS_MODE_STRAGGLED);
        switch (property) {
        case PUMAS_PROPERTY_KINETIC_ENERGY:
                *value = p.K[row];
                return PUMAS_RETURN_SUCCESS;
        case PUMAS_PROPERTY_RANGE:
                if (!scheme_ok) break;
                *value = pb::col_sm(p, p.X, s, material)[row];
                return PUMAS_RETURN_SUCCESS;
        case PUMAS_PROPERTY_PROPER_TIME:
                if (!scheme_ok) break;
                *value = pb::col_sm(p, p.T, s, material)[row];
                return PUMAS_RETURN_SUCCESS;
        case PUMAS_PROPERTY_STOPPING_POWER:
                if (!scheme_ok) break;
                *value = pb::col_sm(p, p.dE, s, material)[row];
                return PUMAS_RETURN_SUCCESS;
        case PUMAS_PROPERTY_MAGNETIC_ROTATION:
                *value = pb::cel_magnetic_rotation(p, material, p.K[row]);
                return PUMAS_RETURN_SUCCESS;
        case PUMAS_PROPERTY_CROSS_SECTION:
                *value = pb::col_m(p, p.CS, material)[row];
                return PUMAS_RETURN_SUCCESS;
        default:
                return RAISE(pumas_physics_table_value,
                    PUMAS_RETURN_INDEX_ERROR, "invalid `property' index [%d]",
                    (int)property);
        }
        return RAISE(pumas_physics_table_value, PUMAS_RETURN_INDEX_ERROR,
            "invalid energy loss scheme [%d]", s);
}

enum pumas_return pumas_physics_table_index(
    const struct pumas_physics * physics, enum pumas_property property,
    enum pumas_mode scheme, int material, double value, int * index)
{
        *index = 0;
        if (physics == nullptr)
                return RAISE(pumas_physics_table_index,
                    PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        const pb::Physics & p = physics->host.view;
        if ((material < 0) || (material >= p.n_materials))
                return RAISE(pumas_physics_table_index,
                    PUMAS_RETURN_INDEX_ERROR, "invalid material index [%d]",
                    material);
        const int s = (int)scheme;
        const bool scheme_ok =
            (s > PUMAS_MODE_DISABLED) && (s < PUMAS_MODE_STRAGGLED);
        const double * table = nullptr;
        if (property == PUMAS_PROPERTY_KINETIC_ENERGY)
                table = p.K;
        else if (property == PUMAS_PROPERTY_RANGE) {
                if (scheme_ok) table = pb::col_sm(p, p.X, s, material);
        } else if (property == PUMAS_PROPERTY_PROPER_TIME) {
                if (scheme_ok) table = pb::col_sm(p, p.T, s, material);
        } else if (property == PUMAS_PROPERTY_MAGNETIC_ROTATION) {
                /* pumas.c:3596-3607: decreasing function of the energy. */
                const int imax = p.n_energies - 1;
                const double * T = pb::col_sm(p, p.T, PUMAS_MODE_CSDA, material);
                const double factor = pb::LARMOR_FACTOR / p.mass;
                const double t = T[imax] - value / factor;
                if (value < 0.)
                        return RAISE(pumas_physics_table_index,
                            PUMAS_RETURN_VALUE_ERROR,
                            "out of range `value' [%g]", value);
                const int i = pb::table_index(p, T, t);
                if ((i < 0) || (i >= imax))
                        return RAISE(pumas_physics_table_index,
                            PUMAS_RETURN_VALUE_ERROR,
                            "out of range `value' [%g]", value);
                *index = i;
                return PUMAS_RETURN_SUCCESS;
        } else {
                return RAISE(pumas_physics_table_index,
                    PUMAS_RETURN_INDEX_ERROR, "invalid `property' index [%d]",
                    (int)property);
        }
        if (table == nullptr)
                return RAISE(pumas_physics_table_index,
                    PUMAS_RETURN_INDEX_ERROR,
                    "invalid energy loss scheme [%d]", s);
        const int imax = p.n_energies - 1;
        if (value < table[0])
                return RAISE(pumas_physics_table_index,
                    PUMAS_RETURN_VALUE_ERROR, "out of range `value' [%g]",
                    value);
        if (value == table[imax]) {
                *index = imax;
                return PUMAS_RETURN_SUCCESS;
        }
        if (value > table[imax])
                return RAISE(pumas_physics_table_index,
                    PUMAS_RETURN_VALUE_ERROR, "out of range `value' [%g]",
                    value);
        *index = pb::table_index(p, table, value);
        return PUMAS_RETURN_SUCCESS;
}

/* ---- context --------------------------------------------------------------- */
enum pumas_return pumas_context_create(struct pumas_context ** context_,
    const struct pumas_physics * physics, int extra_memory)
{
        if (context_ == nullptr)
                return RAISE(pumas_context_create, PUMAS_RETURN_VALUE_ERROR,
                    "a NULL context pointer was provided");
        *context_ = nullptr;
        if (physics == nullptr)
                return RAISE(pumas_context_create, PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        if (extra_memory < 0) extra_memory = 0;
        const size_t size = sizeof(ContextImpl) + (size_t)extra_memory + 16;
        ContextImpl * c = static_cast<ContextImpl *>(g_allocate(size));
        if (c == nullptr)
                return RAISE(pumas_context_create, PUMAS_RETURN_MEMORY_ERROR,
                    "could not allocate memory");
        memset(c, 0x0, size);
        c->physics = physics;
        c->device = -1;
        c->extra_memory = extra_memory;
        c->api.user_data = (extra_memory > 0) ? c->data : nullptr;
        c->api.medium = nullptr;
        c->api.random = nullptr;
        c->api.recorder = nullptr;
        /* Defaults, pumas.c:2592-2605. */
        c->api.mode.decay =
            (physics->host.view.particle == PUMAS_PARTICLE_MUON) ?
            PUMAS_MODE_WEIGHTED :
            PUMAS_MODE_RANDOMISED;
        c->api.mode.direction = PUMAS_MODE_FORWARD;
        c->api.mode.energy_loss = PUMAS_MODE_STRAGGLED;
        c->api.mode.scattering = PUMAS_MODE_MIXED;
        c->api.event = PUMAS_EVENT_NONE;
        c->api.accuracy = 1E-02;
        *context_ = &c->api;
        return PUMAS_RETURN_SUCCESS;
}

void pumas_context_destroy(struct pumas_context ** context)
{
        if ((context == nullptr) || (*context == nullptr)) return;
        ContextImpl * c = reinterpret_cast<ContextImpl *>(*context);
        context_free_buffers(c);
        if (c->has_planar) cudaFree(const_cast<void *>(c->planar.user_data));
        if (c->d_status) cudaFree(c->d_status);
        if (c->d_workspace) cudaFree(c->d_workspace);
        if (c->events_ready) {
                cudaEventDestroy(c->ev0);
                cudaEventDestroy(c->ev1);
        }
        g_deallocate(c);
        *context = nullptr;
}

const struct pumas_physics * pumas_context_physics_get(
    const struct pumas_context * context)
{
        if (context == nullptr) return nullptr;
        return reinterpret_cast<const ContextImpl *>(context)->physics;
}

enum pumas_return pumas_context_random_seed_set(
    struct pumas_context * context, const unsigned long * seed_ptr)
{
        if (context == nullptr)
                return RAISE(pumas_context_random_seed_set,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        if (seed_ptr == nullptr) {
                /* pumas.c:2372-2388: seed from the OS entropy pool. */
                unsigned long seed = 0;
                FILE * fp = fopen("/dev/urandom", "rb");
                if ((fp == nullptr) || (fread(&seed, sizeof seed, 1, fp) != 1)) {
                        if (fp != nullptr) fclose(fp);
                        return RAISE(pumas_context_random_seed_set,
                            PUMAS_RETURN_PATH_ERROR,
                            "could not read from `/dev/urandom'");
                }
                fclose(fp);
                c->seed = seed;
        } else {
                c->seed = *seed_ptr;
        }
        c->seed_set = 1;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_context_random_seed_get(
    struct pumas_context * context, unsigned long * seed_ptr)
{
        if (context == nullptr)
                return RAISE(pumas_context_random_seed_get,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        if (!c->seed_set) {
                const enum pumas_return rc =
                    pumas_context_random_seed_set(context, nullptr);
                if (rc != PUMAS_RETURN_SUCCESS) return rc;
        }
        *seed_ptr = c->seed;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_b200_context_geometry_set(
    struct pumas_context * context, const struct pumas_b200_geometry * geometry)
{
        if ((context == nullptr) || (geometry == nullptr))
                return RAISE(pumas_b200_context_geometry_set,
                    PUMAS_RETURN_VALUE_ERROR, "a NULL pointer was provided");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        const pb::Physics & p = c->physics->host.view;
        const int n_layers =
            (geometry->kind == PUMAS_B200_GEOMETRY_UNIFORM) ? 1 :
                                                              geometry->n_layers;
        for (int k = 0; (k < n_layers) && (k < PUMAS_B200_MAX_LAYERS); k++) {
                c->media_objects[k].material = geometry->layer[k].material;
                c->media_objects[k].locals = nullptr;
        }
        if ((geometry->kind != PUMAS_B200_GEOMETRY_UNIFORM) &&
            (geometry->kind != PUMAS_B200_GEOMETRY_LAYERS))
                return RAISE(pumas_b200_context_geometry_set,
                    PUMAS_RETURN_MEDIUM_ERROR, "invalid geometry kind [%d]",
                    geometry->kind);
        if ((n_layers < 1) || (n_layers > PUMAS_B200_MAX_LAYERS))
                return RAISE(pumas_b200_context_geometry_set,
                    PUMAS_RETURN_MEDIUM_ERROR, "invalid number of layers [%d]",
                    n_layers);
        for (int i = 0; i < n_layers; i++) {
                const int m = geometry->layer[i].material;
                if ((m < 0) || (m >= p.n_materials))
                        return RAISE(pumas_b200_context_geometry_set,
                            PUMAS_RETURN_INDEX_ERROR,
                            "invalid material index [%d]", m);
                if ((geometry->kind == PUMAS_B200_GEOMETRY_LAYERS) &&
                    !(geometry->z[i] < geometry->z[i + 1]))
                        return RAISE(pumas_b200_context_geometry_set,
                            PUMAS_RETURN_MEDIUM_ERROR,
                            "layer boundaries must be increasing");
                if ((geometry->layer[i].density_model ==
                        PUMAS_B200_DENSITY_EXPONENTIAL) &&
                    !(geometry->layer[i].scale != 0.))
                        return RAISE(pumas_b200_context_geometry_set,
                            PUMAS_RETURN_DENSITY_ERROR,
                            "null scale height for layer %d", i);
        }
        c->geometry = *geometry;
        c->has_geometry = 1;
        if (c->has_planar) {
                cudaFree(const_cast<void *>(c->planar.user_data));
                c->has_planar = 0;
        }
        return PUMAS_RETURN_SUCCESS;
}

/* Device callbacks (pumas.h:509-511, 348-349 on the device): see
 * pumas_b200_geometry.h and kernels_callback.cu. NULL: back to the descriptor. */
enum pumas_return pumas_b200_context_callbacks_set(struct pumas_context * context,
    const struct pumas_b200_callbacks * callbacks)
{
        if (context == nullptr)
                return RAISE(pumas_b200_context_callbacks_set,
                    PUMAS_RETURN_VALUE_ERROR, "a NULL pointer was provided");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        if (callbacks == nullptr) {
                c->has_callbacks = 0;
                return PUMAS_RETURN_SUCCESS;
        }
        const pb::Physics & p = c->physics->host.view;
        if (callbacks->medium == nullptr)
                return RAISE(pumas_b200_context_callbacks_set,
                    PUMAS_RETURN_MEDIUM_ERROR, "no medium function");
        if ((callbacks->n_media < 1) ||
            (callbacks->n_media > PUMAS_B200_MAX_LAYERS))
                return RAISE(pumas_b200_context_callbacks_set,
                    PUMAS_RETURN_MEDIUM_ERROR, "invalid number of media [%d]",
                    callbacks->n_media);
        for (int i = 0; i < callbacks->n_media; i++) {
                const int m = callbacks->material[i];
                if ((m < 0) || (m >= p.n_materials))
                        return RAISE(pumas_b200_context_callbacks_set,
                            PUMAS_RETURN_INDEX_ERROR,
                            "invalid material index [%d]", m);
                if (callbacks->has_locals[i] && (callbacks->locals == nullptr))
                        return RAISE(pumas_b200_context_callbacks_set,
                            PUMAS_RETURN_MEDIUM_ERROR,
                            "medium %d has locals but there is no locals "
                            "function", i);
        }
        c->callbacks = *callbacks;
        c->has_callbacks = 1;
        return PUMAS_RETURN_SUCCESS;
}

/* The library's own client of the callback interface: the planar media of a
 * descriptor as device functions, user_data = a device copy of `geometry`. */
enum pumas_return pumas_b200_planar_callbacks(
    const struct pumas_b200_geometry * geometry,
    struct pumas_b200_callbacks * callbacks)
{
        if ((geometry == nullptr) || (callbacks == nullptr))
                return RAISE(pumas_b200_planar_callbacks,
                    PUMAS_RETURN_VALUE_ERROR, "a NULL pointer was provided");
        if (((geometry->kind != PUMAS_B200_GEOMETRY_UNIFORM) &&
                (geometry->kind != PUMAS_B200_GEOMETRY_LAYERS)) ||
            (geometry->n_layers < 1) ||
            (geometry->n_layers > PUMAS_B200_MAX_LAYERS))
                return RAISE(pumas_b200_planar_callbacks,
                    PUMAS_RETURN_MEDIUM_ERROR, "invalid geometry");
        std::string why;
        if (!device_usable(&why))
                return RAISE(pumas_b200_planar_callbacks,
                    PUMAS_RETURN_INTERNAL_ERROR, "%s", why.c_str());
        const cudaError_t e = pb::planar_callbacks(*geometry, callbacks);
        if (e != cudaSuccess)
                return RAISE(pumas_b200_planar_callbacks,
                    PUMAS_RETURN_INTERNAL_ERROR, "CUDA error (%s)",
                    cudaGetErrorString(e));
        return PUMAS_RETURN_SUCCESS;
}

void pumas_b200_planar_callbacks_free(struct pumas_b200_callbacks * callbacks)
{
        if ((callbacks == nullptr) || (callbacks->user_data == nullptr)) return;
        cudaFree(const_cast<void *>(callbacks->user_data));
        callbacks->user_data = nullptr;
}

/* pumas.c:2415-2492. The state of the random streams of a context is its seed
 * and the id of its next first track (Philox is counter based): a tag, then
 * the two 64 bit words. */
static const int RANDOM_DUMP_TAG = 0xB200;

enum pumas_return pumas_context_random_dump(
    struct pumas_context * context, FILE * stream)
{
        if (context == nullptr)
                return RAISE(pumas_context_random_dump,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        if (stream == nullptr)
                return RAISE(pumas_context_random_dump,
                    PUMAS_RETURN_PATH_ERROR, "invalid input stream (null)");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        if (!c->seed_set) {
                /* pumas.c:2432-2438: initialise the stream first */
                const enum pumas_return rc =
                    pumas_context_random_seed_set(context, nullptr);
                if (rc != PUMAS_RETURN_SUCCESS) return rc;
        }
        const unsigned long long words[2] = {
                (unsigned long long)c->seed,
                (unsigned long long)c->track_offset };
        if ((fwrite(&RANDOM_DUMP_TAG, sizeof RANDOM_DUMP_TAG, 1, stream) !=
                1) ||
            (fwrite(words, sizeof words, 1, stream) != 1))
                return RAISE(pumas_context_random_dump, PUMAS_RETURN_IO_ERROR,
                    "could not not write to stream");
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_context_random_load(
    struct pumas_context * context, FILE * stream)
{
        if (context == nullptr)
                return RAISE(pumas_context_random_load,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        if (stream == nullptr)
                return RAISE(pumas_context_random_load,
                    PUMAS_RETURN_PATH_ERROR, "invalid input stream (null)");
        int tag;
        unsigned long long words[2];
        if (fread(&tag, sizeof tag, 1, stream) != 1)
                return RAISE(pumas_context_random_load, PUMAS_RETURN_IO_ERROR,
                    "could not read from stream");
        if (tag != RANDOM_DUMP_TAG)
                return RAISE(pumas_context_random_load,
                    PUMAS_RETURN_FORMAT_ERROR,
                    "incompatible version of binary dump");
        if (fread(words, sizeof words, 1, stream) != 1)
                return RAISE(pumas_context_random_load, PUMAS_RETURN_IO_ERROR,
                    "could not read from stream");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        c->seed = (unsigned long)words[0];
        c->seed_set = 1;
        c->track_offset = words[1];
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_b200_context_track_offset_set(
    struct pumas_context * context, unsigned long long offset)
{
        if (context == nullptr)
                return RAISE(pumas_b200_context_track_offset_set,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        reinterpret_cast<ContextImpl *>(context)->track_offset = offset;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_b200_context_kernel_set(
    struct pumas_context * context, int kernel)
{
        if (context == nullptr)
                return RAISE(pumas_b200_context_kernel_set,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        if ((kernel < pb::KERNEL_RESIDENT) ||
            (kernel > pb::KERNEL_RESIDENT_SPEC))
                return RAISE(pumas_b200_context_kernel_set,
                    PUMAS_RETURN_INDEX_ERROR, "invalid kernel index [%d]",
                    kernel);
        reinterpret_cast<ContextImpl *>(context)->kernel = kernel;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_b200_context_stream_set(
    struct pumas_context * context, void * cuda_stream)
{
        if (context == nullptr)
                return RAISE(pumas_b200_context_stream_set,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        reinterpret_cast<ContextImpl *>(context)->stream =
            static_cast<cudaStream_t>(cuda_stream);
        return PUMAS_RETURN_SUCCESS;
}

/* ---- recorder, pumas.c:495-505, 2796-2851, 7738-7780 ------------------------ */
namespace {

struct FrameStack {
        FrameStack * next;
        int used, capacity;
        struct pumas_frame frames[1];
};

struct RecorderImpl {
        struct pumas_recorder api;
        struct pumas_frame * last;
        FrameStack * stack;
        double data[1]; /* user memory follows */
};

/* record_state, pumas.c:7738-7780 (frames in blocks of ~4 KB) */
void recorder_append(struct pumas_context * context,
    struct pumas_recorder * recorder, struct pumas_medium * medium,
    enum pumas_event event, struct pumas_state * state)
{
        if (recorder->record != nullptr) {
                recorder->record(context, state, medium, event);
                return;
        }
        RecorderImpl * rec = reinterpret_cast<RecorderImpl *>(recorder);
        FrameStack * stack = rec->stack;
        if ((stack == nullptr) || (stack->used >= stack->capacity)) {
                const int capacity = 32;
                stack = static_cast<FrameStack *>(g_allocate(sizeof(FrameStack) +
                    (capacity - 1) * sizeof(struct pumas_frame)));
                if (stack == nullptr) return;
                stack->used = 0;
                stack->capacity = capacity;
                stack->next = rec->stack;
                rec->stack = stack;
        }
        struct pumas_frame * frame = stack->frames + stack->used++;
        if (recorder->first == nullptr)
                recorder->first = frame;
        else
                rec->last->next = frame;
        rec->last = frame;
        frame->next = nullptr;
        frame->medium = medium;
        frame->event = event;
        memcpy(&frame->state, state, sizeof *state);
        recorder->length++;
}

} /* namespace */

extern "C" {

enum pumas_return pumas_recorder_create(
    struct pumas_recorder ** recorder_, int extra_memory)
{
        if (recorder_ == nullptr)
                return RAISE(pumas_recorder_create, PUMAS_RETURN_VALUE_ERROR,
                    "a NULL recorder pointer was provided");
        *recorder_ = nullptr;
        if (extra_memory < 0) extra_memory = 0;
        RecorderImpl * rec = static_cast<RecorderImpl *>(
            g_allocate(sizeof(RecorderImpl) + (size_t)extra_memory));
        if (rec == nullptr)
                return RAISE(pumas_recorder_create, PUMAS_RETURN_MEMORY_ERROR,
                    "could not allocate memory");
        rec->api.period = 1;
        rec->api.record = nullptr;
        rec->api.length = 0;
        rec->api.first = nullptr;
        rec->api.user_data = (extra_memory > 0) ? rec->data : nullptr;
        rec->last = nullptr;
        rec->stack = nullptr;
        *recorder_ = &rec->api;
        return PUMAS_RETURN_SUCCESS;
}

void pumas_recorder_clear(struct pumas_recorder * recorder)
{
        if (recorder == nullptr) return;
        RecorderImpl * rec = reinterpret_cast<RecorderImpl *>(recorder);
        while (rec->stack != nullptr) {
                FrameStack * next = rec->stack->next;
                g_deallocate(rec->stack);
                rec->stack = next;
        }
        rec->last = nullptr;
        recorder->first = nullptr;
        recorder->length = 0;
}

void pumas_recorder_destroy(struct pumas_recorder ** recorder)
{
        if ((recorder == nullptr) || (*recorder == nullptr)) return;
        pumas_recorder_clear(*recorder);
        g_deallocate(*recorder);
        *recorder = nullptr;
}

/* New: the frames of a recorder as flat arrays (bindings): up to `max` frames
 * are copied; media as layer indices of the context's geometry (-1 = none).
 * Returns the number of frames of the recorder. */
int pumas_b200_recorder_export(const struct pumas_context * context,
    const struct pumas_recorder * recorder, long max,
    struct pumas_state * states, int * media, int * events)
{
        if ((context == nullptr) || (recorder == nullptr)) return 0;
        const ContextImpl * c = reinterpret_cast<const ContextImpl *>(context);
        long k = 0;
        for (const struct pumas_frame * f = recorder->first; f != nullptr;
             f = f->next, k++) {
                if (k >= max) continue;
                if (states != nullptr) states[k] = f->state;
                if (media != nullptr)
                        media[k] = (f->medium == nullptr) ?
                            -1 :
                            (int)(f->medium - c->media_objects);
                if (events != nullptr) events[k] = (int)f->event;
        }
        return (int)k;
}

} /* extern "C" */

/* ---- transport -------------------------------------------------------------- */
namespace {

/* A batch with a recorder attached to the context. The states at `d_states`
 * are transported in place; the frames are handed to the recorder in track
 * order (record_state, pumas.c:7738). */
enum pumas_return transport_recorded(pumas_function_t * caller, ContextImpl * c,
    const pb::Physics & d_view, const pb::Physics & host_view,
    const pb::Settings & cfg, long n, pb::State * d_states, int * d_events,
    int * d_media, cudaStream_t stream)
{
        (void)host_view;
#define RAISE_R(rc, ...) raise(caller, rc, __FILE__, __LINE__, __VA_ARGS__)
        struct pumas_context * context = &c->api;
        struct pumas_recorder * recorder = context->recorder;
        pb::State * d_copy = nullptr;
        int * d_counts = nullptr;
        long * d_first = nullptr;
        pb::DeviceFrame * d_frames = nullptr;
        std::vector<int> counts(n);
        std::vector<long> first(n + 1, 0);
        std::vector<pb::DeviceFrame> frames;
        cudaError_t e = cudaMalloc(&d_copy, n * sizeof(pb::State));
        if (e == cudaSuccess) e = cudaMalloc(&d_counts, n * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&d_first, (n + 1) * sizeof(long));
        /* pass 1: count (on a copy of the inputs) */
        if (e == cudaSuccess)
                e = cudaMemcpyAsync(d_copy, d_states, n * sizeof(pb::State),
                    cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess)
                e = cudaMemsetAsync(c->d_status, 0, sizeof(int), stream);
        if (e == cudaSuccess)
                e = pb::launch_transport_recorded(d_view, cfg, c->geometry,
                    (uint64_t)c->seed, (uint64_t)c->track_offset, n, d_copy,
                    nullptr, nullptr, c->d_status, nullptr, nullptr,
                    recorder->period, nullptr, nullptr, d_counts, stream);
        if (e == cudaSuccess)
                e = cudaMemcpyAsync(counts.data(), d_counts, n * sizeof(int),
                    cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess) {
                for (long i = 0; i < n; i++) first[i + 1] = first[i] + counts[i];
                e = cudaMalloc(&d_frames,
                    (first[n] > 0 ? first[n] : 1) * sizeof(pb::DeviceFrame));
        }
        /* pass 2: record, on the caller's states */
        if (e == cudaSuccess)
                e = cudaMemcpyAsync(d_first, first.data(),
                    (n + 1) * sizeof(long), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess)
                e = cudaMemsetAsync(c->d_status, 0, sizeof(int), stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev0, stream);
        if (e == cudaSuccess)
                e = pb::launch_transport_recorded(d_view, cfg, c->geometry,
                    (uint64_t)c->seed, (uint64_t)c->track_offset, n, d_states,
                    d_events, d_media, c->d_status, c->d_random, c->d_steps,
                    recorder->period, d_frames, d_first, d_counts, stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev1, stream);
        if (e == cudaSuccess) {
                frames.resize(first[n]);
                if (first[n] > 0)
                        e = cudaMemcpyAsync(frames.data(), d_frames,
                            first[n] * sizeof(pb::DeviceFrame),
                            cudaMemcpyDeviceToHost, stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (d_copy) cudaFree(d_copy);
        if (d_counts) cudaFree(d_counts);
        if (d_first) cudaFree(d_first);
        if (d_frames) cudaFree(d_frames);
        if (e != cudaSuccess)
                return RAISE_R(PUMAS_RETURN_INTERNAL_ERROR, "CUDA error (%s)",
                    cudaGetErrorString(e));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->last_ms = ms;
        for (long k = 0; k < first[n]; k++) {
                pb::DeviceFrame & f = frames[k];
                struct pumas_medium * medium = ((f.medium < 0) ||
                    (f.medium >= PUMAS_B200_MAX_LAYERS)) ?
                    nullptr :
                    &c->media_objects[f.medium];
                recorder_append(context, recorder, medium,
                    (enum pumas_event)f.event,
                    reinterpret_cast<struct pumas_state *>(&f.state));
        }
        return PUMAS_RETURN_SUCCESS;
#undef RAISE_R
}

/* One batch, for the public entry point `caller` (errors are attributed to
 * it, as the reference does). `async`: enqueue on the context's stream and
 * return; every pointer must then be a device pointer, and
 * pumas_b200_context_synchronize collects the status. */
enum pumas_return transport_batch(pumas_function_t * caller,
    struct pumas_context * context, long n, struct pumas_state * states,
    int * events, int * media, bool async)
{
#define RAISE_C(rc, ...) raise(caller, rc, __FILE__, __LINE__, __VA_ARGS__)
        if (context == nullptr)
                return RAISE_C(PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        if ((states == nullptr) && (n > 0))
                return RAISE_C(PUMAS_RETURN_VALUE_ERROR, "no state (null)");
        if (n < 0)
                return RAISE_C(PUMAS_RETURN_VALUE_ERROR, "negative number of states");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        pumas_physics * physics = const_cast<pumas_physics *>(c->physics);
        if (physics == nullptr)
                return RAISE_C(PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        if (!c->has_geometry && !c->has_callbacks) {
                /* a host `pumas_medium_cb` cannot run inside a kernel: say so
                 * instead of ignoring it */
                if (context->medium != nullptr)
                        return RAISE_C(PUMAS_RETURN_MEDIUM_ERROR,
                            "context->medium is a host callback: it cannot "
                            "run on the device. Describe the geometry with "
                            "pumas_b200_context_geometry_set or give device "
                            "functions to pumas_b200_context_callbacks_set");
                return RAISE_C(PUMAS_RETURN_MEDIUM_ERROR, "no medium specified");
        }
        if ((context->recorder != nullptr) && async)
                return RAISE_C(PUMAS_RETURN_VALUE_ERROR,
                    "a recorder needs the synchronous entry points");
        if ((context->recorder != nullptr) && c->has_callbacks)
                return RAISE_C(PUMAS_RETURN_VALUE_ERROR,
                    "the recorder runs on the built-in media only (set the "
                    "geometry with pumas_b200_context_geometry_set)");

        /* Configuration checks, pumas.c:3152-3177. */
        const pb::Settings cfg = settings_from(context);
        const pb::Physics & hp = physics->host.view;
        if ((hp.particle == PUMAS_PARTICLE_TAU) &&
            (cfg.direction == pb::MODE_FORWARD) &&
            (cfg.decay == pb::MODE_WEIGHTED))
                return RAISE_C(PUMAS_RETURN_DECAY_ERROR,
                    "`PUMAS_MODE_WEIGHTED' mode is not valid for forward taus");
        if ((cfg.accuracy <= 0) || (cfg.accuracy > 1))
                return RAISE_C(PUMAS_RETURN_ACCURACY_ERROR,
                    "bad accuracy value (expected a value in ]0,1], got %g)",
                    cfg.accuracy);
        if ((cfg.direction == pb::MODE_BACKWARD) &&
            (cfg.energy_loss > pb::MODE_CSDA) && (hp.cutoff < 1E-02))
                return RAISE_C(PUMAS_RETURN_ACCURACY_ERROR,
                    "bad cutoff value for backward transport (expected a "
                    "value greater than or equal to 0.01, got %g)",
                    hp.cutoff);
        if ((cfg.energy_loss < pb::MODE_DISABLED) ||
            (cfg.energy_loss > pb::MODE_STRAGGLED))
                return RAISE_C(PUMAS_RETURN_INDEX_ERROR,
                    "invalid energy loss scheme [%d]", cfg.energy_loss);
        c->last_n = 0;
        c->last_ms = 0.;
        c->last_launches = 0;
        if (n == 0) return PUMAS_RETURN_SUCCESS;

        /* The device: there is no CPU path. A context lives on the device
         * that was current at its first batch. */
        std::string why;
        if (!device_usable(&why))
                return RAISE_C(PUMAS_RETURN_INTERNAL_ERROR, "%s", why.c_str());
        {
                int dev = 0;
                cudaGetDevice(&dev);
                if (c->device < 0) c->device = dev;
                if (c->device != dev)
                        return RAISE_C(PUMAS_RETURN_INTERNAL_ERROR,
                            "the context was first used on CUDA device %d, "
                            "the current device is %d (one context per "
                            "device)",
                            c->device, dev);
        }
        pb::Physics d_view;
        enum pumas_return rc = physics_upload(caller, physics, &d_view);
        if (rc != PUMAS_RETURN_SUCCESS) return rc;
        /* user DCS models (pumas_dcs_register): the device functions, called
         * by the device-linked kernel (kernels_callback.cu) only */
        bool user_models = false;
        for (int i = 0; i < 3; i++) {
                if (hp.model[i] != pb::MODEL_USER) continue;
                std::lock_guard<std::mutex> lock(g_user_models_mutex);
                const UserModel * m =
                    user_model_find(i, physics->host.model[i].c_str());
                if ((m == nullptr) || (m->device == nullptr))
                        return RAISE_C(PUMAS_RETURN_MODEL_ERROR,
                            "the %s model of the %s process has no device "
                            "function (pumas_b200_dcs_register_device)",
                            physics->host.model[i].c_str(), PROCESS_NAME[i]);
                d_view.user_dcs[i] = m->device;
                user_models = true;
        }
        if (user_models && (context->recorder != nullptr))
                return RAISE_C(PUMAS_RETURN_MODEL_ERROR,
                    "the recorder kernel does not call user DCS models");

        if (!c->seed_set) {
                rc = pumas_context_random_seed_set(context, nullptr);
                if (rc != PUMAS_RETURN_SUCCESS) return rc;
        }

        cudaError_t e = context_reserve(c, n);
        if (e != cudaSuccess)
                return RAISE_C(PUMAS_RETURN_MEMORY_ERROR,
                    "could not allocate device memory (%s)",
                    cudaGetErrorString(e));

        const cudaStream_t stream = c->stream;
        const bool states_on_device = is_device_pointer(states);
        if (async &&
            (!states_on_device ||
                ((events != nullptr) && !is_device_pointer(events)) ||
                ((media != nullptr) && !is_device_pointer(media))))
                return RAISE_C(PUMAS_RETURN_VALUE_ERROR,
                    "asynchronous transport needs device pointers");
        const bool events_on_device =
            (events != nullptr) && is_device_pointer(events);
        const bool media_on_device =
            (media != nullptr) && is_device_pointer(media);
        pb::State * d_states = states_on_device ?
            reinterpret_cast<pb::State *>(states) :
            c->d_states;
        int * d_events = events_on_device ? events : c->d_events;
        int * d_media = media_on_device ? media : c->d_media;

        if (!states_on_device)
                e = cudaMemcpyAsync(d_states, states, n * sizeof(pb::State),
                    cudaMemcpyHostToDevice, stream);
        if ((e == cudaSuccess) && (context->recorder != nullptr)) {
                /* the recorder (pumas.h:432-458): two passes of the
                 * recording kernel over the same inputs and streams, the
                 * first counts the frames of every track, the second stores
                 * them; then the frames go to the recorder in track order */
                const enum pumas_return rr = transport_recorded(caller, c, d_view,
                    physics->host.view, cfg, n, d_states, d_events, d_media,
                    stream);
                if (rr != PUMAS_RETURN_SUCCESS) return rr;
                int status = 0;
                if (!states_on_device)
                        e = cudaMemcpyAsync(states, d_states,
                            n * sizeof(pb::State), cudaMemcpyDeviceToHost,
                            stream);
                if ((e == cudaSuccess) && (events != nullptr) &&
                    !events_on_device)
                        e = cudaMemcpyAsync(events, d_events, n * sizeof(int),
                            cudaMemcpyDeviceToHost, stream);
                if ((e == cudaSuccess) && (media != nullptr) &&
                    !media_on_device)
                        e = cudaMemcpyAsync(media, d_media, 2 * n * sizeof(int),
                            cudaMemcpyDeviceToHost, stream);
                if (e == cudaSuccess)
                        e = cudaMemcpyAsync(&status, c->d_status, sizeof(int),
                            cudaMemcpyDeviceToHost, stream);
                if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
                if (e != cudaSuccess)
                        return RAISE_C(PUMAS_RETURN_INTERNAL_ERROR,
                            "CUDA error (%s)", cudaGetErrorString(e));
                c->last_launches = 2;
                c->last_n = n;
                c->track_offset += (unsigned long long)n;
                if (status != 0)
                        return RAISE_C((enum pumas_return)status, "%s",
                            return_text(status));
                return PUMAS_RETURN_SUCCESS;
        }
        if ((e == cudaSuccess) && !c->pending)
                e = cudaMemsetAsync(c->d_status, 0, sizeof(int), stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev0, stream);
        if ((e == cudaSuccess) && user_models && !c->has_callbacks) {
                /* the built-in media behind the callback interface */
                if (!c->has_planar) {
                        e = pb::planar_callbacks(c->geometry, &c->planar);
                        if (e == cudaSuccess) c->has_planar = 1;
                }
        }
        if ((e == cudaSuccess) && (c->has_callbacks || user_models))
                e = pb::launch_transport_callbacks(d_view, cfg,
                    c->has_callbacks ? c->callbacks : c->planar,
                    (uint64_t)c->seed, (uint64_t)c->track_offset, n, d_states,
                    d_events, d_media, c->d_status, c->d_random, c->d_steps,
                    c->d_workspace, stream);
        else if (e == cudaSuccess)
                e = pb::launch_transport(d_view, physics->host.view, cfg,
                    c->geometry,
                    (uint64_t)c->seed, (uint64_t)c->track_offset, n, d_states,
                    d_events, d_media, c->d_status, c->d_random, c->d_steps,
                    c->d_workspace, c->kernel, stream);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev1, stream);
        if (async) {
                if (e != cudaSuccess)
                        return RAISE_C(PUMAS_RETURN_INTERNAL_ERROR,
                            "CUDA error (%s)", cudaGetErrorString(e));
                c->pending = 1;
                c->last_launches = pb::transport_last_launches();
                c->last_n = n;
                c->track_offset += (unsigned long long)n;
                return PUMAS_RETURN_SUCCESS;
        }
        int status = 0;
        if ((e == cudaSuccess) && !states_on_device)
                e = cudaMemcpyAsync(states, d_states, n * sizeof(pb::State),
                    cudaMemcpyDeviceToHost, stream);
        if ((e == cudaSuccess) && (events != nullptr) && !events_on_device)
                e = cudaMemcpyAsync(events, d_events, n * sizeof(int),
                    cudaMemcpyDeviceToHost, stream);
        if ((e == cudaSuccess) && (media != nullptr) && !media_on_device)
                e = cudaMemcpyAsync(media, d_media, 2 * n * sizeof(int),
                    cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess)
                e = cudaMemcpyAsync(&status, c->d_status, sizeof(int),
                    cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess)
                return RAISE_C(PUMAS_RETURN_INTERNAL_ERROR, "CUDA error (%s)",
                    cudaGetErrorString(e));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->last_ms = ms;
        c->last_launches = pb::transport_last_launches();
        c->last_n = n;
        c->track_offset += (unsigned long long)n;

        c->pending = 0;
        if (status != 0)
                return RAISE_C((enum pumas_return)status, "%s", return_text(status));
        return PUMAS_RETURN_SUCCESS;
#undef RAISE_C
}

} /* namespace */

enum pumas_return pumas_context_transport_n(struct pumas_context * context,
    long n, struct pumas_state * states, int * events, int * media)
{
        return transport_batch(FN(pumas_context_transport_n), context, n,
            states, events, media, false);
}

enum pumas_return pumas_b200_context_transport_async(
    struct pumas_context * context, long n, struct pumas_state * states,
    int * events, int * media)
{
        return transport_batch(FN(pumas_b200_context_transport_async), context,
            n, states, events, media, true);
}

enum pumas_return pumas_b200_context_synchronize(struct pumas_context * context)
{
        if (context == nullptr)
                return RAISE(pumas_b200_context_synchronize,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        if (!c->pending) return PUMAS_RETURN_SUCCESS;
        int status = 0;
        cudaError_t e = cudaMemcpyAsync(&status, c->d_status, sizeof(int),
            cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        c->pending = 0;
        if (e != cudaSuccess)
                return RAISE(pumas_b200_context_synchronize,
                    PUMAS_RETURN_INTERNAL_ERROR, "CUDA error (%s)",
                    cudaGetErrorString(e));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->last_ms = ms; /* (the last batch of the series) */
        if (status != 0)
                return RAISE(pumas_b200_context_synchronize,
                    (enum pumas_return)status, "%s", return_text(status));
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_context_transport(struct pumas_context * context,
    struct pumas_state * state, enum pumas_event * event,
    struct pumas_medium * media[2])
{
        if (context == nullptr)
                return RAISE(pumas_context_transport, PUMAS_RETURN_VALUE_ERROR,
                    "no context (null)");
        if (state == nullptr)
                return RAISE(pumas_context_transport, PUMAS_RETURN_VALUE_ERROR,
                    "no state (null)");
        if (media != nullptr) media[0] = media[1] = nullptr;
        int e = PUMAS_EVENT_NONE;
        const enum pumas_return rc = transport_batch(
            FN(pumas_context_transport), context, 1, state, &e, nullptr, false);
        if (event != nullptr) *event = (enum pumas_event)e;
        return rc;
}

enum pumas_return pumas_b200_context_last_counters(
    struct pumas_context * context, long n, unsigned int * n_random,
    unsigned int * n_steps)
{
        if (context == nullptr)
                return RAISE(pumas_b200_context_last_counters,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        if ((n < 0) || (n > c->last_n))
                return RAISE(pumas_b200_context_last_counters,
                    PUMAS_RETURN_VALUE_ERROR,
                    "the last batch held %ld tracks, %ld requested", c->last_n,
                    n);
        cudaError_t e = cudaSuccess;
        if (n_random != nullptr)
                e = cudaMemcpy(n_random, c->d_random, n * sizeof(unsigned int),
                    cudaMemcpyDeviceToHost);
        if ((e == cudaSuccess) && (n_steps != nullptr))
                e = cudaMemcpy(n_steps, c->d_steps, n * sizeof(unsigned int),
                    cudaMemcpyDeviceToHost);
        if (e != cudaSuccess)
                return RAISE(pumas_b200_context_last_counters,
                    PUMAS_RETURN_INTERNAL_ERROR, "CUDA error (%s)",
                    cudaGetErrorString(e));
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_b200_context_last_kernel_time(
    struct pumas_context * context, double * milliseconds, int * launches)
{
        if (context == nullptr)
                return RAISE(pumas_b200_context_last_kernel_time,
                    PUMAS_RETURN_VALUE_ERROR, "no context (null)");
        ContextImpl * c = reinterpret_cast<ContextImpl *>(context);
        if (milliseconds != nullptr) *milliseconds = c->last_ms;
        if (launches != nullptr) *launches = c->last_launches;
        return PUMAS_RETURN_SUCCESS;
}

} /* extern "C" */

/* ---- DCS models on the host: pumas.c:13673-13699, 15376-15543 ---------------
 * The functions below evaluate the very code the kernels run (pb_dcs.h,
 * pb_dcs_models.h are single-source host/device), behind the reference's
 * pumas_dcs_t signature. */
namespace {

const pb::Physics & model_constants()
{
        static pb::Physics constants;
        static std::once_flag once;
        std::call_once(once, [] {
                std::memset(&constants, 0, sizeof constants);
                /* as load_physics_bytes (pb_dump.hpp): run time libm */
                volatile double one = 1., half = 0.5, m16 = -1. / 6.,
                                t13 = 1. / 3., m13 = -1. / 3.;
                constants.c_exp_1 = exp(one);
                constants.c_exp_1_2 = exp(half);
                constants.c_exp_m1_6 = exp(m16);
                constants.c_exp_1_3 = exp(t13);
                constants.c_exp_m1_3 = exp(m13);
        });
        return constants;
}

#define MODEL_WITH_CONSTANTS(name)                                             \
        double api_##name(double Z, double A, double m, double K, double q)    \
        {                                                                      \
                pb::Element el;                                                \
                pb::detail::element_fill(el, Z, A, 0.);                        \
                return pb::name(model_constants(), el, m, K, q);               \
        }
#define MODEL(name)                                                            \
        double api_##name(double Z, double A, double m, double K, double q)    \
        {                                                                      \
                pb::Element el;                                                \
                pb::detail::element_fill(el, Z, A, 0.);                        \
                return pb::name(el, m, K, q);                                  \
        }
MODEL_WITH_CONSTANTS(dcs_bremsstrahlung_SSR)
MODEL(dcs_bremsstrahlung_KKP)
MODEL(dcs_bremsstrahlung_ABB)
MODEL_WITH_CONSTANTS(dcs_pair_production_SSR)
MODEL(dcs_pair_production_KKP)
MODEL(dcs_photonuclear_DRSS)
MODEL(dcs_photonuclear_BM)
MODEL(dcs_photonuclear_BBKS)
#undef MODEL
#undef MODEL_WITH_CONSTANTS

/* Indexed like pb::Physics::model (pb_dump.hpp, `known`). */
struct ModelEntry {
        const char * name;
        pumas_dcs_t * dcs;
};
const ModelEntry MODELS[3][3] = {
        { { "SSR", &api_dcs_bremsstrahlung_SSR },
            { "KKP", &api_dcs_bremsstrahlung_KKP },
            { "ABB", &api_dcs_bremsstrahlung_ABB } },
        { { "SSR", &api_dcs_pair_production_SSR },
            { "KKP", &api_dcs_pair_production_KKP }, { nullptr, nullptr } },
        { { "DRSS", &api_dcs_photonuclear_DRSS },
            { "BM", &api_dcs_photonuclear_BM },
            { "BBKS", &api_dcs_photonuclear_BBKS } }
};
} /* namespace */

extern "C" {

enum pumas_return pumas_physics_dcs(const struct pumas_physics * physics,
    enum pumas_process process, const char ** model, pumas_dcs_t ** dcs)
{
        if (physics == nullptr)
                return RAISE(pumas_physics_dcs, PUMAS_RETURN_PHYSICS_ERROR,
                    "the Physics has not been initialised");
        if (((int)process < 0) || (process > PUMAS_PROCESS_PHOTONUCLEAR))
                return RAISE(pumas_physics_dcs, PUMAS_RETURN_INDEX_ERROR,
                    "bad process (expected a value in [0, 2], got %u)",
                    (unsigned int)process);
        if (physics->host.view.model[process] == pb::MODEL_USER) {
                if (dcs != nullptr)
                        *dcs = physics->host.view.user_dcs[process];
                if (model != nullptr)
                        *model = physics->host.model[process].c_str();
                return PUMAS_RETURN_SUCCESS;
        }
        const ModelEntry & entry =
            MODELS[process][physics->host.view.model[process]];
        if (dcs != nullptr) *dcs = entry.dcs;
        if (model != nullptr) *model = entry.name;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_dcs_get(
    enum pumas_process process, const char * model, pumas_dcs_t ** dcs)
{
        if (((int)process < 0) || (process > PUMAS_PROCESS_PHOTONUCLEAR))
                return RAISE(pumas_dcs_get, PUMAS_RETURN_INDEX_ERROR,
                    "bad process (expected an index in [0, %d], got %u)",
                    (int)PUMAS_PROCESS_PHOTONUCLEAR, (unsigned int)process);
        if (model == nullptr) model = MODELS[process][0].name;
        for (int k = 0; k < 3; k++) {
                const ModelEntry & entry = MODELS[process][k];
                if ((entry.name != nullptr) &&
                    (strcmp(entry.name, model) == 0)) {
                        *dcs = entry.dcs;
                        return PUMAS_RETURN_SUCCESS;
                }
        }
        {
                std::lock_guard<std::mutex> lock(g_user_models_mutex);
                const UserModel * m = user_model_find((int)process, model);
                if (m != nullptr) {
                        *dcs = m->host;
                        return PUMAS_RETURN_SUCCESS;
                }
        }
        *dcs = nullptr;
        return RAISE(pumas_dcs_get, PUMAS_RETURN_MODEL_ERROR,
            "model %s not found for %s process", model, PROCESS_NAME[process]);
}

/* pumas.c:15422-15460. The host function serves the host side API
 * (pumas_dcs_get, pumas_physics_dcs) as in the reference; the transport
 * kernels need its `__device__` counterpart as well, given by
 * pumas_b200_dcs_register_device. */
enum pumas_return pumas_dcs_register(
    enum pumas_process process, const char * model, pumas_dcs_t * dcs)
{
        if (((int)process < 0) || (process > PUMAS_PROCESS_PHOTONUCLEAR))
                return RAISE(pumas_dcs_register, PUMAS_RETURN_INDEX_ERROR,
                    "bad process (expected an index in [0, %d], got %u)",
                    (int)PUMAS_PROCESS_PHOTONUCLEAR, (unsigned int)process);
        if (dcs == nullptr)
                return RAISE(pumas_dcs_register, PUMAS_RETURN_VALUE_ERROR,
                    "bad dcs (expected a function, got nil)");
        if (model == nullptr)
                return RAISE(pumas_dcs_register, PUMAS_RETURN_VALUE_ERROR,
                    "bad model (expected a string, got nil)");
        bool known = false;
        for (int k = 0; k < 3; k++) {
                const ModelEntry & entry = MODELS[process][k];
                if ((entry.name != nullptr) && (strcmp(entry.name, model) == 0))
                        known = true;
        }
        std::lock_guard<std::mutex> lock(g_user_models_mutex);
        if (known || (user_model_find((int)process, model) != nullptr))
                return RAISE(pumas_dcs_register, PUMAS_RETURN_MODEL_ERROR,
                    "model %s already registered for %s process", model,
                    PROCESS_NAME[process]);
        if (g_user_models.size() >= USER_MODEL_STACK)
                return RAISE(pumas_dcs_register, PUMAS_RETURN_MEMORY_ERROR,
                    "max stack size reached");
        UserModel m;
        m.process = (int)process;
        m.name = model;
        m.host = dcs;
        m.device = nullptr;
        g_user_models.push_back(m);
        return PUMAS_RETURN_SUCCESS;
}

/* New. The `__device__` function of a model registered with
 * pumas_dcs_register: a device function pointer of the device link unit of the
 * kernels (`python -m pumas_b200.build --user`). */
enum pumas_return pumas_b200_dcs_register_device(
    enum pumas_process process, const char * model, pumas_dcs_t * device_dcs)
{
        if (((int)process < 0) || (process > PUMAS_PROCESS_PHOTONUCLEAR))
                return RAISE(pumas_b200_dcs_register_device,
                    PUMAS_RETURN_INDEX_ERROR,
                    "bad process (expected an index in [0, %d], got %u)",
                    (int)PUMAS_PROCESS_PHOTONUCLEAR, (unsigned int)process);
        if ((device_dcs == nullptr) || (model == nullptr))
                return RAISE(pumas_b200_dcs_register_device,
                    PUMAS_RETURN_VALUE_ERROR, "a NULL pointer was provided");
        std::lock_guard<std::mutex> lock(g_user_models_mutex);
        for (size_t i = 0; i < g_user_models.size(); i++) {
                UserModel & m = g_user_models[i];
                if ((m.process == (int)process) && (m.name == model)) {
                        m.device = device_dcs;
                        return PUMAS_RETURN_SUCCESS;
                }
        }
        return RAISE(pumas_b200_dcs_register_device, PUMAS_RETURN_MODEL_ERROR,
            "cannot find %s model for %s process (register the host function "
            "with pumas_dcs_register first)", model, PROCESS_NAME[process]);
}

const char * pumas_dcs_default(enum pumas_process process)
{
        if (((int)process < 0) || (process > PUMAS_PROCESS_PHOTONUCLEAR))
                return nullptr;
        return MODELS[process][0].name;
}

enum pumas_return pumas_dcs_range(enum pumas_process process, double Z,
    double mass, double energy, double * min, double * max)
{
        if (((int)process < 0) || (process > PUMAS_PROCESS_PHOTONUCLEAR))
                return RAISE(pumas_dcs_range, PUMAS_RETURN_INDEX_ERROR,
                    "bad process (expected an index in [0, %d], got %u)",
                    (int)PUMAS_PROCESS_PHOTONUCLEAR, (unsigned int)process);
        double qmin, qmax;
        if (process == PUMAS_PROCESS_PHOTONUCLEAR) {
                pb::dcs_photonuclear_range(mass, energy, &qmin, &qmax);
        } else {
                pb::Element el;
                pb::detail::element_fill(el, Z, 1., 0.);
                pb::dcs_pair_production_range(el, mass, energy, &qmin, &qmax);
                if (process == PUMAS_PROCESS_BREMSSTRAHLUNG) qmin = 0.;
        }
        if (min != nullptr) *min = qmin;
        if (max != nullptr) *max = qmax;
        return PUMAS_RETURN_SUCCESS;
}

/* pumas.c:15545-15581, on the Coulomb functions of the kernels (pb_coulomb.h). */
double pumas_elastic_dcs(double Z, double A, double m, double K, double theta)
{
        if ((Z < 1.) || (A < 1.) || (m <= 0.) || (K <= 0.)) return 0.;

        pb::Element el;
        pb::detail::element_fill(el, Z, A, 0.);
        int n_parameters;
        double kinetic0, amplitude[3], screening[pb::MAX_SCREENING + 1];
        double fCM[2];
        pb::coulomb_frame_parameters(Z, A, m, K, &kinetic0, fCM);
        pb::coulomb_screening_parameters(
            el, m, K, kinetic0, &n_parameters, amplitude, screening);
        const double fspin = pb::coulomb_spin_factor(m, K);

        const double c = cos(theta);
        const double s = sin(theta);
        const double gs = fCM[0] * s;
        const double sd = sqrt(c * c + gs * gs * (1. - fCM[1] * fCM[1]));
        const double cgs = c * c + gs * gs;
        const double mu0 =
            0.5 * (gs * gs * (1. + fCM[1]) + c * (c - sd)) / cgs;
        const double num = fCM[0] * (c * fCM[1] + sd);
        const double jac = num * num / (cgs * cgs * sd);

        const int n = n_parameters - 1;
        double dcs = 0.;
        for (int i = 0; i < n; i++) dcs += amplitude[i] / (screening[i] + mu0);
        dcs *= pb::coulomb_nuclear_form_factor(mu0, screening[n]);
        dcs *= dcs;

        const double p0 = sqrt(kinetic0 * (kinetic0 + 2. * m));
        const double p = sqrt(K * (K + 2. * m));
        const double factor =
            0.5 * pb::ALPHA_EM * Z * (K + m) * pb::HBAR_C / (p0 * p);
        return factor * factor * dcs * jac * (1. - fspin * mu0);
}

/* pumas.c:15584-15609. Host only: the pole reduction and the transport
 * coefficients run in `long double`, as in the reference (the kernels use the
 * double-double instantiation of the same templates, pb_coulomb.h). */
#pragma nv_diag_suppress 20208
double pumas_elastic_path(
    int order, double Z, double A, double mass, double kinetic)
{
        if ((order < 0) || (order > 1) || (kinetic <= 0.)) return 0.;

        typedef long double real;
        pb::Element el;
        pb::detail::element_fill(el, Z, A, 0.);
        int n_parameters;
        double kinetic0, amplitude[3], screening[pb::MAX_SCREENING + 1];
        double fCM[2];
        pb::coulomb_frame_parameters(Z, A, mass, kinetic, &kinetic0, fCM);
        pb::coulomb_screening_parameters(
            el, mass, kinetic, kinetic0, &n_parameters, amplitude, screening);
        const double fspin = pb::coulomb_spin_factor(mass, kinetic);
        pb::CoulombPoles<real> poles;
        pb::coulomb_pole_reduction(n_parameters, amplitude, screening, &poles);

        double coefficient;
        if (order == 0) {
                coefficient = pb::coulomb_transport_cs0(
                    1., fspin, n_parameters, screening, poles);
        } else {
                /* Order 1 coefficient over the whole angular range (mu = 1),
                 * pumas.c:7623-7683: per pole, the moments of 1 / (x + mu) and
                 * 1 / (x + mu)^2 weighted by mu (1 - S mu); the nuclear term
                 * adds the poles of order 3 and 4. The accumulator is binary64
                 * (each term is rounded when added). */
                const real S = fspin, mu = 1.;
                const int n = n_parameters - 1;
                double cs1 = 0.;
                real x, r, L, I0, I1, J1, I2, J2;
                auto moments = [&](double pole) {
                        x = pole;
                        r = mu / (mu + x);
                        L = logl(1. + mu / x);
                        I0 = r / x;
                        I1 = L - r;
                        J1 = -x * L;
                        I2 = mu + x * (r - 2. * L);
                        J2 = x * (x * L - mu);
                };
                for (int i = 0; i < n; i++) {
                        moments(screening[i]);
                        cs1 += poles.a[i] * (J1 - S * J2) +
                            poles.b[i] * (I1 - S * I2);
                }
                moments(screening[n]);
                const real rn = 1. / (1. + mu / x);
                const real K0 = (1. - rn) * (1. + rn) / (2 * x * x);
                const real L0 = (1. - rn) * ((1. + rn) * (1. + rn) - rn) /
                    (3 * x * x * x);
                const real K1 = I0 - x * K0;
                const real L1 = K0 - x * L0;
                const real K2 = I1 - x * K1;
                const real L2 = K1 - x * L1;
                cs1 += poles.c[0] * (J1 - S * J2) + poles.c[1] * (I1 - S * I2) +
                    poles.c[2] * (K1 - S * K2) + poles.c[3] * (L1 - S * L2);
                const double d = 1. / (fCM[0] * (1. + fCM[1]));
                coefficient = 2 * cs1;
                coefficient *= d * d;
        }
        return pb::coulomb_normalisation(Z, A, mass, kinetic, kinetic0) /
            coefficient;
}
#pragma nv_diag_default 20208

double pumas_electronic_dcs(double Z, double I, double m, double K, double q)
{
        return pb::electronic_dcs(Z, I, m, K, q);
}

} /* extern "C" */
