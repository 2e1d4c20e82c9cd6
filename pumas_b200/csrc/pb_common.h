/*
 * pb_common.h -- shared definitions of the transport core.
 *
 * The transport core (pb_*.h) is single-source C++: nvcc compiles it into the
 * sm_100a kernels of libpumas_b200.so; g++ compiles the very same headers into
 * the CPU debugging build under oracle/_port (tests only), which is how the
 * algorithm is checked against the reference on a box without a GPU.
 *
 * Constants are those of the reference (src/pumas.c:60-200) and must stay
 * bit-identical.
 */
#ifndef PB_COMMON_H
#define PB_COMMON_H

#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PB_FN __host__ __device__ __forceinline__
#define PB_NOINLINE inline __host__ __device__ __noinline__
#else
#define PB_FN inline
#define PB_NOINLINE inline
#endif

namespace pb {

/* Out-of-line libm calls. Inlined, each CUDA log / exp / sincos / pow is 60 to
 * 300 instructions; the transport code calls them at ~150 sites, which alone
 * made the kernel several times larger than the instruction cache and
 * `no_instruction` its first stall reason (profiles/r1_v1). One shared copy
 * each keeps the stepping loop cache resident; results are unchanged. */
#if defined(__CUDACC__)
#define PB_MATH inline __host__ __device__ __noinline__
#else
#define PB_MATH inline
#endif
PB_MATH double xlog(double x) { return log(x); }
PB_MATH double xexp(double x) { return exp(x); }
PB_MATH double xsin(double x) { return sin(x); }
PB_MATH double xcos(double x) { return cos(x); }
/* sin and cos of the same argument. The device sincos() shares the argument
 * reduction and returns the values sin() and cos() return. */
struct SinCos {
        double s, c;
};
PB_MATH SinCos xsincos(double x)
{
        SinCos r;
#if defined(__CUDA_ARCH__)
        sincos(x, &r.s, &r.c);
#else
        r.c = cos(x);
        r.s = sin(x);
#endif
        return r;
}
PB_MATH double xatan(double x) { return atan(x); }
PB_MATH double xacos(double x) { return acos(x); }
PB_MATH double xlog10(double x) { return log10(x); }
PB_MATH double xpow(double x, double y) { return pow(x, y); }

/* pumas.c:60-125 */
constexpr int N_SCHEMES = 2;
constexpr int N_LARMOR_ORDERS = 8;
constexpr int N_DEL_PROCESSES = 4;
/* Largest number of elements per material the hard elastic event handles. */
constexpr int MAX_COMPONENTS = 8;
constexpr double BMC_ALPHA = 2.0;
constexpr double EPSILON_X = 3E-03;
constexpr double STEP_MIN = 1E-07;
/* pumas.c:130-188 */
constexpr double ALPHA_EM = 7.2973525693E-03;
constexpr double HBAR_C = 1.973269804E-16;
constexpr double BOHR_RADIUS = 0.529177210903E-10;
constexpr double MUON_C_TAU = 658.654;
constexpr double TAU_C_TAU = 87.03E-06;
constexpr double LARMOR_FACTOR = 0.299792458;
constexpr double ELECTRON_MASS = 0.510998910E-03;
constexpr double ELECTRON_RADIUS = 2.817940285E-15;
constexpr double MUON_MASS = 0.10565839;
constexpr double TAU_MASS = 1.77682;
constexpr double PROTON_MASS = 0.938272;
constexpr double NEUTRON_MASS = 0.939565;
constexpr double PION_MASS = 0.13957018;
constexpr double PI = 3.14159265358979323846;
constexpr double AVOGADRO_NUMBER = 6.02214076E+23;

/* enum pumas_mode (pumas.h:90-140) */
constexpr int MODE_DISABLED = -1;
constexpr int MODE_CSDA = 0;
constexpr int MODE_MIXED = 1;
constexpr int MODE_STRAGGLED = 2;
constexpr int MODE_WEIGHTED = 0;
constexpr int MODE_RANDOMISED = 1;
constexpr int MODE_FORWARD = 0;
constexpr int MODE_BACKWARD = 1;

/* enum pumas_event (pumas.h:199-236) */
constexpr int EVENT_NONE = 0;
constexpr int EVENT_LIMIT_ENERGY = 1;
constexpr int EVENT_LIMIT_DISTANCE = 2;
constexpr int EVENT_LIMIT_GRAMMAGE = 4;
constexpr int EVENT_LIMIT_TIME = 8;
constexpr int EVENT_LIMIT = 15;
constexpr int EVENT_MEDIUM = 16;
constexpr int EVENT_VERTEX_BREMSSTRAHLUNG = 32;
constexpr int EVENT_VERTEX_PAIR_CREATION = 64;
constexpr int EVENT_VERTEX_PHOTONUCLEAR = 128;
constexpr int EVENT_VERTEX_DELTA_RAY = 256;
constexpr int EVENT_VERTEX_DEL = 480;
constexpr int EVENT_VERTEX_COULOMB = 512;
constexpr int EVENT_VERTEX_DECAY = 1024;
constexpr int EVENT_VERTEX = 2016;
constexpr int EVENT_WEIGHT = 2048;

/* enum pumas_step (pumas.h:461) */
constexpr int STEP_CHECK = 0;
constexpr int STEP_RAW = 1;

/* enum pumas_return codes raised from inside the transport (pumas.h:143) */
constexpr int RETURN_SUCCESS = 0;
constexpr int RETURN_DENSITY_ERROR = 4;
constexpr int RETURN_VALUE_ERROR = 24;

/* Binary layout of struct pumas_state (pumas.h:289-309). */
struct State {
        double charge;
        double energy;
        double distance;
        double grammage;
        double time;
        double weight;
        double position[3];
        double direction[3];
        int decayed;
};

/* The public, per batch, settings of a pumas_context (pumas.h:636-672). */
struct Settings {
        int energy_loss;
        int decay;
        int direction;
        int scattering;
        int event;
        double limit_energy;
        double limit_distance;
        double limit_grammage;
        double limit_time;
        double accuracy;
};

/* The same settings with the four modes fixed at compile time: the transport
 * functions are templates on the settings type and read `cfg.energy_loss` etc.
 * either way, so a kernel instantiated on StaticSettings keeps only the code
 * of its own configuration (no CSDA / mixed branches in a straggled kernel, no
 * backward weights in a forward one ...). */
template <int ENERGY_LOSS, int SCATTERING, int DIRECTION, int DECAY>
struct StaticSettings {
        static constexpr int energy_loss = ENERGY_LOSS;
        static constexpr int decay = DECAY;
        static constexpr int direction = DIRECTION;
        static constexpr int scattering = SCATTERING;
        int event;
        double limit_energy;
        double limit_distance;
        double limit_grammage;
        double limit_time;
        double accuracy;

        StaticSettings() = default;
        PB_FN explicit StaticSettings(const Settings & s)
            : event(s.event), limit_energy(s.limit_energy),
              limit_distance(s.limit_distance),
              limit_grammage(s.limit_grammage), limit_time(s.limit_time),
              accuracy(s.accuracy)
        {
        }
        PB_FN static bool matches(const Settings & s)
        {
                return (s.energy_loss == ENERGY_LOSS) &&
                    (s.scattering == SCATTERING) &&
                    (s.direction == DIRECTION) && (s.decay == DECAY);
        }
};

} /* namespace pb */

#endif
