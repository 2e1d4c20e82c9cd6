/*
 * pb_hot.h -- the lean Monte Carlo step of the straggled energy loss mode in a
 * single uniform medium without locals (examples/pumas/straight.c geometry;
 * BASELINE.json headline, configs[0] and configs[2]).
 *
 * Same algorithm and same results, bit for bit, as the general restatement
 * (pb_transport.h: step_transport + the tail of stepping_step, i.e.
 * src/pumas.c:6296-6925, 6943-7095 and 5277-5335), re-organised for the GPU:
 *
 *  - Tables. The reference evaluates a dozen PCHIP columns per step, each from
 *    four scalars found in four different arrays (pumas.c:4146-4157). Here the
 *    columns a step needs are re-packed at load time into ONE record per grid
 *    cell (`HOTK_WORDS` doubles per cell of the kinetic grid, `HOTX_WORDS` per
 *    cell of the grammage column): the cell bounds and, per column, the four
 *    coefficients (p0, m0 * dx, c2, c3) of math_pchip_interpolate
 *    (pumas.c:4118-4125) evaluated by the host with the reference's own
 *    expressions and without contraction, so the device computes
 *    p0 + t (m0 + t (c2 + t c3)) from two 16 byte loads off one base address.
 *    Same IEEE operations on the same operands: identical values.
 *  - State. A track is a `slot` of HOT_WORDS 64 bit words behind an accessor
 *    (shared memory, word major, on the GPU; a plain array on the CPU); every
 *    field is read where it is used and written where it changes, so nothing
 *    but the working set of the arithmetic lives in registers.
 *  - The kinetic energy of a track is located once: the cell and reduced
 *    abscissa found for the end of step multiple scattering look-up are kept
 *    for the start of the next step.
 *  - Anything unusual (energies off the grid or in its first cell, above the
 *    1e9 GeV step reduction) is detected before the step draws a variate or
 *    writes the state, and is handed to the general step_transport.
 *
 * The number and order of the uniform variates drawn is unchanged (contract
 * with the oracle, tests/test_port_vs_reference.py runs this file on the CPU
 * against the reference).
 */
#ifndef PB_HOT_H
#define PB_HOT_H

#include <cstring>

#include "pb_hot_math.h"
#include "pb_hot_tables.h"
#include "pb_transport.h"

namespace pb {

/* ---- loads -----------------------------------------------------------------
 * A `HotPtr` addresses a word of the packed tables: a pointer (CPU build;
 * kernels that read them from HBM through L1), or, with PB_HOT_SMEM_TABLES (set
 * by hot_kernel.cuh), a 64 bit word index into the CTA's shared memory, where
 * the kernel stages the records of the medium's material (50 KB). */
struct D2 {
        double x, y;
};
struct D4 {
        double a, b, c, d;
};

#if defined(__CUDACC__) && defined(PB_HOT_SMEM_TABLES)
typedef unsigned int HotPtr;
extern __shared__ __align__(16) unsigned long long hot_smem64[];

PB_FN double hot_ld1(HotPtr p)
{
#if defined(__CUDA_ARCH__)
        return __longlong_as_double((long long)hot_smem64[p]);
#else
        (void)p;
        return 0.;
#endif
}

PB_FN D2 hot_ld2(HotPtr p)
{
        D2 r;
#if defined(__CUDA_ARCH__)
        const double2 v = *reinterpret_cast<const double2 *>(&hot_smem64[p]);
        r.x = v.x;
        r.y = v.y;
#else
        (void)p;
        r.x = r.y = 0.;
#endif
        return r;
}

PB_FN D4 hot_ld4(HotPtr p)
{
        D4 r;
#if defined(__CUDA_ARCH__)
        const double2 u = *reinterpret_cast<const double2 *>(&hot_smem64[p]);
        const double2 v =
            *reinterpret_cast<const double2 *>(&hot_smem64[p + 2]);
        r.a = u.x;
        r.b = u.y;
        r.c = v.x;
        r.d = v.y;
#else
        (void)p;
        r.a = r.b = r.c = r.d = 0.;
#endif
        return r;
}
#else
typedef const double * HotPtr;

PB_FN double hot_ld1(HotPtr p)
{
#if defined(__CUDA_ARCH__)
        return __ldg(p);
#else
        return *p;
#endif
}

PB_FN D2 hot_ld2(HotPtr p)
{
        D2 r;
#if defined(__CUDA_ARCH__)
        const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
        r.x = v.x;
        r.y = v.y;
#else
        r.x = p[0];
        r.y = p[1];
#endif
        return r;
}

PB_FN D4 hot_ld4(HotPtr p)
{
        D4 r;
#if defined(__CUDA_ARCH__)
        const double2 u = __ldg(reinterpret_cast<const double2 *>(p));
        const double2 v = __ldg(reinterpret_cast<const double2 *>(p) + 1);
        r.a = u.x;
        r.b = u.y;
        r.c = v.x;
        r.d = v.y;
#else
        r.a = p[0];
        r.b = p[1];
        r.c = p[2];
        r.d = p[3];
#endif
        return r;
}
#endif

/* A packed column at reduced abscissa t: the polynomial of pumas.c:4124. */
PB_FN double hot_pchip(HotPtr column, double t)
{
        const D4 c = hot_ld4(column);
        return c.a + t * (c.b + t * (c.c + t * c.d));
}

/* ---- per launch constants ------------------------------------------------- */
struct HotCtx {
        HotPtr hotK; /* K records of the medium's material */
        HotPtr hotX; /* X records of the same */
        /* the same in HBM (the kernels that stage them to shared memory) */
        const double * gK;
        const double * gX;
        int n_cells;         /* n_energies - 1 */
        int material;
        double K0, K1, Kmax; /* K[0], K[1], K[n_energies - 1] */
        double X0, Xmax;     /* grammage column: first and last node */
        double Ms1_1, Lb_1;  /* Ms1[1], Lb[1]: below K[1], pumas.c:4088, 7453 */
        double density, density_i;
        double mass, ctau;
        double step_X_limit; /* pumas.c:5262-5266, constant in one material */
        double grammage_max; /* transport_limit, pumas.c:5612-5617: constant in
                              * the straggled mode */
};

template <class Cfg>
inline HotCtx hot_context(const Physics & host_view, const Physics & view,
    const double * hotK, const double * hotX, const Cfg & cfg, int material,
    double density = 0.)
{
        /* `host_view` is read here; `view`, hotK, hotX are the pointers the
         * step will use (device or host). */
        (void)view;
        HotCtx h;
        const Physics & p = host_view;
        const int nc = p.n_energies - 1;
        h.gK = hotK + (size_t)material * nc * HOTK_WORDS;
        h.gX = hotX + (size_t)material * nc * HOTX_WORDS;
#if defined(__CUDACC__) && defined(PB_HOT_SMEM_TABLES)
        h.hotK = h.hotX = 0; /* (set by the launcher: shared memory layout) */
#else
        h.hotK = h.gK;
        h.hotX = h.gX;
#endif
        h.n_cells = nc;
        h.material = material;
        h.K0 = p.K[0];
        h.K1 = p.K[1];
        h.Kmax = p.K[nc];
        const double * X = col_sm(p, p.X, MODE_MIXED, material);
        h.X0 = X[0];
        h.Xmax = X[nc];
        h.Ms1_1 = col_Ms1(p, p.Ms1, MODE_MIXED, material)[1];
        h.Lb_1 = col_m(p, p.Lb, material)[1];
        h.density = (density > 0.) ? density : p.material_density[material];
        h.density_i = 1. / h.density;
        h.mass = p.mass;
        h.ctau = p.ctau;
        h.step_X_limit = (cfg.event & EVENT_LIMIT_ENERGY) ?
            cel_grammage(p, cfg.energy_loss, material, cfg.limit_energy) :
            0.;
        h.grammage_max =
            (cfg.event & EVENT_LIMIT_GRAMMAGE) ? cfg.limit_grammage : 0.;
        return h;
}

/* ---- the slot ------------------------------------------------------------- */
enum {
        HF_ENERGY = 0,
        HF_DISTANCE,
        HF_GRAMMAGE,
        HF_TIME,
        HF_PX,
        HF_PY,
        HF_PZ,
        HF_UX,
        HF_UY,
        HF_UZ,
        HF_INVLB1, /* Track::step_invlb1 */
        HF_RANDN,  /* Track::randn_next */
        HF_T0,     /* reduced abscissa of the energy in cell iK, when located */
        HF_FLAGS,  /* packed flags (low) and n_steps (high) */
        HF_ID,     /* position in the batch (low) and variates drawn (high) */
        /* the words only events touch (1 step in 6): the kernel keeps them in
         * the HBM mirror of the slot instead of shared memory */
        HOT_RESIDENT_WORDS,
        HF_WEIGHT = HOT_RESIDENT_WORDS,
        HF_TI,     /* Loop::ti */
        HF_WI,     /* Loop::wi */
        HF_XF,     /* Loop::Xf  (backward weights) */
        HF_DEI,    /* Loop::dei (backward weights) */
        HOT_WORDS
};

/* flags word, low 32 bits */
constexpr uint32_t HFL_EVENT_MASK = 0x3fffu;      /* Track::step_event */
constexpr uint32_t HFL_FIRST = 1u << 14;          /* Track::step_first */
constexpr uint32_t HFL_RANDN = 1u << 15;          /* Track::randn_done */
constexpr uint32_t HFL_LOCATED = 1u << 16;        /* HF_T0 / iK are valid */
constexpr int HFL_IK_SHIFT = 17;                  /* Track::iK, 8 bits */
constexpr uint32_t HFL_IK_MASK = 0xffu << HFL_IK_SHIFT;
constexpr int HFL_PHASE_SHIFT = 25;               /* scheduler phase, 3 bits */
constexpr uint32_t HFL_PHASE_MASK = 7u << HFL_PHASE_SHIFT;
constexpr int HFL_RC_SHIFT = 28;                  /* Track::error != 0 */

/* A slot in plain memory (CPU build, and the service code of the kernel). */
struct PlainSlot {
        unsigned long long w[HOT_WORDS];
        PB_FN double ld(int k) const
        {
                double d;
                const unsigned long long u = w[k];
                memcpy(&d, &u, sizeof d);
                return d;
        }
        PB_FN void st(int k, double v)
        {
                unsigned long long u;
                memcpy(&u, &v, sizeof u);
                w[k] = u;
        }
        PB_FN unsigned long long ldu(int k) const { return w[k]; }
        PB_FN void stu(int k, unsigned long long v) { w[k] = v; }
};

/* (State, Track, Loop) <-> slot. `index` and `drawn` are kept by the caller
 * in HF_ID. Only what a straggled track in a uniform medium needs: the rest of
 * Track and Loop is constant there (loop_canonical, pb_transport.h). */
template <class S>
PB_FN void hot_pack(S & s, const State & st, const Track & t, const Loop & l,
    int phase, bool located, double t0)
{
        s.st(HF_ENERGY, st.energy);
        s.st(HF_DISTANCE, st.distance);
        s.st(HF_GRAMMAGE, st.grammage);
        s.st(HF_TIME, st.time);
        s.st(HF_WEIGHT, st.weight);
        s.st(HF_PX, st.position[0]);
        s.st(HF_PY, st.position[1]);
        s.st(HF_PZ, st.position[2]);
        s.st(HF_UX, st.direction[0]);
        s.st(HF_UY, st.direction[1]);
        s.st(HF_UZ, st.direction[2]);
        s.st(HF_INVLB1, t.step_invlb1);
        s.st(HF_RANDN, t.randn_next);
        s.st(HF_TI, l.ti);
        s.st(HF_WI, l.wi);
        s.st(HF_XF, l.Xf);
        s.st(HF_DEI, l.dei);
        s.st(HF_T0, t0);
        const uint32_t f = ((uint32_t)t.step_event & HFL_EVENT_MASK) |
            (t.step_first ? HFL_FIRST : 0u) | (t.randn_done ? HFL_RANDN : 0u) |
            (located ? HFL_LOCATED : 0u) |
            (((uint32_t)t.iK & 0xffu) << HFL_IK_SHIFT) |
            (((uint32_t)phase & 7u) << HFL_PHASE_SHIFT) |
            (((uint32_t)t.error & 0xfu) << HFL_RC_SHIFT);
        s.stu(HF_FLAGS,
            (unsigned long long)f | ((unsigned long long)t.n_steps << 32));
}

template <class S, class Cfg>
PB_FN void hot_unpack(const S & s, const Physics & p, const Cfg & cfg,
    const UniformGeometry & geo, State & st, Track & t, Loop & l)
{
        st.energy = s.ld(HF_ENERGY);
        st.distance = s.ld(HF_DISTANCE);
        st.grammage = s.ld(HF_GRAMMAGE);
        st.time = s.ld(HF_TIME);
        st.weight = s.ld(HF_WEIGHT);
        st.position[0] = s.ld(HF_PX);
        st.position[1] = s.ld(HF_PY);
        st.position[2] = s.ld(HF_PZ);
        st.direction[0] = s.ld(HF_UX);
        st.direction[1] = s.ld(HF_UY);
        st.direction[2] = s.ld(HF_UZ);
        st.charge = 0.; /* (not used by the stepping; never written back) */
        st.decayed = 0;
        const unsigned long long fw = s.ldu(HF_FLAGS);
        const uint32_t f = (uint32_t)fw;
        t.lifetime = 0.;
        t.step_first = (f & HFL_FIRST) ? 1 : 0;
        t.step_event = (int)(f & HFL_EVENT_MASK);
        t.step_foreseen =
            (cfg.event & EVENT_LIMIT_GRAMMAGE) ? EVENT_LIMIT_GRAMMAGE : EVENT_NONE;
        t.step_X_limit = 0.; /* (set by the caller from HotCtx) */
        t.step_invlb1 = s.ld(HF_INVLB1);
        t.randn_done = (f & HFL_RANDN) ? 1 : 0;
        t.randn_next = s.ld(HF_RANDN);
        t.iK = (int)((f >> HFL_IK_SHIFT) & 0xffu);
        t.error = (int)((f >> HFL_RC_SHIFT) & 0xfu);
        t.n_steps = (unsigned int)(fw >> 32);
        l.medium = 0;
        l.first = 0;
        l.ti = s.ld(HF_TI);
        l.wi = s.ld(HF_WI);
        l.Xi = st.grammage;
        l.Xf = s.ld(HF_XF);
        l.dei = s.ld(HF_DEI);
        l.grammage_max =
            (cfg.event & EVENT_LIMIT_GRAMMAGE) ? cfg.limit_grammage : 0.;
        loop_canonical(p, cfg, geo, t, l);
}

/* ---- located look-ups ------------------------------------------------------ */
struct HLoc {
        int i;    /* cell, 0 <= i < n_cells */
        double t; /* (x - X[i]) / dX */
};

/* Cell of x in the kinetic grid, walking from `hint`. Requires K0 <= x < Kmax
 * (the caller checks): on a strictly increasing grid this is the cell the
 * reference's bisection returns (pumas.c:4174-4233). */
PB_HOTCALL HLoc hot_locate_K(const HotCtx & h, double x, int hint)
{
        int i = hint;
        if (i < 0) i = 0;
        if (i > h.n_cells - 1) i = h.n_cells - 1;
        D2 b = hot_ld2(h.hotK + i * HOTK_WORDS);
        while (x < b.x) {
                --i;
                b = hot_ld2(h.hotK + i * HOTK_WORDS);
        }
        while (x >= b.y) {
                ++i;
                b = hot_ld2(h.hotK + i * HOTK_WORDS);
        }
        HLoc l;
        l.i = i;
        l.t = hot_div(x - b.x, b.y - b.x);
        return l;
}

/* Same in the grammage column. Requires X0 <= x < Xmax. */
PB_HOTCALL HLoc hot_locate_X(const HotCtx & h, double x, int hint)
{
        int i = hint;
        if (i < 0) i = 0;
        if (i > h.n_cells - 1) i = h.n_cells - 1;
        D2 b = hot_ld2(h.hotX + i * HOTX_WORDS);
        while (x < b.x) {
                --i;
                b = hot_ld2(h.hotX + i * HOTX_WORDS);
        }
        while (x >= b.y) {
                ++i;
                b = hot_ld2(h.hotX + i * HOTX_WORDS);
        }
        HLoc l;
        l.i = i;
        l.t = hot_div(x - b.x, b.y - b.x);
        return l;
}

PB_FN HotPtr hot_K(const HotCtx & h, const HLoc & l)
{
        return h.hotK + l.i * HOTK_WORDS;
}

/* cel_kinetic_energy (pumas.c:3758-3784) of a grammage inside the column. */
PB_FN double hot_kinetic_of_grammage(const HotCtx & h, double X, int hint)
{
        const HLoc l = hot_locate_X(h, X, hint);
        return hot_pchip(h.hotX + l.i * HOTX_WORDS + HX_K, l.t);
}

/* coulomb_ehs_length (pumas.c:7446-7466) at an energy located in the grid. */
PB_FN double hot_ehs_length(
    const HotCtx & h, const HLoc & l, double kinetic)
{
        const double p2 = kinetic * (kinetic + 2. * h.mass);
        if (kinetic < h.K1) return hot_div(h.Lb_1, p2);
        const D2 lb = hot_ld2(hot_K(h, l) + HK_LB);
        return hot_div(lb.x + l.t * lb.y, p2);
}

/* Inverse first transport grammage, table_get_msc (pumas.c:4077-4111), at an
 * energy located in the grid. */
PB_FN double hot_msc_invlb1(const HotCtx & h, const HLoc & l, double kinetic)
{
        if (kinetic < h.K1)
                return h.Ms1_1 * hot_sqrt(hot_div(h.K1, kinetic));
        return hot_pchip(hot_K(h, l) + HK_MS1, l.t);
}

/* Is the energy inside the grid? */
PB_FN bool hot_in_grid(const HotCtx & h, double k)
{
        return (k >= h.K0) && (k < h.Kmax);
}

/* Is the energy one the lean step handles: inside the grid and below the
 * 1e9 GeV step reduction (pumas.c:6331). The stopping tracks spend a tenth of
 * all steps below K[1] = 1 MeV (steps of STEP_MIN): the lean step covers the
 * first cell, with the switches of table_get_msc and coulomb_ehs_length at
 * K[1] (pumas.c:4088, 7453). */
PB_FN bool hot_energy_ok(const HotCtx & h, double k)
{
        return hot_in_grid(h, k) && !(k > 1E+09);
}

/* The thinning candidates of a step (pumas.c:6606-6622, 6637-6660): kinetic
 * energy at the grammage of the candidate vertex, located in the grid when it
 * lies there. One out of line copy for the two call sites; the look-ups walk
 * from the cell of the start energy (the candidate lies within the step). */
/* In line since the loop's footprint stopped being what bounds the kernel
 * (profiles/r2_experiments.md sections 12, 14): a call saves and restores the
 * live registers through local memory. -DPB_HOT_OUTLINE_CALLS: the one-copy
 * versions of rounds 1 and 2 (A/B). */
#if defined(PB_HOT_OUTLINE_CALLS)
#define PB_HOT_VERTEX_CALL PB_NOINLINE
#else
#define PB_HOT_VERTEX_CALL PB_FN
#endif
struct HotVertex {
        double k;
        HLoc l;
        int located; /* K[0] <= k < Kmax: `l` is valid */
};

template <class Cfg>
PB_HOT_VERTEX_CALL HotVertex hot_vertex(const Physics & p, const Cfg & cfg,
    const HotCtx & h, double energy, double Xtot, double X_ev, double ratio,
    int hint)
{
        const double sgn = (cfg.direction == MODE_FORWARD) ? 1. : -1.;
        const double X = Xtot - sgn * X_ev;
        double k_h;
        if ((X >= h.X0) && (X < h.Xmax))
                k_h = hot_kinetic_of_grammage(h, X, hint);
        else
                k_h = cel_kinetic_energy(p, cfg.energy_loss, h.material, X);
        HotVertex v;
        v.k = energy - sgn * ratio * fabs(energy - k_h);
        v.l.i = hint;
        v.l.t = 0.;
        v.located = hot_in_grid(h, v.k) ? 1 : 0;
        if (v.located) v.l = hot_locate_K(h, v.k, hint);
        return v;
}

/* General look-ups for energies off the grid (rare), out of line. */
PB_NOINLINE double hot_cel_straggling(
    const Physics & p, int material, double kinetic)
{
        return cel_straggling(p, material, kinetic);
}

PB_NOINLINE double hot_cel_energy_loss(
    const Physics & p, int scheme, int material, double kinetic)
{
        return cel_energy_loss(p, scheme, material, kinetic);
}

PB_NOINLINE double hot_table_invlb1(
    const Physics & p, int scheme, int material, double kinetic)
{
        double mu0, v;
        table_get_msc(p, scheme, material, kinetic, &mu0, &v);
        return v;
}

PB_NOINLINE double hot_del_cross_section(
    const Physics & p, int material, double kinetic)
{
        return del_cross_section(p, material, kinetic);
}

PB_NOINLINE double hot_coulomb_ehs_length(
    const Physics & p, int material, double kinetic)
{
        return coulomb_ehs_length(p, material, kinetic);
}

/* The random stream of the lean step (`Rng`): besides uniform(), which may
 * generate a block when its buffer is empty,
 *   uniform_fast(), log_uniform_fast()   no check: the caller of hot_step
 *       guarantees 8 buffered variates, what a step draws outside its retry
 *       loops;
 *   uniform_keep(), log_uniform_keep()   for the retry loops: refill so that 8
 *       stay buffered.
 * log_uniform*() return the logarithm of the next variate (pb_hot_math.h on
 * the GPU, the host libm on the CPU). */

/* step_randn, pumas.c:7016-7029, on the slot's Box-Muller cache. */
template <bool KEEP, class Rng>
PB_FN double hot_randn(Rng & rng, uint32_t & flags, double & randn_next)
{
        flags ^= HFL_RANDN;
        if (!(flags & HFL_RANDN)) return randn_next;
        const double r = hot_sqrt(
            -2. * (KEEP ? rng.log_uniform_keep() : rng.log_uniform_fast()));
        /* sin, cos of 2 pi u */
        const SinCos sc = hot_sincospi(
            2. * (KEEP ? rng.uniform_keep() : rng.uniform_fast()));
        randn_next = r * sc.c;
        return r * sc.s;
}

/* The general restatement of the step on the unpacked track (rare). */
template <class Cfg, class S, class Rng>
PB_NOINLINE int hot_step_general(const Physics & p, const Cfg & cfg,
    const UniformGeometry & geo, const HotCtx & h, S & s, Rng & rng)
{
        State st;
        Track t;
        Loop l;
        hot_unpack(s, p, cfg, geo, st, t, l);
        t.step_X_limit = h.step_X_limit;
        const int phase = stepping_step(p, cfg, geo, t, rng, st, l);
        hot_pack(s, st, t, l, phase, false, 0.);
        return phase;
}

/* A proper time limit reached inside a step, pumas.c:6762-6797 (uniform
 * density). Out of line: only met with PUMAS_EVENT_LIMIT_TIME. */
struct HotTimeLimit {
        double step, k1, grammage;
};
PB_NOINLINE HotTimeLimit hot_time_limit(double momentum, double p_f, double sf1,
    double time_max, double ti, double mass, double density, double grammage0,
    double dX)
{
        HotTimeLimit r;
        const double a = (momentum / p_f - 1.) / sf1;
        const double c = -2. * (time_max - ti) * momentum / mass;
        double step;
        if (a != 0.) {
                double delta = 1. - a * c;
                if (delta < 0.) delta = 0.;
                step = 1. / a * (sqrt(delta) - 1.);
        } else {
                step = -0.5 * c;
        }
        if (step < 0.) step = 0.;
        const double p1 = momentum / (1 + a * step);
        double k1 = sqrt(p1 * p1 - mass * mass) - mass;
        if (k1 < 0.) k1 = 0.;
        const double Xf = grammage0 + dX;
        double grammage = grammage0 + step * density;
        if (grammage > Xf) grammage = Xf;
        r.step = step;
        r.k1 = k1;
        r.grammage = grammage;
        return r;
}

/*
 * One Monte Carlo step of a track in PHASE_STEP, then the event dispatch of
 * the stepping loop (pumas.c:5277-5335). Returns the next phase. The slot is
 * updated in place. `p`, `cfg`, `geo` serve the general fall-backs.
 */
template <class Cfg, class S, class Rng>
PB_FN int hot_step(const Physics & p, const Cfg & cfg,
    const UniformGeometry & geo, const HotCtx & h, S & s, Rng & rng)
{
        static_assert(Cfg::energy_loss == MODE_STRAGGLED, "straggled only");
        static_assert(Cfg::decay == MODE_WEIGHTED, "weighted decays only");
        constexpr bool mixed = (Cfg::scattering == MODE_MIXED);
        constexpr bool forward = (Cfg::direction == MODE_FORWARD);
        const double sgn = forward ? 1. : -1.;

        const double e0 = s.ld(HF_ENERGY);
        const unsigned long long fw = s.ldu(HF_FLAGS);
        uint32_t flags = (uint32_t)fw;

        /* ---- anything unusual: the general step ------------------------- */
        bool lean = hot_energy_ok(h, e0);
        /* (the end of the CSDA step must stay inside the grammage column;
         * checked below, still before any variate is drawn) */
        HLoc l0;
        double Xtot = 0., step = 0., invlb1 = 0., dX = 0., Xend = 0.;
        int event = EVENT_NONE;
        if (lean) {
                /* the energy of the track in the grid: kept from the end of
                 * the last step, or located now */
                l0.i = (int)((flags >> HFL_IK_SHIFT) & 0xffu);
                if (flags & HFL_LOCATED)
                        l0.t = s.ld(HF_T0);
                else
                        l0 = hot_locate_K(h, e0, l0.i);
                const HotPtr r0 = hot_K(h, l0);
                Xtot = hot_pchip(r0 + HK_X, l0.t); /* pumas.c:6318 */

                /* local step length, pumas.c:6324-6347, 6386-6395 */
                double step_loc = cfg.accuracy * h.density_i * Xtot;
                if (mixed) {
                        if (flags & HFL_FIRST)
                                invlb1 =
                                    hot_msc_invlb1(h, l0, e0) * h.density;
                        else
                                invlb1 = s.ld(HF_INVLB1);
                        const double stepT = hot_div(cfg.accuracy, invlb1);
                        if (stepT < step_loc) step_loc = stepT;
                }
                if (step_loc < STEP_MIN) step_loc = STEP_MIN;
                step = step_loc;
                if (cfg.event & EVENT_LIMIT_DISTANCE) {
                        const double d =
                            cfg.limit_distance - s.ld(HF_DISTANCE);
                        if (d <= step) {
                                step = d;
                                event = EVENT_LIMIT_DISTANCE;
                        }
                }
                dX = 0.5 * step * (h.density + h.density); /* pumas.c:6518 */
                Xend = Xtot - sgn * dX;
                lean = (Xend < h.Xmax);
        }
        if (!lean) return hot_step_general(p, cfg, geo, h, s, rng);

        /* ---- the lean step ------------------------------------------------ */
        flags &= ~HFL_FIRST;
        const double momentum = hot_sqrt(e0 * (e0 + 2. * h.mass));

        /* position, pumas.c:6410-6414 */
        {
                const double a = step * sgn;
                s.st(HF_PX, s.ld(HF_PX) + a * s.ld(HF_UX));
                s.st(HF_PY, s.ld(HF_PY) + a * s.ld(HF_UY));
                s.st(HF_PZ, s.ld(HF_PZ) + a * s.ld(HF_UZ));
        }

        /* straggled end of step energy: step_fluctuate, pumas.c:6943-7005 */
        double k1, ratio = 0.;
        const HotPtr r0 = hot_K(h, l0);
        HLoc l1 = { 0, 0. }; /* the located end energy, when `have_l1` */
        bool have_l1 = false;
        {
                /* (below the column: 0, pumas.c:3765) */
                k1 = (Xend >= h.X0) ?
                    hot_kinetic_of_grammage(h, Xend, l0.i) :
                    0.;
                const double dk0 = fabs(e0 - k1);
                if ((k1 > 0.) && (dk0 > 0.)) {
                        double Omega1, de1 = 0.;
                        const double tmp = hot_div(dX, Xtot);
                        if ((k1 >= h.K0) && (k1 < h.Kmax)) {
                                const HLoc lc = hot_locate_K(h, k1, l0.i);
                                const HotPtr rc = hot_K(h, lc);
                                Omega1 = hot_pchip(rc + HK_OMEGA, lc.t);
                                if (tmp > 5E-02)
                                        de1 = hot_pchip(rc + HK_DE, lc.t);
                        } else {
                                Omega1 =
                                    hot_cel_straggling(p, h.material, k1);
                                if (tmp > 5E-02)
                                        de1 = hot_cel_energy_loss(p,
                                            cfg.energy_loss, h.material, k1);
                        }
                        const double Omega0 = hot_pchip(r0 + HK_OMEGA, l0.t);
                        double dk12 = 0.5 * dX * (Omega0 + Omega1);
                        if (tmp > 5E-02) {
                                const double de0 =
                                    hot_pchip(r0 + HK_DE, l0.t);
                                dk12 *= 1. +
                                    hot_div(sgn * dX, dk0) * (de1 - de0);
                        }
                        const double dk1 = hot_sqrt(dk12);
                        if (dk0 >= 3. * dk1) {
                                double randn_next = s.ld(HF_RANDN);
                                double u =
                                    hot_randn<false>(rng, flags, randn_next);
                                while (fabs(u) > 3.)
                                        u = hot_randn<true>(
                                            rng, flags, randn_next);
                                s.st(HF_RANDN, randn_next);
                                u /= 1.015387;
                                k1 += u * dk1;
                        } else if (dk0 >= 1.7320508 * dk1) {
                                const double u =
                                    1.7320508 * (1. - 2. * rng.uniform_fast());
                                k1 += u * dk1;
                        } else {
                                const double dk32 = 3. * dk12;
                                const double dk02 = dk0 * dk0;
                                const double a =
                                    1. - hot_div(dk32 - dk02, dk32 + 3. * dk02);
                                if (rng.uniform_fast() <= a) {
                                        const double b =
                                            hot_div(0.5 * (dk32 + 3. * dk02), dk0);
                                        const double u = rng.uniform_fast();
                                        k1 = e0 - sgn * b * u;
                                        if (k1 < 0.) k1 = 0.;
                                } else {
                                        k1 = e0;
                                }
                        }
                        ratio = hot_div(fabs(e0 - k1), dk0);
                }
                if (ratio == 0.) ratio = 1.;
        }
        const double dk = fabs(e0 - k1);

        /* kinetic limit, pumas.c:6573-6596 */
        double grammage_max = h.grammage_max;
        const double grammage0 = s.ld(HF_GRAMMAGE);
        {
                double kinetic_limit = -1.;
                if (forward) {
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
                        const double grammage = grammage0 +
                            hot_div(fabs(e0 - kinetic_limit), dk) * dX;
                        k1 = kinetic_limit;
                        if ((grammage_max <= 0.) ||
                            (grammage < grammage_max)) {
                                grammage_max = grammage;
                                event = EVENT_LIMIT_ENERGY;
                        }
                }
        }

        /* discrete events by thinning, pumas.c:6598-6669 */
        double Xmax = dX;
        if (grammage_max > 0.) {
                const double dX1 = grammage_max - grammage0;
                if (Xmax > dX1) Xmax = dX1;
        }
        double k_x = -1.;

        /* the end energy in the grid, for the cross-sections below and the
         * end of step scattering length */
        if ((k1 >= h.K0) && (k1 < h.Kmax)) {
                l1 = hot_locate_K(h, k1, l0.i);
                have_l1 = true;
        }
        {
                double xs_del = hot_pchip(r0 + HK_CS, l0.t);
                const double tmp_del = have_l1 ?
                    hot_pchip(hot_K(h, l1) + HK_CS, l1.t) :
                    hot_del_cross_section(p, h.material, k1);
                if (tmp_del > xs_del) xs_del = tmp_del;
                if (xs_del > 0.) {
                        const double X_del =
                            hot_div(-rng.log_uniform_fast(), xs_del);
                        if (X_del < Xmax) {
                                const HotVertex v = hot_vertex(p, cfg, h, e0,
                                    Xtot, X_del, ratio, l0.i);
                                const double k_del = v.k;
                                if (k_del > 0.) {
                                        const double r = hot_div(
                                            v.located ?
                                                hot_pchip(
                                                    hot_K(h, v.l) + HK_CS,
                                                    v.l.t) :
                                                hot_del_cross_section(
                                                    p, h.material, k_del),
                                            xs_del);
                                        if (!(rng.uniform_fast() > r)) {
                                                k_x = k_del;
                                                Xmax = X_del;
                                                event = EVENT_VERTEX_DEL;
                                        }
                                }
                        }
                }
        }
        if (mixed) {
                double kmin, kmax;
                if (k1 <= e0)
                        kmin = k1, kmax = e0;
                else
                        kmax = k1, kmin = e0;
                if (kmin < h.K1) {
                        const double tmp = kmin;
                        kmin = kmax, kmax = tmp;
                }
                (void)kmax;
                /* kmin is e0 or k1: both located when inside the grid */
                double lb_ehs;
                if (kmin == e0)
                        lb_ehs = hot_ehs_length(h, l0, kmin);
                else if (have_l1)
                        lb_ehs = hot_ehs_length(h, l1, kmin);
                else
                        lb_ehs = hot_coulomb_ehs_length(p, h.material, kmin);
                if (lb_ehs > 0.) {
                        const double X_ehs = -rng.log_uniform_fast() * lb_ehs;
                        if (X_ehs < Xmax) {
                                const HotVertex v = hot_vertex(p, cfg, h, e0,
                                    Xtot, X_ehs, ratio, l0.i);
                                const double k_ehs = v.k;
                                if (k_ehs > 0.) {
                                        const double r = hot_div(
                                            v.located ?
                                                hot_ehs_length(
                                                    h, v.l, k_ehs) :
                                                hot_coulomb_ehs_length(
                                                    p, h.material, k_ehs),
                                            lb_ehs);
                                        if (!((r <= 0.) ||
                                                (rng.uniform_fast() >
                                                    hot_div(1., r)))) {
                                                k_x = k_ehs;
                                                Xmax = X_ehs;
                                                event = EVENT_VERTEX_COULOMB;
                                        }
                                }
                        }
                }
        }
        if (k_x >= 0.) {
                k1 = k_x;
                grammage_max = grammage0 + Xmax;
                have_l1 = false;
        }

        /* grammage, pumas.c:6671-6713 (uniform density) */
        double grammage = grammage0 + dX;
        const double sf0 = step;
        if ((grammage_max > 0.) && (grammage >= grammage_max)) {
                const double dX_ = grammage_max - grammage0;
                step = dX_ * h.density_i;
                grammage = grammage_max;
                if (!(event & EVENT_LIMIT_ENERGY)) {
                        if (!(event &
                                (EVENT_VERTEX_COULOMB | EVENT_VERTEX_DEL))) {
                                k1 = e0 -
                                    hot_div(sgn * (grammage - grammage0) * dk,
                                        dX);
                                if (k1 < 0.) k1 = 0.;
                                event = EVENT_LIMIT_GRAMMAGE;
                                have_l1 = false;
                        }
                }
                const double a = (step - sf0) * sgn;
                s.st(HF_PX, s.ld(HF_PX) + a * s.ld(HF_UX));
                s.st(HF_PY, s.ld(HF_PY) + a * s.ld(HF_UY));
                s.st(HF_PZ, s.ld(HF_PZ) + a * s.ld(HF_UZ));
        }

        /* proper time, pumas.c:6715-6798 (weighted decays: no lifetime) */
        const double time_max =
            (cfg.event & EVENT_LIMIT_TIME) ? cfg.limit_time : 0.;
        const double sf1 = step;
        {
                const double p_f =
                    (k1 <= 0) ? momentum : hot_sqrt(k1 * (k1 + 2. * h.mass));
                const double ti = s.ld(HF_TIME);
                double time =
                    ti +
                    0.5 * step * h.mass *
                        (hot_div(1., momentum) + hot_div(1., p_f));
                if ((time_max > 0.) && (time >= time_max)) {
                        event = EVENT_LIMIT_TIME;
                        time = time_max;
                        const HotTimeLimit tl = hot_time_limit(momentum, p_f,
                            sf1, time_max, ti, h.mass, h.density, grammage0,
                            dX);
                        step = tl.step;
                        k1 = tl.k1;
                        grammage = tl.grammage;
                        have_l1 = false;

                        const double b = (step - sf1) * sgn;
                        s.st(HF_PX, s.ld(HF_PX) + b * s.ld(HF_UX));
                        s.st(HF_PY, s.ld(HF_PY) + b * s.ld(HF_UY));
                        s.st(HF_PZ, s.ld(HF_PZ) + b * s.ld(HF_UZ));
                }
                s.st(HF_TIME, time);
        }
        s.st(HF_GRAMMAGE, grammage);
        s.st(HF_ENERGY, k1);
        s.st(HF_DISTANCE, s.ld(HF_DISTANCE) + step);

        /* end of step scattering length, pumas.c:6808-6818; the located
         * energy is kept for the next step */
        double invlb1_end = 0.;
        flags &= ~HFL_LOCATED;
        if (mixed) {
                if (k1 <= 0.) {
                        invlb1_end = invlb1;
                } else {
                        if (!have_l1 && (k1 >= h.K0) && (k1 < h.Kmax)) {
                                l1 = hot_locate_K(h, k1, l0.i);
                                have_l1 = true;
                        }
                        if (have_l1) {
                                invlb1_end =
                                    h.density * hot_msc_invlb1(h, l1, k1);
                        } else {
                                invlb1_end = h.density *
                                    hot_table_invlb1(p, cfg.energy_loss,
                                        h.material, k1);
                        }
                }
                s.st(HF_INVLB1, invlb1_end);
        } else if (!have_l1 && (k1 >= h.K0) && (k1 < h.Kmax)) {
                l1 = hot_locate_K(h, k1, l0.i);
                have_l1 = true;
        }
        int iK = l0.i;
        if (have_l1) {
                iK = l1.i;
                flags |= HFL_LOCATED;
                s.st(HF_T0, l1.t);
        }
        flags = (flags & ~HFL_IK_MASK) | ((uint32_t)iK << HFL_IK_SHIFT);

        /* soft multiple scattering, pumas.c:6913-6922 */
        if (mixed && ((invlb1 > 0.) || (invlb1_end > 0.))) {
                double ilb1 = 0.25 * step * (invlb1 + invlb1_end);
                if (ilb1 > 1.) ilb1 = 1.;
                double mu = -ilb1 * rng.log_uniform_fast();
                while (mu > 1.) mu = -ilb1 * rng.log_uniform_keep();
                /* step_rotate_direction, pumas.c:7042-7095 */
                const double cos_theta = 1 - 2 * mu;
                const double stsq = 4 * mu * (1. - mu);
                if (stsq > 0.) {
                        const double st = hot_sqrt(stsq);
                        const double d0 = s.ld(HF_UX), d1 = s.ld(HF_UY),
                                     d2 = s.ld(HF_UZ);
                        const double a0 = fabs(d0), a1 = fabs(d1),
                                     a2 = fabs(d2);
                        const bool case_xz = (a0 > a1) && (a0 > a2);
                        const bool case_xy = !(a0 > a1) && (a1 > a2);
                        const double pa = (case_xz || case_xy) ? d0 : d1;
                        const double pb_ = case_xy ? d1 : d2;
                        const double nrm =
                            hot_div(1., hot_sqrt(pa * pa + pb_ * pb_));
                        const double u0x = case_xz ? -d2 * nrm :
                            (case_xy ? d1 * nrm : 0.);
                        const double u0y = case_xz ? 0. :
                            (case_xy ? -d0 * nrm : d2 * nrm);
                        const double u0z = case_xz ? d0 * nrm :
                            (case_xy ? 0. : -d1 * nrm);
                        const double u1x = u0y * d2 - u0z * d1;
                        const double u1y = u0z * d0 - u0x * d2;
                        const double u1z = u0x * d1 - u0y * d0;
                        /* sin, cos of phi = pi (1 - 2 u) */
                        const SinCos sc =
                            hot_sincospi(1. - 2. * rng.uniform_fast());
                        s.st(HF_UX,
                            cos_theta * d0 + st * (sc.c * u0x + sc.s * u1x));
                        s.st(HF_UY,
                            cos_theta * d1 + st * (sc.c * u0y + sc.s * u1y));
                        s.st(HF_UZ,
                            cos_theta * d2 + st * (sc.c * u0z + sc.s * u1z));
                }
        }

        /* ---- event dispatch, pumas.c:5297-5335 ----------------------------- */
        const unsigned int n_steps = (unsigned int)(fw >> 32) + 1u;
        int phase = PHASE_STEP;
        if (event != EVENT_NONE) {
                /* weights, pumas.c:5309-5321 */
                const double wt = hot_exp(
                    hot_div(-fabs(s.ld(HF_TIME) - s.ld(HF_TI)), h.ctau));
                if (!forward) {
                        double de;
                        if (have_l1)
                                de = hot_pchip(hot_K(h, l1) + HK_DE, l1.t);
                        else
                                de = hot_cel_energy_loss(
                                    p, cfg.energy_loss, h.material, k1);
                        s.st(HF_WEIGHT, s.ld(HF_WI) * wt * s.ld(HF_DEI) * de);
                } else {
                        s.st(HF_WEIGHT, s.ld(HF_WI) * wt);
                }
                if (event & (EVENT_LIMIT | EVENT_WEIGHT | EVENT_VERTEX_DECAY)) {
                        /* stepping_finish, pumas.c:5492-5513 */
                        if (forward) {
                                const double kinetic_min =
                                    (cfg.event & EVENT_LIMIT_ENERGY) ?
                                    cfg.limit_energy :
                                    0.;
                                if (fabs(k1 - kinetic_min) <
                                    (double)FLT_EPSILON)
                                        s.st(HF_ENERGY, kinetic_min);
                        } else if ((cfg.event & EVENT_LIMIT_ENERGY) &&
                            (fabs(k1 - cfg.limit_energy) <
                                (double)FLT_EPSILON)) {
                                s.st(HF_ENERGY, cfg.limit_energy);
                        }
                        phase = PHASE_DONE;
                } else if (event & EVENT_VERTEX_DEL) {
                        phase = PHASE_DEL;
                } else {
                        phase = PHASE_EHS;
                }
        }
        flags = (flags & ~(HFL_EVENT_MASK | HFL_PHASE_MASK)) |
            ((uint32_t)event & HFL_EVENT_MASK) |
            ((uint32_t)phase << HFL_PHASE_SHIFT);
        s.stu(HF_FLAGS,
            (unsigned long long)flags | ((unsigned long long)n_steps << 32));
        return phase;
}

/* The general restatement of a hard elastic event on the unpacked track. */
template <class Cfg, class S, class Rng>
PB_NOINLINE int hot_ehs_general(const Physics & p, const Cfg & cfg,
    const UniformGeometry & geo, const HotCtx & h, S & s, Rng & rng)
{
        State st;
        Track t;
        Loop l;
        hot_unpack(s, p, cfg, geo, st, t, l);
        t.step_X_limit = h.step_X_limit;
        const int phase = stepping_ehs(p, cfg, geo, t, rng, st, l);
        hot_pack(s, st, t, l, phase, false, 0.);
        return phase;
}

/*
 * A hard elastic event of a track in PHASE_EHS: transport_do_ehs
 * (pumas.c:5768-5885), then the stop check and the re-arming of the stepping
 * loop (pumas.c:5398-5402, 5476-5488 and transport_limit, pumas.c:5569-5631).
 * Only the energy and the direction of the track take part in the event; the
 * rest of the slot is touched by the re-arming alone.
 */
#if defined(PB_HOT_OUTLINE_CALLS)
#define PB_HOT_EHS_CALL PB_NOINLINE
#else
#define PB_HOT_EHS_CALL PB_FN
#endif
template <class Cfg, class S, class Rng>
PB_HOT_EHS_CALL int hot_ehs(const Physics & p, const Cfg & cfg,
    const UniformGeometry & geo, const HotCtx & h, S & s, Rng & rng)
{
        constexpr bool forward = (Cfg::direction == MODE_FORWARD);
        const double kinetic = s.ld(HF_ENERGY);
        if (!hot_in_grid(h, kinetic))
                return hot_ehs_general(p, cfg, geo, h, s, rng);

        const unsigned long long fw = s.ldu(HF_FLAGS);
        uint32_t flags = (uint32_t)fw;
        State st;
        st.energy = kinetic;
        st.direction[0] = s.ld(HF_UX);
        st.direction[1] = s.ld(HF_UY);
        st.direction[2] = s.ld(HF_UZ);
        transport_do_ehs(p, cfg, rng, st, h.material,
            (int)((flags >> HFL_IK_SHIFT) & 0xffu));
        s.st(HF_UX, st.direction[0]);
        s.st(HF_UY, st.direction[1]);
        s.st(HF_UZ, st.direction[2]);

        int event = (int)(flags & HFL_EVENT_MASK);
        int phase = PHASE_STEP;
        if (cfg.event & event) {
                phase = PHASE_DONE; /* (stepping_finish below) */
        } else {
                /* re-arm: pumas.c:5476-5488 */
                const double weight = s.ld(HF_WEIGHT);
                s.st(HF_TI, s.ld(HF_TIME));
                s.st(HF_WI, weight);
                if (!forward) {
                        /* (Xf and dei only enter the backward weights) */
                        HLoc l;
                        l.i = (int)((flags >> HFL_IK_SHIFT) & 0xffu);
                        if (flags & HFL_LOCATED)
                                l.t = s.ld(HF_T0);
                        else
                                l = hot_locate_K(h, kinetic, l.i);
                        const HotPtr r = hot_K(h, l);
                        s.st(HF_XF, hot_pchip(r + HK_X, l.t));
                        s.st(HF_DEI, 1. / hot_pchip(r + HK_DE, l.t));
                }
                /* transport_limit in the straggled mode */
                event = EVENT_NONE;
                if (weight <= 0.) {
                        event = EVENT_WEIGHT;
                } else if (forward) {
                        const double kinetic_min =
                            (cfg.event & EVENT_LIMIT_ENERGY) ?
                            cfg.limit_energy :
                            0.;
                        if (kinetic <= kinetic_min) event = EVENT_LIMIT_ENERGY;
                } else if ((cfg.event & EVENT_LIMIT_ENERGY) &&
                    (kinetic >= cfg.limit_energy)) {
                        event = EVENT_LIMIT_ENERGY;
                }
                if (event != EVENT_NONE) phase = PHASE_DONE;
        }
        if (phase == PHASE_DONE) {
                /* stepping_finish, pumas.c:5492-5513 */
                if (forward) {
                        const double kinetic_min =
                            (cfg.event & EVENT_LIMIT_ENERGY) ?
                            cfg.limit_energy :
                            0.;
                        if (fabs(kinetic - kinetic_min) < (double)FLT_EPSILON)
                                s.st(HF_ENERGY, kinetic_min);
                } else if ((cfg.event & EVENT_LIMIT_ENERGY) &&
                    (fabs(kinetic - cfg.limit_energy) < (double)FLT_EPSILON)) {
                        s.st(HF_ENERGY, cfg.limit_energy);
                }
        }
        flags = (flags & ~(HFL_EVENT_MASK | HFL_PHASE_MASK)) |
            ((uint32_t)event & HFL_EVENT_MASK) |
            ((uint32_t)phase << HFL_PHASE_SHIFT);
        s.stu(HF_FLAGS,
            (unsigned long long)flags | (fw & 0xffffffff00000000ull));
        return phase;
}

} /* namespace pb */

#endif
