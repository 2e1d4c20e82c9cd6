/*
 * pb_media.h -- built-in media: the device side replacement of the user
 * `pumas_medium_cb` (pumas.h:509) and `pumas_locals_cb` (pumas.h:348).
 *
 * Media are identified by their layer index (-1 = outside, the reference's
 * NULL medium). Semantics follow the two example geometries of the reference:
 * examples/pumas/straight.c:61-74 (uniform) and examples/pumas/geometry.c:64-163
 * (horizontal layers with constant or exponential density and a uniform
 * magnetic field). oracle/oracle_driver.c implements the same descriptor with
 * host callbacks for the CPU oracle.
 */
#ifndef PB_MEDIA_H
#define PB_MEDIA_H

#include "pb_physics.h"
#include "pumas_b200_geometry.h"

namespace pb {

typedef struct pumas_b200_geometry Geometry;

/* struct medium_locals, pumas.c:398-405 */
struct Locals {
        double density;
        double magnet[3];
        int magnetized;
};

/* The medium callback. `want_medium` / `want_step` mirror the NULL-ness of the
 * two out pointers of pumas_medium_cb. Returns the step type. */
PB_FN int medium_locate(const Geometry & g, int direction, const State & state,
    bool want_medium, int * medium, bool want_step, double * step_ptr)
{
        if (g.kind == PUMAS_B200_GEOMETRY_UNIFORM) {
                if (want_medium) *medium = 0;
                if (want_step) *step_ptr = 0.;
                return g.step_type;
        }
        if (!want_medium && !want_step) return g.step_type;

        const double z = state.position[2];
        const double sgn = (direction == MODE_FORWARD) ? 1 : -1;
        const double uz = state.direction[2] * sgn;
        double step = g.step_default;
        const int n = g.n_layers;
        if ((z < g.z[0]) || !(z < g.z[n])) {
                if (want_medium) *medium = -1;
                step = -1.;
        } else {
                int i = 0;
                while ((i < n - 1) && !(z < g.z[i + 1])) i++;
                if (want_medium) *medium = i;
                if (uz > (double)FLT_EPSILON)
                        step = (g.z[i + 1] - z) / uz;
                else if (uz < -(double)FLT_EPSILON)
                        step = (g.z[i] - z) / uz;
        }
        if (want_step) {
                if (step > 0) step += g.step_epsilon;
                *step_ptr = step;
        }
        return g.step_type;
}

/* The locals callback of a layer (examples/pumas/geometry.c:64-97): density,
 * field, and the size of the inhomogeneities (0: uniform). */
PB_FN double layer_locals(const struct pumas_b200_layer & l, const State & state,
    double * density, double * magnet)
{
        double step_max;
        magnet[0] = l.magnet[0];
        magnet[1] = l.magnet[1];
        magnet[2] = l.magnet[2];
        if (l.density_model == PUMAS_B200_DENSITY_EXPONENTIAL) {
                *density = l.density * xexp(-state.position[2] / l.scale);
                const double uz = fabs(state.direction[2]);
                step_max = l.scale / ((uz <= l.eps) ? l.eps : uz);
        } else {
                *density = l.density;
                step_max = 0.;
        }
        return step_max;
}

/* transport_set_locals, pumas.c:5528-5552. Returns the proposed locals step
 * times the accuracy (<= 0: uniform). */
PB_FN double set_locals(const Physics & p, const Geometry & g, double accuracy,
    int medium, const State & state, Locals * locals)
{
        const struct pumas_b200_layer & l = g.layer[medium];
        if (!l.has_locals) {
                locals->density = p.material_density[l.material];
                locals->magnet[0] = locals->magnet[1] = locals->magnet[2] = 0.;
                locals->magnetized = 0;
                return 0;
        }
        const double step_max =
            layer_locals(l, state, &locals->density, locals->magnet);
        if (locals->density <= 0)
                locals->density = p.material_density[l.material];
        const double * b = locals->magnet;
        locals->magnetized =
            ((b[0] != 0.) || (b[1] != 0.) || (b[2] != 0.)) ? 1 : 0;
        return step_max * accuracy;
}

/* ---- compile time specialisation of the geometry ---------------------------
 * The transport functions are templates on the geometry type. `Geometry` (the
 * C descriptor) is the general, run time case. `UniformGeometry` is the same
 * thing as a PUMAS_B200_GEOMETRY_UNIFORM descriptor whose single layer has no
 * locals callback (examples/pumas/straight.c), with everything that cannot
 * happen there (boundaries, density gradients, magnetic field) known at
 * compile time, so that the kernels specialised for it carry none of that
 * code. Results are identical to the general case. */
struct UniformGeometry {
        int material;
        int step_type;
        /* <= 0: the density of the material; else the constant density of a
         * locals callback without field (examples/pumas/straight.c:61-74 sets
         * 2.65E+03 that way): transport_set_locals (pumas.c:5528-5552) then
         * returns that density and no step limit, everything else is the
         * uniform case. */
        double density;
};

/* Is the descriptor a single infinite medium of constant density without
 * field? Fills `ug`. */
inline bool uniform_geometry_of(const Geometry & geo, UniformGeometry * ug)
{
        if (geo.kind != PUMAS_B200_GEOMETRY_UNIFORM) return false;
        const struct pumas_b200_layer & l = geo.layer[0];
        ug->material = l.material;
        ug->step_type = geo.step_type;
        ug->density = 0.;
        if (!l.has_locals) return true;
        if ((l.density_model != PUMAS_B200_DENSITY_CONSTANT) ||
            (l.magnet[0] != 0.) || (l.magnet[1] != 0.) || (l.magnet[2] != 0.))
                return false;
        ug->density = (l.density > 0.) ? l.density : 0.;
        return true;
}

/* The density of the medium. */
PB_FN double uniform_density(const Physics & p, const UniformGeometry & g)
{
        return (g.density > 0.) ? g.density : p.material_density[g.material];
}

PB_FN int geo_material(const Geometry & g, int medium)
{
        return g.layer[medium].material;
}
PB_FN int geo_material(const UniformGeometry & g, int) { return g.material; }

/* Can a step end in another medium? */
PB_FN bool geo_has_boundaries(const Geometry &) { return true; }
PB_FN constexpr bool geo_has_boundaries(const UniformGeometry &) { return false; }

/* Can the locals (density, field) vary or the medium be magnetised? */
PB_FN bool geo_has_locals(const Geometry &) { return true; }
PB_FN constexpr bool geo_has_locals(const UniformGeometry &) { return false; }

PB_FN int medium_locate(const UniformGeometry & g, int, const State &,
    bool want_medium, int * medium, bool want_step, double * step_ptr)
{
        if (want_medium) *medium = 0;
        if (want_step) *step_ptr = 0.;
        return g.step_type;
}

PB_FN double set_locals(const Physics & p, const UniformGeometry & g, double,
    int, const State &, Locals * locals)
{
        locals->density = uniform_density(p, g);
        locals->magnet[0] = locals->magnet[1] = locals->magnet[2] = 0.;
        locals->magnetized = 0;
        return 0;
}

/* ---- callbacks -------------------------------------------------------------
 * The geometry behind two function pointers (device functions on the GPU,
 * host functions in the CPU build): pumas_medium_cb / pumas_locals_cb. */
struct CallbackGeometry {
        struct pumas_b200_callbacks cb;
};

PB_FN int geo_material(const CallbackGeometry & g, int medium)
{
        return g.cb.material[medium];
}
PB_FN bool geo_has_boundaries(const CallbackGeometry &) { return true; }
PB_FN bool geo_has_locals(const CallbackGeometry &) { return true; }

PB_FN int medium_locate(const CallbackGeometry & g, int direction,
    const State & state, bool want_medium, int * medium, bool want_step,
    double * step_ptr)
{
        return g.cb.medium(g.cb.user_data,
            reinterpret_cast<const struct pumas_state *>(&state),
            (direction == MODE_BACKWARD) ? 1 : 0, want_medium ? medium : nullptr,
            want_step ? step_ptr : nullptr);
}

/* transport_set_locals, pumas.c:5528-5552 */
PB_FN double set_locals(const Physics & p, const CallbackGeometry & g,
    double accuracy, int medium, const State & state, Locals * locals)
{
        const int material = g.cb.material[medium];
        if (!g.cb.has_locals[medium] || (g.cb.locals == nullptr)) {
                locals->density = p.material_density[material];
                locals->magnet[0] = locals->magnet[1] = locals->magnet[2] = 0.;
                locals->magnetized = 0;
                return 0;
        }
        const double step_max = g.cb.locals(g.cb.user_data, medium,
            reinterpret_cast<const struct pumas_state *>(&state),
            &locals->density, locals->magnet);
        if (locals->density <= 0)
                locals->density = p.material_density[material];
        const double * b = locals->magnet;
        locals->magnetized =
            ((b[0] != 0.) || (b[1] != 0.) || (b[2] != 0.)) ? 1 : 0;
        return step_max * accuracy;
}

} /* namespace pb */

#endif
