/*
 * straight_b200.c -- the user program of the reference's
 * examples/pumas/straight.c (backward muon flux through a uniform rock layer)
 * written against libpumas_b200.so: same physics, same context settings, one
 * batched call instead of the per-event loop.
 *
 *   cc examples/straight_b200.c -Iinclude -Lpumas_b200 -lpumas_b200 -lm \
 *       -Wl,-rpath,$PWD/pumas_b200 -o straight_b200
 *   ./straight_b200 DUMP DEPTH[m] KINETIC_MIN[GeV] [KINETIC_MAX[GeV]] [N]
 *
 * DUMP is a pumas_physics_dump of the reference (gunzip
 * tests/golden/materials.muon.pumas.gz: the dump of examples/data/materials.xml
 * for muons). The primary spectrum is a plain
 * power law, E^-2.7, in place of the sea level parametrisation of the
 * reference example: the program demonstrates the API, not a flux model.
 *
 * Mapping to the reference program (examples/pumas/straight.c):
 *   :98-104   pumas_physics_create       -> pumas_physics_load (same tables)
 *   :61-74    medium1 callback           -> PUMAS_B200_GEOMETRY_UNIFORM
 *   :110-116  context->mode, limits      -> unchanged
 *   :84-92    context->random            -> Philox seed + track ids
 *   :134-179  loop over pumas_context_transport -> pumas_context_transport_n
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pumas_b200.h"

#define MUON_MASS 0.10566

/* SplitMix64: host side sampling of the final states (the transport itself
 * draws from the library's per-track Philox streams). */
static double uniform01(unsigned long long * s)
{
        unsigned long long z = (*s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        z ^= z >> 31;
        return (double)(z >> 11) * (1. / 9007199254740992.);
}

static double primary_spectrum(double kinetic)
{
        return pow(kinetic + MUON_MASS, -2.7);
}

int main(int argc, char * argv[])
{
        if (argc < 4) {
                fprintf(stderr, "usage: %s DUMP DEPTH KINETIC_MIN "
                                "[KINETIC_MAX] [N]\n", argv[0]);
                return EXIT_FAILURE;
        }
        const double depth = strtod(argv[2], NULL);
        const double kinetic_min = strtod(argv[3], NULL);
        const double kinetic_max =
            (argc > 4) ? strtod(argv[4], NULL) : kinetic_min;
        const long n = (argc > 5) ? strtol(argv[5], NULL, 10) : 1000000;
        const int spectrum = kinetic_max > kinetic_min;

        /* Errors go to the default handler: message on stderr and exit, as in
         * the reference (pumas.c:871-880). */
        struct pumas_physics * physics = NULL;
        FILE * stream = fopen(argv[1], "rb");
        if (stream == NULL) {
                perror(argv[1]);
                return EXIT_FAILURE;
        }
        pumas_physics_load(&physics, stream);
        fclose(stream);

        int material;
        pumas_physics_material_index(physics, "StandardRock", &material);

        struct pumas_context * context = NULL;
        pumas_context_create(&context, physics, 0);
        context->mode.energy_loss = PUMAS_MODE_STRAGGLED;
        context->mode.direction = PUMAS_MODE_BACKWARD;
        context->mode.scattering = PUMAS_MODE_DISABLED;
        context->limit.distance = depth;
        context->event |= PUMAS_EVENT_LIMIT_DISTANCE;

        struct pumas_b200_geometry geometry = { 0 };
        geometry.kind = PUMAS_B200_GEOMETRY_UNIFORM;
        geometry.n_layers = 1;
        geometry.step_type = PUMAS_STEP_CHECK;
        geometry.layer[0].material = material;
        pumas_b200_context_geometry_set(context, &geometry);

        unsigned long seed = 0;
        pumas_context_random_seed_set(context, &seed);

        /* The final states, all at once: vertical, at the bottom of the layer;
         * kinetic energy fixed or log-uniform with the matching weight; charge
         * drawn with probability 1/2 and weight 2. */
        struct pumas_state * states = calloc(n, sizeof(*states));
        int * events = calloc(n, sizeof(*events));
        if ((states == NULL) || (events == NULL)) return EXIT_FAILURE;
        unsigned long long rng = 0x1234ABCDULL;
        const double rk = spectrum ? log(kinetic_max / kinetic_min) : 0.;
        long i;
        for (i = 0; i < n; i++) {
                double kf = kinetic_min, wf = 1.;
                if (spectrum) {
                        kf = kinetic_min * exp(rk * uniform01(&rng));
                        wf = kf * rk;
                }
                states[i].charge = (uniform01(&rng) > 0.5) ? 1. : -1.;
                states[i].energy = kf;
                states[i].weight = 2. * wf;
                states[i].direction[2] = -1.;
        }

        pumas_context_transport_n(context, n, states, events, NULL);

        double w = 0., w2 = 0.;
        for (i = 0; i < n; i++) {
                if (!(events[i] & PUMAS_EVENT_LIMIT_DISTANCE)) continue;
                const double wi =
                    states[i].weight * primary_spectrum(states[i].energy);
                w += wi;
                w2 += wi * wi;
        }
        w /= n;
        const double sigma =
            (n > 1) ? sqrt(((w2 / n) - w * w) / (n - 1)) : 0.;
        printf("Flux : %.5lE \\pm %.5lE %s\n", w, sigma,
            spectrum ? "(integrated, arbitrary normalisation)" :
                       "GeV^{-1} (arbitrary normalisation)");

        free(states);
        free(events);
        pumas_context_destroy(&context);
        pumas_physics_destroy(&physics);
        return EXIT_SUCCESS;
}
