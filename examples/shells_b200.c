/*
 * shells_b200.c -- a C user program on user `__device__` media: muons started
 * inside concentric spherical shells (examples/user_media.h / user_media.cu),
 * transported forward until they stop or leave. The reference equivalent sets
 * context->medium to a host function (pumas.h:641, 509-511); here the same
 * geometry code runs on the device behind struct pumas_b200_callbacks.
 *
 *   python -m pumas_b200.build --user examples/user_media.cu \
 *       -o examples/libpumas_b200_user.so
 *   cc examples/shells_b200.c -Iinclude -Iexamples -Lexamples \
 *       -lpumas_b200_user -lm -Wl,-rpath,$PWD/examples -o shells_b200
 *   ./shells_b200 DUMP [N] [KINETIC]
 *
 * DUMP: a pumas_physics_dump of the reference holding StandardRock, Water and
 * Air (gunzip tests/golden/materials.muon.pumas.gz).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "pumas_b200.h"
#include "user_media.h"

/* (exported by examples/user_media.cu, built into libpumas_b200_user.so) */
int user_media_callbacks(const struct user_media * shells,
    const int * materials, struct pumas_b200_callbacks * callbacks);

int main(int argc, char * argv[])
{
        if (argc < 2) {
                fprintf(stderr, "usage: %s DUMP [N] [KINETIC]\n", argv[0]);
                return EXIT_FAILURE;
        }
        const long n = (argc > 2) ? atol(argv[2]) : 10000;
        const double kinetic = (argc > 3) ? atof(argv[3]) : 30.;

        FILE * stream = fopen(argv[1], "rb");
        if (stream == NULL) {
                perror(argv[1]);
                return EXIT_FAILURE;
        }
        struct pumas_physics * physics = NULL;
        if (pumas_physics_load(&physics, stream) != PUMAS_RETURN_SUCCESS)
                return EXIT_FAILURE;
        fclose(stream);

        /* rock core with a density gradient, water shell, air */
        struct user_media shells = { 0 };
        shells.n_shells = 3;
        shells.r[0] = 0.;
        shells.r[1] = 50.;
        shells.r[2] = 120.;
        shells.r[3] = 400.;
        shells.density[0] = 2.65E+03;
        shells.gradient[0] = 1E-03;
        shells.density[1] = 1.00E+03;
        shells.density[2] = 1.205;
        shells.gradient[2] = 2E-04;
        shells.step_epsilon = 1E-07;
        int materials[3];
        pumas_physics_material_index(physics, "StandardRock", &materials[0]);
        pumas_physics_material_index(physics, "Water", &materials[1]);
        pumas_physics_material_index(physics, "Air", &materials[2]);

        struct pumas_b200_callbacks callbacks;
        if (user_media_callbacks(&shells, materials, &callbacks) != 0) {
                fprintf(stderr, "no device functions (is a B200 visible?)\n");
                return EXIT_FAILURE;
        }

        struct pumas_context * context = NULL;
        pumas_context_create(&context, physics, 0);
        context->mode.energy_loss = PUMAS_MODE_STRAGGLED;
        context->mode.scattering = PUMAS_MODE_MIXED;
        context->mode.direction = PUMAS_MODE_FORWARD;
        unsigned long seed = 1;
        pumas_context_random_seed_set(context, &seed);
        pumas_b200_context_callbacks_set(context, &callbacks);

        struct pumas_state * states = calloc((size_t)n, sizeof *states);
        int * events = calloc((size_t)n, sizeof *events);
        int * media = calloc(2 * (size_t)n, sizeof *media);
        long i;
        for (i = 0; i < n; i++) {
                /* a fan of directions in the (x, z) plane from (10, 5, -20) */
                const double phi = 2. * M_PI * ((double)i + 0.5) / (double)n;
                states[i].charge = -1.;
                states[i].energy = kinetic;
                states[i].weight = 1.;
                states[i].position[0] = 10.;
                states[i].position[1] = 5.;
                states[i].position[2] = -20.;
                states[i].direction[0] = cos(phi);
                states[i].direction[2] = sin(phi);
        }
        if (pumas_context_transport_n(context, n, states, events, media) !=
            PUMAS_RETURN_SUCCESS)
                return EXIT_FAILURE;

        long left = 0, stopped[3] = { 0, 0, 0 };
        double range = 0.;
        for (i = 0; i < n; i++) {
                if (media[2 * i + 1] < 0)
                        left++;
                else
                        stopped[media[2 * i + 1]]++;
                range += states[i].distance;
        }
        printf("Tracks  : %ld of %.3g GeV\n", n, kinetic);
        printf("Left    : %ld\n", left);
        printf("Stopped : %ld in rock, %ld in water, %ld in air\n",
            stopped[0], stopped[1], stopped[2]);
        printf("Path    : %.6f m on average\n", range / (double)n);

        free(states);
        free(events);
        free(media);
        pumas_context_destroy(&context);
        pumas_physics_destroy(&physics);
        return EXIT_SUCCESS;
}
