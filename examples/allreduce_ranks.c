/*
 * allreduce_ranks.c -- the multi-GPU path of libpumas_b200.so from C, without
 * Python: one host thread per GPU ("rank"), every rank transports its own range
 * of track ids of the same workload (examples/pumas/straight.c geometry, forward),
 * tallies its batch on the device, accumulates over two batches and the ranks
 * sum their tallies with ONE pumas_b200_comm_allreduce (NCCL) at the end.
 * Check: the reduced number of tracks is ranks x batches x n, and every rank
 * holds the same reduced vector.
 *
 *   gcc -O2 -I include examples/allreduce_ranks.c -o allreduce_ranks \
 *       -L pumas_b200 -lpumas_b200 -lcudart -lpthread
 *   ./allreduce_ranks <physics dump> <ranks>
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "pumas_b200.h"
#include "pumas_b200_geometry.h"

#define N_TRACKS 20000
#define N_BATCHES 2
#define N_BINS 16
#define N_TALLY (4 + N_BINS)

struct rank_job {
        int rank, world;
        struct pumas_physics * physics;
        char id[PUMAS_B200_UNIQUE_ID_BYTES];
        double result[N_TALLY];
        int rc;
};

static void * rank_main(void * arg)
{
        struct rank_job * job = arg;
        job->rc = 1;
        if (cudaSetDevice(job->rank) != cudaSuccess) return NULL;

        struct pumas_b200_comm * comm = NULL;
        if (pumas_b200_comm_create(&comm, job->rank, job->world, job->id,
                sizeof job->id) != PUMAS_RETURN_SUCCESS)
                return NULL;

        struct pumas_context * context = NULL;
        if (pumas_context_create(&context, job->physics, 0) !=
            PUMAS_RETURN_SUCCESS)
                return NULL;
        context->mode.direction = PUMAS_MODE_FORWARD;
        context->event = PUMAS_EVENT_LIMIT_DISTANCE;
        context->limit.distance = 100.;
        unsigned long seed = 7;
        pumas_context_random_seed_set(context, &seed);
        int material = 0;
        pumas_physics_material_index(job->physics, "StandardRock", &material);
        struct pumas_b200_geometry geometry;
        memset(&geometry, 0x0, sizeof geometry);
        geometry.kind = PUMAS_B200_GEOMETRY_UNIFORM;
        geometry.n_layers = 1;
        geometry.layer[0].material = material;
        pumas_b200_context_geometry_set(context, &geometry);

        struct pumas_state * host = calloc(N_TRACKS, sizeof *host);
        struct pumas_state * states = NULL;
        int * events = NULL;
        double *tally = NULL, *total = NULL;
        void * scratch = NULL;
        cudaMalloc((void **)&states, N_TRACKS * sizeof *states);
        cudaMalloc((void **)&events, N_TRACKS * sizeof *events);
        cudaMalloc((void **)&tally, N_TALLY * sizeof *tally);
        cudaMalloc((void **)&total, N_TALLY * sizeof *total);
        cudaMalloc(&scratch, pumas_b200_tally_scratch_bytes(N_BINS));
        cudaMemset(total, 0x0, N_TALLY * sizeof *total);

        for (int batch = 0; batch < N_BATCHES; batch++) {
                for (int i = 0; i < N_TRACKS; i++) {
                        memset(&host[i], 0x0, sizeof host[i]);
                        host[i].charge = -1.;
                        host[i].energy = 10.;
                        host[i].weight = 1.;
                        host[i].direction[2] = 1.;
                }
                cudaMemcpy(states, host, N_TRACKS * sizeof *states,
                    cudaMemcpyHostToDevice);
                /* disjoint track ids per rank and batch */
                pumas_b200_context_track_offset_set(context,
                    (unsigned long long)(batch * job->world + job->rank) *
                        N_TRACKS);
                if (pumas_b200_context_transport_async(context, N_TRACKS,
                        states, events, NULL) != PUMAS_RETURN_SUCCESS)
                        return NULL;
                pumas_b200_tally(states, events, N_TRACKS, 0, 1E-03, 1E+03,
                    N_BINS, tally, scratch, NULL);
                pumas_b200_tally_accumulate(total, tally, N_TALLY, NULL);
        }
        /* the one collective of the run */
        if (pumas_b200_comm_allreduce(comm, total, N_TALLY, NULL) !=
            PUMAS_RETURN_SUCCESS)
                return NULL;
        if (pumas_b200_context_synchronize(context) != PUMAS_RETURN_SUCCESS)
                return NULL;
        cudaDeviceSynchronize();
        cudaMemcpy(job->result, total, N_TALLY * sizeof *total,
            cudaMemcpyDeviceToHost);

        cudaFree(states);
        cudaFree(events);
        cudaFree(tally);
        cudaFree(total);
        cudaFree(scratch);
        free(host);
        pumas_context_destroy(&context);
        pumas_b200_comm_destroy(&comm);
        job->rc = 0;
        return NULL;
}

int main(int argc, char * argv[])
{
        if (argc < 3) {
                fprintf(stderr, "usage: %s <physics dump> <ranks>\n", argv[0]);
                return 2;
        }
        const int world = atoi(argv[2]);
        int devices = 0;
        cudaGetDeviceCount(&devices);
        if ((world < 1) || (world > devices) || (world > 8)) {
                fprintf(stderr, "%d ranks asked for, %d devices\n", world,
                    devices);
                return 2;
        }
        if (!pumas_b200_comm_available()) {
                fprintf(stderr, "NCCL could not be loaded\n");
                return 2;
        }
        struct pumas_physics * physics = NULL;
        FILE * stream = fopen(argv[1], "rb");
        if ((stream == NULL) ||
            (pumas_physics_load(&physics, stream) != PUMAS_RETURN_SUCCESS)) {
                fprintf(stderr, "cannot load %s\n", argv[1]);
                return 2;
        }
        fclose(stream);

        struct rank_job jobs[8];
        pthread_t threads[8];
        char id[PUMAS_B200_UNIQUE_ID_BYTES];
        if (pumas_b200_comm_unique_id(id, sizeof id) != PUMAS_RETURN_SUCCESS)
                return 1;
        for (int r = 0; r < world; r++) {
                jobs[r].rank = r;
                jobs[r].world = world;
                jobs[r].physics = physics;
                memcpy(jobs[r].id, id, sizeof id);
                pthread_create(&threads[r], NULL, &rank_main, &jobs[r]);
        }
        int failed = 0;
        for (int r = 0; r < world; r++) {
                pthread_join(threads[r], NULL);
                if (jobs[r].rc != 0) failed = 1;
        }
        pumas_physics_destroy(&physics);
        if (failed) {
                fprintf(stderr, "a rank failed\n");
                return 1;
        }
        const double expected = (double)world * N_BATCHES * N_TRACKS;
        for (int r = 0; r < world; r++) {
                if (jobs[r].result[2] != expected) {
                        fprintf(stderr, "rank %d: %g tracks, expected %g\n",
                            r, jobs[r].result[2], expected);
                        return 1;
                }
                if (memcmp(jobs[r].result, jobs[0].result,
                        sizeof jobs[0].result) != 0) {
                        fprintf(stderr, "rank %d differs from rank 0\n", r);
                        return 1;
                }
        }
        printf("ranks %d: tracks %.0f, selected %.0f, sum w %.6f, sum w2 %.6f\n",
            world, jobs[0].result[2], jobs[0].result[3], jobs[0].result[0],
            jobs[0].result[1]);
        return 0;
}
