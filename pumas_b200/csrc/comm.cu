/*
 * comm.cu -- the multi-GPU side of the library: tallies accumulated on the
 * device across batches and ONE all-reduce over NCCL at the end of a run
 * (BASELINE.json north_star; SURVEY.md sections 5 and 8e).
 *
 * Tracks shard trivially (every rank transports its own range of track ids,
 * pumas_b200_context_track_offset_set): the only exchange of a run is the sum
 * of the tally vectors, a few hundred bytes to a few hundred kilobytes. NCCL is
 * loaded at run time (dlopen of libnccl.so.2: the copy the process already
 * holds, e.g. PyTorch's, when there is one), so a single-GPU user needs no NCCL
 * at all. A communicator is either created here from a unique id the caller
 * distributes (pumas_b200_comm_unique_id on one rank, then
 * pumas_b200_comm_create on all), or wrapped around the caller's own
 * ncclComm_t (pumas_b200_comm_wrap).
 *
 * Reproducibility: the per-rank tallies are reduced in a fixed order
 * (aux_kernels.cu), accumulated across batches in launch order, and summed over
 * the ranks by NCCL, whose reduction order is a function of the communicator
 * (number of ranks, topology, algorithm) and not of timing.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "pumas_b200.h"

namespace {

/* The few NCCL entry points used, by their public C signatures (nccl.h):
 * ncclResult_t is an int (0 = ncclSuccess), ncclComm_t an opaque pointer,
 * ncclUniqueId 128 bytes, ncclDouble = 8, ncclSum = 0. */
struct NcclUniqueId {
        char internal[128];
};
typedef int (*nccl_get_unique_id_t)(NcclUniqueId *);
typedef int (*nccl_comm_init_rank_t)(void **, int, NcclUniqueId, int);
typedef int (*nccl_comm_destroy_t)(void *);
typedef int (*nccl_all_reduce_t)(
    const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char * (*nccl_get_error_string_t)(int);

struct Nccl {
        void * handle;
        nccl_get_unique_id_t get_unique_id;
        nccl_comm_init_rank_t comm_init_rank;
        nccl_comm_destroy_t comm_destroy;
        nccl_all_reduce_t all_reduce;
        nccl_get_error_string_t get_error_string;
};

const Nccl * nccl()
{
        static Nccl n = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
        static std::once_flag once;
        std::call_once(once, []() {
                const char * names[] = { getenv("PUMAS_B200_NCCL"),
                        "libnccl.so.2", "libnccl.so" };
                for (int i = 0; (i < 3) && (n.handle == nullptr); i++)
                        if (names[i] != nullptr)
                                n.handle = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
                if (n.handle == nullptr) return;
                n.get_unique_id = reinterpret_cast<nccl_get_unique_id_t>(
                    dlsym(n.handle, "ncclGetUniqueId"));
                n.comm_init_rank = reinterpret_cast<nccl_comm_init_rank_t>(
                    dlsym(n.handle, "ncclCommInitRank"));
                n.comm_destroy = reinterpret_cast<nccl_comm_destroy_t>(
                    dlsym(n.handle, "ncclCommDestroy"));
                n.all_reduce = reinterpret_cast<nccl_all_reduce_t>(
                    dlsym(n.handle, "ncclAllReduce"));
                n.get_error_string = reinterpret_cast<nccl_get_error_string_t>(
                    dlsym(n.handle, "ncclGetErrorString"));
        });
        if ((n.handle == nullptr) || (n.get_unique_id == nullptr) ||
            (n.comm_init_rank == nullptr) || (n.comm_destroy == nullptr) ||
            (n.all_reduce == nullptr))
                return nullptr;
        return &n;
}

__global__ void accumulate_kernel(
    double * __restrict__ total, const double * __restrict__ batch, long count)
{
        const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < count) total[i] += batch[i];
}

/* Sums of the weights and squared weights of consecutive segments of
 * `per_bin` tracks (bin b = tracks [b per_bin, (b + 1) per_bin)): one warp
 * per bin, lanes stride the segment, then a shuffle tree: a fixed order. */
__global__ void __launch_bounds__(256)
    tally_segments_kernel(const struct pumas_state * __restrict__ states,
        const int * __restrict__ events, long n_bins, long per_bin,
        int event_mask, double * __restrict__ out)
{
        const long bin = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int lane = threadIdx.x & 31;
        if (bin >= n_bins) return;
        double w = 0., w2 = 0., k = 0.;
        for (long j = lane; j < per_bin; j += 32) {
                const long i = bin * per_bin + j;
                const int e = (events != nullptr) ? events[i] : 0;
                if ((event_mask != 0) && !(e & event_mask)) continue;
                const double wi = states[i].weight;
                w += wi;
                w2 += wi * wi;
                k += 1.;
        }
        for (int d = 16; d > 0; d >>= 1) {
                w += __shfl_down_sync(0xffffffffu, w, d);
                w2 += __shfl_down_sync(0xffffffffu, w2, d);
                k += __shfl_down_sync(0xffffffffu, k, d);
        }
        if (lane == 0) {
                out[3 * bin] = w;
                out[3 * bin + 1] = w2;
                out[3 * bin + 2] = k;
        }
}

} /* namespace */

struct pumas_b200_comm {
        void * nccl_comm;
        int rank, world;
        int owned; /* created here: destroyed here */
};

extern "C" {

enum pumas_return pumas_b200_tally_accumulate(
    double * total, const double * batch, long count, void * cuda_stream)
{
        if ((total == nullptr) || (batch == nullptr) || (count < 0))
                return PUMAS_RETURN_VALUE_ERROR;
        if (count == 0) return PUMAS_RETURN_SUCCESS;
        accumulate_kernel<<<(unsigned int)((count + 255) / 256), 256, 0,
            static_cast<cudaStream_t>(cuda_stream)>>>(total, batch, count);
        return (cudaGetLastError() == cudaSuccess) ?
            PUMAS_RETURN_SUCCESS :
            PUMAS_RETURN_INTERNAL_ERROR;
}

enum pumas_return pumas_b200_tally_segments(const struct pumas_state * states,
    const int * events, long n_bins, long per_bin, int event_mask, double * out,
    void * cuda_stream)
{
        if ((states == nullptr) || (out == nullptr) || (n_bins < 0) ||
            (per_bin <= 0))
                return PUMAS_RETURN_VALUE_ERROR;
        if (n_bins == 0) return PUMAS_RETURN_SUCCESS;
        const long threads = n_bins * 32;
        tally_segments_kernel<<<(unsigned int)((threads + 255) / 256), 256, 0,
            static_cast<cudaStream_t>(cuda_stream)>>>(
            states, events, n_bins, per_bin, event_mask, out);
        return (cudaGetLastError() == cudaSuccess) ?
            PUMAS_RETURN_SUCCESS :
            PUMAS_RETURN_INTERNAL_ERROR;
}

int pumas_b200_comm_available(void) { return (nccl() != nullptr) ? 1 : 0; }

enum pumas_return pumas_b200_comm_unique_id(void * id, int bytes)
{
        const Nccl * n = nccl();
        if ((n == nullptr) || (id == nullptr) ||
            (bytes < (int)sizeof(NcclUniqueId)))
                return PUMAS_RETURN_VALUE_ERROR;
        NcclUniqueId u;
        if (n->get_unique_id(&u) != 0) return PUMAS_RETURN_INTERNAL_ERROR;
        memcpy(id, &u, sizeof u);
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_b200_comm_create(struct pumas_b200_comm ** comm,
    int rank, int world, const void * id, int bytes)
{
        const Nccl * n = nccl();
        if (comm == nullptr) return PUMAS_RETURN_VALUE_ERROR;
        *comm = nullptr;
        if ((n == nullptr) || (id == nullptr) ||
            (bytes < (int)sizeof(NcclUniqueId)) || (world < 1) || (rank < 0) ||
            (rank >= world))
                return PUMAS_RETURN_VALUE_ERROR;
        NcclUniqueId u;
        memcpy(&u, id, sizeof u);
        void * c = nullptr;
        if (n->comm_init_rank(&c, world, u, rank) != 0)
                return PUMAS_RETURN_INTERNAL_ERROR;
        pumas_b200_comm * p =
            static_cast<pumas_b200_comm *>(malloc(sizeof(pumas_b200_comm)));
        if (p == nullptr) {
                n->comm_destroy(c);
                return PUMAS_RETURN_MEMORY_ERROR;
        }
        p->nccl_comm = c;
        p->rank = rank;
        p->world = world;
        p->owned = 1;
        *comm = p;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_b200_comm_wrap(
    struct pumas_b200_comm ** comm, void * nccl_comm, int rank, int world)
{
        if ((comm == nullptr) || (nccl_comm == nullptr) || (nccl() == nullptr))
                return PUMAS_RETURN_VALUE_ERROR;
        pumas_b200_comm * p =
            static_cast<pumas_b200_comm *>(malloc(sizeof(pumas_b200_comm)));
        if (p == nullptr) return PUMAS_RETURN_MEMORY_ERROR;
        p->nccl_comm = nccl_comm;
        p->rank = rank;
        p->world = world;
        p->owned = 0;
        *comm = p;
        return PUMAS_RETURN_SUCCESS;
}

enum pumas_return pumas_b200_comm_allreduce(struct pumas_b200_comm * comm,
    double * values, long count, void * cuda_stream)
{
        const Nccl * n = nccl();
        if ((comm == nullptr) || (n == nullptr) || (values == nullptr) ||
            (count < 0))
                return PUMAS_RETURN_VALUE_ERROR;
        if ((count == 0) || (comm->world == 1)) return PUMAS_RETURN_SUCCESS;
        const int rc = n->all_reduce(values, values, (size_t)count,
            8 /* ncclDouble */, 0 /* ncclSum */, comm->nccl_comm,
            static_cast<cudaStream_t>(cuda_stream));
        return (rc == 0) ? PUMAS_RETURN_SUCCESS : PUMAS_RETURN_INTERNAL_ERROR;
}

void pumas_b200_comm_destroy(struct pumas_b200_comm ** comm)
{
        if ((comm == nullptr) || (*comm == nullptr)) return;
        const Nccl * n = nccl();
        if ((*comm)->owned && (n != nullptr)) n->comm_destroy((*comm)->nccl_comm);
        free(*comm);
        *comm = nullptr;
}

} /* extern "C" */
