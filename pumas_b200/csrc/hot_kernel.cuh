/*
 * hot_kernel.cuh -- the transport kernel of the straggled energy loss mode in a
 * single uniform medium (BASELINE.json headline, configs[0], configs[2]).
 *
 * A persistent kernel, one CTA per SM, around the lean step of pb_hot.h. Most
 * CTAs are *stepping CTAs*, a few are *service CTAs*.
 *
 *  - Tracks in flight live in the SHARED MEMORY of a stepping CTA as slots of
 *    HOT_WORDS 64 bit words, word major per warp (word k of slot e at
 *    slots[k * HOT_STRIDE + e]). A lane owns HOT_SLOTS slots (lane, lane + 32,
 *    ...): its accesses are conflict free whichever of them it works on, and
 *    switching a lane to another track costs nothing: no exchange of 31 words
 *    through a pool as in the register resident kernel (resident.cuh), and
 *    ~100 registers per thread instead of 168.
 *  - Every round a lane runs one Monte Carlo step (hot_step) on one of its
 *    slots that is ready to step. A slot whose step ended in a hard elastic
 *    event (1 step in ~6) waits while the lane goes on with its other slot;
 *    every few rounds -- the same rounds for all warps -- a warp serves the
 *    waiting events of its lanes together (hot_ehs): the wavefront /
 *    event-queue regrouping of BASELINE.json's north_star, the queue being the
 *    slots themselves.
 *  - Hard inelastic events (1 step in ~340, thousands of divergent
 *    instructions: DCS models, rejection loops), finished tracks and empty
 *    slots are handed to the service CTAs: the lane copies the slot to its
 *    *mirror* in HBM (a fixed address per slot: no allocation, no return
 *    queue, nothing to dead lock), pushes its id on a global ring and goes on
 *    with its other slot, polling the mirror's phase word once per round. A
 *    service warp pops up to 32 ids, works on the mirrors, and publishes the
 *    new phase. The stepping CTAs never run that code: the stepping loop is
 *    small enough for the SM's instruction cache to hold (with the cold code
 *    on the same SM, `no_instruction` was 3/4 of all stall cycles:
 *    profiles/r2_hot_v1_ncu.txt).
 *  - The stepping warps of a CTA run in lockstep (one named barrier per round)
 *    so that they share the instruction lines they fetch.
 *  - Random numbers: the Philox4x32-10 stream of a track (philox.h) is
 *    addressed by the number of variates drawn so far, kept in the slot. Each
 *    thread owns a 16 word ring in shared memory topped up at the start of a
 *    step by all lanes together.
 *
 * Results per track do not depend on the schedule (per track streams) and are
 * those of the general kernels bit for bit (tests/test_gpu_schedulers.py).
 */
#ifndef PB_HOT_KERNEL_CUH
#define PB_HOT_KERNEL_CUH

/* 1: the packed tables of the medium's material are staged to shared memory
 * (50 KB) by every stepping CTA; 0: read from HBM through L1. Measured on B200
 * (profiles/r2_experiments.md): staging removes the table loads from the long
 * scoreboard but costs two warps per CTA and most of the L1 that serves the
 * register saves around calls: 1.96e6 against 2.15e6 headline tracks/s. */
#ifndef PB_HOT_SMEM_TABLES_ON
#define PB_HOT_SMEM_TABLES_ON 0
#endif
#if PB_HOT_SMEM_TABLES_ON
#define PB_HOT_SMEM_TABLES
#endif

#include "device_common.cuh"
#include "pb_hot.h"

namespace pb {

#ifndef PB_HOT_WARPS
#if PB_HOT_SMEM_TABLES_ON
#define PB_HOT_WARPS 14
#else
#define PB_HOT_WARPS 16
#endif
#endif
#ifndef PB_HOT_SLOTS
#define PB_HOT_SLOTS 2
#endif
/* 1: scheduler statistics (PUMAS_B200_STATS); costs registers, off in the
 * product build */
#ifndef PB_HOT_STATS
#define PB_HOT_STATS 0
#endif
/* N > 0: one barrier every N rounds for the stepping warps of a CTA */
#ifndef PB_HOT_LOCKSTEP
#define PB_HOT_LOCKSTEP 1
#endif
constexpr int HOT_WARPS = PB_HOT_WARPS;
constexpr int HOT_THREADS = 32 * HOT_WARPS;
constexpr int HOT_SLOTS = PB_HOT_SLOTS;        /* slots per lane */
constexpr int HOT_STRIDE = 32 * HOT_SLOTS;     /* slots per warp */
constexpr int HOT_CTA_SLOTS = HOT_STRIDE * HOT_WARPS;
constexpr int HOT_RING_WORDS = 16;             /* random words per thread */
constexpr unsigned int HOT_FULL = 0xffffffffu;
constexpr unsigned int HOT_NONE = 0xffffffffu;
/* The warp pools are served every that many rounds. */
constexpr int HOT_EHS_PERIOD = 4;

/* Phases of a slot: PHASE_STEP, PHASE_EHS, PHASE_DEL, PHASE_DONE
 * (pb_transport.h), then */
constexpr int HPHASE_EMPTY = 4;   /* wants a fresh track */
constexpr int HPHASE_RETIRED = 5; /* the batch is exhausted */

/* Global workspace (device_common.cuh: WORKSPACE_BYTES): header, the ring of
 * slot ids waiting for a service warp, then the mirrors. A slot is in the ring
 * at most once, and the ring has at least twice as many entries as there are
 * slots: a wrapped producer cannot reach an entry its consumer has not read. */
struct HotHeader {
        unsigned long long cursor; /* (first: BatchArgs::cursor) */
        unsigned int q_head, q_tail;
        int retired;  /* slots retired, all CTAs */
        int n_slots;  /* slots of the launch */
};
constexpr size_t HOT_Q_OFFSET = 256;
constexpr int HOT_QUEUE = 1 << 19;
constexpr size_t HOT_MIRROR_OFFSET =
    HOT_Q_OFFSET + (size_t)HOT_QUEUE * sizeof(unsigned int);
static_assert(2 * 160 * HOT_CTA_SLOTS <= HOT_QUEUE, "service queue too small");
static_assert(HOT_MIRROR_OFFSET +
            (size_t)160 * HOT_CTA_SLOTS * HOT_WORDS * sizeof(double) <=
        WORKSPACE_BYTES,
    "workspace too small");

/* The dynamic shared memory of a stepping CTA, as 64 bit words: control block
 * (2 words), random rings, slots, then the staged tables. */
extern __shared__ __align__(16) unsigned long long hot_smem64[];
constexpr unsigned int HOT_SLOTS0 = 2u +
    (unsigned int)(HOT_RING_WORDS * HOT_THREADS * sizeof(uint32_t) / 8);
constexpr unsigned int HOT_TABLES0 =
    HOT_SLOTS0 + (unsigned int)(HOT_RESIDENT_WORDS * HOT_STRIDE * HOT_WARPS);
static_assert((HOT_TABLES0 & 1u) == 0, "16 byte aligned tables");

/* A slot in shared memory: word k at hot_smem64[off + k * HOT_STRIDE] (a 32
 * bit index: one register, and LDS / STS with immediate offsets). */
struct SmemSlot {
        unsigned int off; /* index of word 0 */
        /* Words >= HOT_RESIDENT_WORDS (weight, ti, wi, Xf, dei: read and
         * written by events only) stay in the slot's mirror in HBM / L2:
         * 15 words per slot instead of 20 leave 92 KB of L1 to the packed
         * tables and the register saves instead of 60 (+5 % tracks/s,
         * profiles/r2_experiments.md section 11). `k` is a constant at every
         * call site: the test folds away. */
        unsigned long long * cold; /* word 0 of the mirror of this slot */
        __device__ __forceinline__ double ld(int k) const
        {
                return __longlong_as_double((long long)ldu(k));
        }
        __device__ __forceinline__ void st(int k, double v)
        {
                stu(k, (unsigned long long)__double_as_longlong(v));
        }
        __device__ __forceinline__ unsigned long long ldu(int k) const
        {
                if (k >= HOT_RESIDENT_WORDS)
                        return __ldcg(&cold[(size_t)k * HOT_STRIDE]);
                return hot_smem64[off + k * HOT_STRIDE];
        }
        __device__ __forceinline__ void stu(int k, unsigned long long v)
        {
                if (k >= HOT_RESIDENT_WORDS)
                        __stcg(&cold[(size_t)k * HOT_STRIDE], v);
                else
                        hot_smem64[off + k * HOT_STRIDE] = v;
        }
};

__device__ __forceinline__ int hot_phase_of(unsigned long long flags_word)
{
        return (int)(((uint32_t)flags_word & HFL_PHASE_MASK) >> HFL_PHASE_SHIFT);
}

__device__ __forceinline__ unsigned long long hot_with_phase(
    unsigned long long flags_word, int phase)
{
        return (flags_word & ~(unsigned long long)HFL_PHASE_MASK) |
            ((unsigned long long)phase << HFL_PHASE_SHIFT);
}

/* Philox stream of one track through a per thread ring in shared memory:
 * `count` words have been consumed, words [count, filled) are in the ring. */
struct HotRng {
        uint32_t key0, key1, trk0, trk1;
        uint32_t count, filled;
        uint32_t ring; /* shared memory address of word 0 of this thread */

        static __device__ __forceinline__ uint32_t lds(uint32_t address)
        {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(address));
                return v;
        }
        static __device__ __forceinline__ void sts(uint32_t address, uint32_t v)
        {
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(address), "r"(v)
                             : "memory");
        }
        __device__ __forceinline__ uint32_t at(uint32_t word) const
        {
                return ring +
                    (word & (HOT_RING_WORDS - 1)) * (uint32_t)(4 * HOT_THREADS);
        }
        __device__ __forceinline__ void attach(uint32_t * base, uint64_t seed)
        {
                ring = (uint32_t)__cvta_generic_to_shared(base);
                key0 = (uint32_t)seed;
                key1 = (uint32_t)(seed >> 32);
                trk0 = trk1 = count = filled = 0;
        }
        __device__ __forceinline__ void put(const uint4 & o)
        {
                sts(at(filled + 0), o.x);
                sts(at(filled + 1), o.y);
                sts(at(filled + 2), o.z);
                sts(at(filled + 3), o.w);
                filled += 4;
        }
        /* Position the stream of `track` after `drawn` variates (nothing is
         * generated here). */
        __device__ __forceinline__ void seek(uint64_t track, uint32_t drawn)
        {
                trk0 = (uint32_t)track;
                trk1 = (uint32_t)(track >> 32);
                count = drawn;
                filled = drawn & ~3u;
        }
        /* Called by the lanes of a step together: at least 12 words ready
         * (a step draws at most 8 outside its retry loops). */
        __device__ __forceinline__ void top_up()
        {
#pragma unroll 1
                for (int r = 0; r < 3; r++) {
                        const bool need =
                            (int)(filled - count) <= (HOT_RING_WORDS - 5);
                        if (!__any_sync(__activemask(), need)) break;
#if defined(PB_HOT_OUTLINE_CALLS)
                        const uint4 o =
                            philox_block(key0, key1, filled >> 2, trk0, trk1);
#else
                        const uint4 o = philox_block_inline(
                            key0, key1, filled >> 2, trk0, trk1);
#endif
                        if (need) put(o);
                }
        }
        __device__ __forceinline__ uint32_t word_fast()
        {
                const uint32_t w = lds(at(count));
                count++;
                return w;
        }
        static __device__ __forceinline__ double to_uniform(uint32_t w)
        {
                return ((double)w + 0.5) * (1.0 / 4294967296.0);
        }
        /* any number of buffered words, refilled when empty */
        __device__ __forceinline__ double uniform()
        {
                if ((int)(filled - count) <= 0) refill();
                return to_uniform(word_fast());
        }
        /* the lean step (pb_hot.h) */
        __device__ __forceinline__ double uniform_fast()
        {
                return to_uniform(word_fast());
        }
        __device__ __forceinline__ double log_uniform_fast()
        {
                return hot_log_uniform(word_fast());
        }
        __device__ __forceinline__ double uniform_keep()
        {
                if ((int)(filled - count) <= 8) refill();
                return to_uniform(word_fast());
        }
        __device__ __forceinline__ double log_uniform_keep()
        {
                if ((int)(filled - count) <= 8) refill();
                return hot_log_uniform(word_fast());
        }
        __device__ __noinline__ void refill()
        {
                put(philox_block(key0, key1, filled >> 2, trk0, trk1));
        }
};

/* A slot's mirror in HBM: same word major layout as the slots of the CTA. */
struct MirrorSlot {
        unsigned long long * base; /* word 0 */
        __device__ __forceinline__ double ld(int k) const
        {
                return __longlong_as_double(
                    (long long)__ldcg(&base[(size_t)k * HOT_STRIDE]));
        }
        __device__ __forceinline__ void st(int k, double v)
        {
                __stcg(&base[(size_t)k * HOT_STRIDE],
                    (unsigned long long)__double_as_longlong(v));
        }
        __device__ __forceinline__ unsigned long long ldu(int k) const
        {
                return __ldcg(&base[(size_t)k * HOT_STRIDE]);
        }
        __device__ __forceinline__ void stu(int k, unsigned long long v)
        {
                __stcg(&base[(size_t)k * HOT_STRIDE], v);
        }
        __device__ __forceinline__ unsigned long long flags_volatile() const
        {
                return *(volatile unsigned long long *)&base[(size_t)HF_FLAGS *
                    HOT_STRIDE];
        }
        __device__ __forceinline__ void flags_publish(unsigned long long v)
        {
                *(volatile unsigned long long *)&base[(size_t)HF_FLAGS *
                    HOT_STRIDE] = v;
        }
};

/* Mirror of slot `e` of warp `warp` of stepping CTA `cta`. */
__device__ __forceinline__ unsigned long long * hot_mirror(
    unsigned char * workspace, int cta, int warp, int e)
{
        return reinterpret_cast<unsigned long long *>(
                   workspace + HOT_MIRROR_OFFSET) +
            ((size_t)cta * HOT_WARPS + warp) * HOT_WORDS * HOT_STRIDE + e;
}

/* Watchdog: a scheduler that makes no progress for HOT_WATCHDOG_NS (a lost
 * slot: nothing queued, the batch not over) must not hang the device; the
 * kernel traps and the call returns PUMAS_RETURN_INTERNAL_ERROR. The longest
 * legitimate wait is one track's hard event: 2.4 s for the runaway backward
 * tracks of profiles/r2_experiments.md section 7 (4 million variates), hence
 * two minutes. */
constexpr unsigned long long HOT_WATCHDOG_NS = 120000000000ull;
__device__ __forceinline__ unsigned long long hot_now_ns()
{
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
}

/* CTA wide control block of a stepping CTA. */
struct HotShared {
        int retired; /* slots of this CTA retired so far */
        int stop[2];
        int pad_;
};

/* Shared memory of a CTA for a grid of `n_cells` cells. */
inline size_t hot_smem_bytes(int n_cells)
{
        size_t bytes = (size_t)HOT_TABLES0 * 8;
#if PB_HOT_SMEM_TABLES_ON
        bytes += (size_t)n_cells * (HOTK_WORDS + HOTX_WORDS) * 8;
#else
        (void)n_cells;
#endif
        return bytes;
}

__device__ __forceinline__ void hot_queue_push(
    HotHeader * g, volatile unsigned int * queue, unsigned int id)
{
        const unsigned int pos = atomicAdd(&g->q_tail, 1u);
        queue[pos & (HOT_QUEUE - 1)] = id;
}

/* Pop one id for each lane that wants one, as many as there are (whole warp).
 * Many service warps pop concurrently. */
__device__ __forceinline__ unsigned int hot_queue_pop(HotHeader * g,
    volatile unsigned int * queue, int wanted, int lane)
{
        unsigned int h = 0;
        int n = 0;
        if (lane == 0) {
                for (;;) {
                        h = *(volatile unsigned int *)&g->q_head;
                        const unsigned int t =
                            *(volatile unsigned int *)&g->q_tail;
                        n = (int)(t - h);
                        if (n > wanted) n = wanted;
                        if (n <= 0) {
                                n = 0;
                                break;
                        }
                        if (atomicCAS(&g->q_head, h, h + (unsigned int)n) == h)
                                break;
                }
        }
        h = __shfl_sync(HOT_FULL, h, 0);
        n = __shfl_sync(HOT_FULL, n, 0);
        unsigned int id = HOT_NONE;
        if (lane < n) {
                const unsigned int at = (h + (unsigned int)lane) & (HOT_QUEUE - 1);
                do {
                        id = queue[at];
                } while (id == HOT_NONE);
                queue[at] = HOT_NONE;
        }
        return id;
}

/* Write a finished track back to HBM: the fields the stepping changes (charge
 * and the decay flag are the caller's, untouched). */
template <class S>
__device__ __forceinline__ void hot_flush(const S & s, long i, uint32_t drawn,
    const BatchArgs & a)
{
        State * const out = a.states + i;
        out->energy = s.ld(HF_ENERGY);
        out->distance = s.ld(HF_DISTANCE);
        out->grammage = s.ld(HF_GRAMMAGE);
        out->time = s.ld(HF_TIME);
        out->weight = s.ld(HF_WEIGHT);
        out->position[0] = s.ld(HF_PX);
        out->position[1] = s.ld(HF_PY);
        out->position[2] = s.ld(HF_PZ);
        out->direction[0] = s.ld(HF_UX);
        out->direction[1] = s.ld(HF_UY);
        out->direction[2] = s.ld(HF_UZ);
        const unsigned long long fw = s.ldu(HF_FLAGS);
        const uint32_t f = (uint32_t)fw;
        if (a.events != nullptr) a.events[i] = (int)(f & HFL_EVENT_MASK);
        if (a.media != nullptr) {
                a.media[2 * i] = 0;
                a.media[2 * i + 1] = 0;
        }
        const int rc = (int)((f >> HFL_RC_SHIFT) & 0xfu);
        if (rc != RETURN_SUCCESS) atomicCAS(a.status, 0, rc);
        if (a.n_random != nullptr) a.n_random[i] = drawn;
        if (a.n_steps != nullptr) a.n_steps[i] = (unsigned int)(fw >> 32);
}

/* Scheduler statistics (diagnostics, PUMAS_B200_STATS + a PB_HOT_STATS
 * build). */
enum {
        HSTAT_ROUNDS = 0,
        HSTAT_STEP_LANES,
        HSTAT_EHS_PASSES,
        HSTAT_EHS_LANES,
        HSTAT_SVC_PASSES,
        HSTAT_SVC_LANES,
        HSTAT_IDLE_ROUNDS,
        N_HSTATS
};

/* Prepares the workspace for one launch. */
static __global__ void hot_workspace_init(unsigned char * workspace, int n_slots)
{
        HotHeader * const g = reinterpret_cast<HotHeader *>(workspace);
        unsigned int * const q =
            reinterpret_cast<unsigned int *>(workspace + HOT_Q_OFFSET);
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i == 0) {
                g->cursor = 0;
                g->q_head = g->q_tail = 0;
                g->retired = 0;
                g->n_slots = n_slots;
        }
        if (i < HOT_QUEUE) q[i] = HOT_NONE;
}

/* The service CTAs: hard inelastic events (pumas.c:5340-5392), finished
 * tracks, fresh tracks (pumas.c:3114-3235, 5210-5273); on the mirrors. */
template <class Cfg>
__device__ __noinline__ void hot_service_loop(const Physics & p, const Cfg & cfg,
    const UniformGeometry & geo, const HotCtx & h, const BatchArgs & a,
    HotRng & rng, unsigned long long * stat, int lane)
{
        HotHeader * const g = reinterpret_cast<HotHeader *>(a.workspace);
        volatile unsigned int * const queue =
            reinterpret_cast<volatile unsigned int *>(a.workspace + HOT_Q_OFFSET);
        int idle = 0;
        unsigned long long idle_since = 0;
        for (;;) {
                __syncwarp();
                int avail = 0, over = 0;
                if (lane == 0) {
                        avail = (int)(*(volatile unsigned int *)&g->q_tail -
                            *(volatile unsigned int *)&g->q_head);
                        over = (*(volatile int *)&g->retired >=
                            *(volatile int *)&g->n_slots);
                }
                avail = __shfl_sync(HOT_FULL, avail, 0);
                over = __shfl_sync(HOT_FULL, over, 0);
                if (over) break;
                if ((avail <= 0) || ((avail < a.del_threshold) && (idle < a.service_patience))) {
                        __nanosleep(2000);
                        idle++;
                        if ((lane == 0) && ((idle & 1023) == 0)) {
                                const unsigned long long now = hot_now_ns();
                                if (idle_since == 0)
                                        idle_since = now;
                                else if (now - idle_since > HOT_WATCHDOG_NS) {
                                        printf("[hot] service warp idle for "
                                               "120 s with the batch not over: "
                                               "retired %d of %d slots\n",
                                            g->retired, g->n_slots);
                                        __trap();
                                }
                        }
                        continue;
                }
                idle = 0;
                idle_since = 0;
                const unsigned int id =
                    hot_queue_pop(g, queue, (avail < 32) ? avail : 32, lane);
                const bool have = (id != HOT_NONE);
                MirrorSlot s;
                s.base = nullptr;
                int phase = HPHASE_RETIRED;
                if (have) {
                        __threadfence();
                        const int cta = (int)(id / (unsigned int)HOT_CTA_SLOTS);
                        const int e = (int)(id % (unsigned int)HOT_CTA_SLOTS);
                        s.base = hot_mirror(
                            a.workspace, cta, e / HOT_STRIDE, e % HOT_STRIDE);
                        phase = hot_phase_of(s.flags_volatile());
                }
#if PB_HOT_STATS
                stat[HSTAT_SVC_PASSES] += 1;
                stat[HSTAT_SVC_LANES] += __popc(__ballot_sync(HOT_FULL, have));
#else
                (void)stat;
#endif
                /* a hard inelastic event */
                if (have && (phase == PHASE_DEL)) {
                        State st;
                        Track t;
                        Loop l;
                        hot_unpack(s, p, cfg, geo, st, t, l);
                        t.step_X_limit = h.step_X_limit;
                        const unsigned long long idw = s.ldu(HF_ID);
                        const uint32_t index = (uint32_t)idw;
                        rng.seek(
                            a.offset + (uint64_t)index, (uint32_t)(idw >> 32));
                        /* (the first 12 variates of the event generated by
                         * the lanes together: inside the DCS code a refill
                         * runs on 2 lanes) */
                        rng.top_up();
                        phase = stepping_del(p, cfg, geo, t, rng, st, l);
                        /* (the owner only sees the phase published below,
                         * after the fence) */
                        hot_pack(s, st, t, l, HPHASE_EMPTY, false, 0.);
                        s.stu(HF_ID, (unsigned long long)index |
                                ((unsigned long long)rng.count << 32));
                }
                /* a finished track */
                if (have && (phase == PHASE_DONE)) {
                        const unsigned long long idw = s.ldu(HF_ID);
                        hot_flush(
                            s, (long)(uint32_t)idw, (uint32_t)(idw >> 32), a);
                        phase = HPHASE_EMPTY;
                }
                /* fresh tracks from the global cursor */
                for (;;) {
                        const unsigned int m = __ballot_sync(
                            HOT_FULL, have && (phase == HPHASE_EMPTY));
                        if (m == 0) break;
                        const int k = __popc(m);
                        unsigned long long base = 0;
                        if (lane == 0)
                                base = atomicAdd(
                                    a.cursor, (unsigned long long)k);
                        base = __shfl_sync(HOT_FULL, base, 0);
                        if (((m >> lane) & 1u) == 0) continue;
                        const long i =
                            (long)base + __popc(m & ((1u << lane) - 1u));
                        if (i >= a.n) {
                                phase = HPHASE_RETIRED;
                                continue;
                        }
                        State st = a.states[i];
                        Track t;
                        Loop l;
                        Outcome out;
                        rng.seek(a.offset + (uint64_t)i, 0);
                        const int ph =
                            track_begin(p, cfg, geo, rng, st, t, l, out);
                        if (ph == PHASE_DONE) {
                                /* over before its first step: the state is
                                 * unchanged */
                                if (a.events != nullptr)
                                        a.events[i] = out.event;
                                if (a.media != nullptr) {
                                        a.media[2 * i] = out.medium[0];
                                        a.media[2 * i + 1] = out.medium[1];
                                }
                                if (out.rc != RETURN_SUCCESS)
                                        atomicCAS(a.status, 0, out.rc);
                                if (a.n_random != nullptr)
                                        a.n_random[i] = rng.count;
                                if (a.n_steps != nullptr)
                                        a.n_steps[i] = t.n_steps;
                                continue; /* still empty */
                        }
                        phase = ph;
                        hot_pack(s, st, t, l, HPHASE_EMPTY, false, 0.);
                        s.stu(HF_ID, (unsigned long long)(uint32_t)i |
                                ((unsigned long long)rng.count << 32));
                }
                if (have) {
                        /* hand the slot back: fields, then the phase */
                        __threadfence();
                        s.flags_publish(hot_with_phase(s.ldu(HF_FLAGS), phase));
                        if (phase == HPHASE_RETIRED) atomicAdd(&g->retired, 1);
                }
        }
}

template <class Cfg>
__global__ void __launch_bounds__(HOT_THREADS, 1)
    transport_hot_kernel(const __grid_constant__ Physics p,
        const __grid_constant__ Cfg cfg,
        const __grid_constant__ UniformGeometry geo,
        const __grid_constant__ HotCtx h, const __grid_constant__ BatchArgs a)
{
        unsigned char * const smem_raw =
            reinterpret_cast<unsigned char *>(hot_smem64);
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;

        HotShared & sh = *reinterpret_cast<HotShared *>(smem_raw);
        uint32_t * const rings = reinterpret_cast<uint32_t *>(smem_raw + 16);
        constexpr unsigned int SLOTS0 = HOT_SLOTS0;

#if PB_HOT_STATS
        unsigned long long stat[N_HSTATS];
#pragma unroll
        for (int k = 0; k < N_HSTATS; k++) stat[k] = 0;
#define HOT_STAT(k, v) stat[k] += (v);
#else
        unsigned long long * const stat = nullptr;
#define HOT_STAT(k, v)
#endif

        HotRng rng;
        rng.attach(rings + threadIdx.x, a.seed);

        if ((int)blockIdx.x < a.n_del_ctas) {
                hot_service_loop(p, cfg, geo, h, a, rng, stat, lane);
        } else {
                /* ---- stepping CTA -------------------------------------------- */
                const int cta = (int)blockIdx.x - a.n_del_ctas;
                HotHeader * const g = reinterpret_cast<HotHeader *>(a.workspace);
                volatile unsigned int * const queue =
                    reinterpret_cast<volatile unsigned int *>(
                        a.workspace + HOT_Q_OFFSET);
                if (threadIdx.x == 0) {
                        sh.retired = 0;
                        sh.stop[0] = sh.stop[1] = 0;
                }
#if PB_HOT_SMEM_TABLES_ON
                /* stage the packed tables of the medium's material */
                for (int k = threadIdx.x; k < h.n_cells * HOTK_WORDS;
                     k += HOT_THREADS)
                        hot_smem64[h.hotK + k] = (unsigned long long)
                            __double_as_longlong(__ldg(h.gK + k));
                for (int k = threadIdx.x; k < h.n_cells * HOTX_WORDS;
                     k += HOT_THREADS)
                        hot_smem64[h.hotX + k] = (unsigned long long)
                            __double_as_longlong(__ldg(h.gX + k));
#endif
                __syncthreads();
                const unsigned int mine = SLOTS0 +
                    (unsigned int)warp * (HOT_RESIDENT_WORDS * HOT_STRIDE) +
                    (unsigned int)lane;
                unsigned long long * const mirror =
                    hot_mirror(a.workspace, cta, warp, lane);
                const unsigned int my_id = (unsigned int)cta * HOT_CTA_SLOTS +
                    (unsigned int)warp * HOT_STRIDE + (unsigned int)lane;
                int phase[HOT_SLOTS];
                bool away[HOT_SLOTS]; /* handed to the service CTAs */
#pragma unroll
                for (int j = 0; j < HOT_SLOTS; j++) {
                        MirrorSlot m;
                        m.base = mirror + 32 * j;
                        m.stu(HF_FLAGS, hot_with_phase(0ull, HPHASE_EMPTY));
                        phase[j] = HPHASE_EMPTY;
                        away[j] = true;
                }
                __threadfence();
#pragma unroll
                for (int j = 0; j < HOT_SLOTS; j++)
                        hot_queue_push(g, queue, my_id + 32u * j);
                int attached = -1; /* slot the random stream is positioned on */
                uint32_t index = 0; /* its track: position in the batch */
                const unsigned int period =
                    (a.ehs_period > 0) ? (unsigned int)a.ehs_period : 1u;
#define HOT_SET(array, j_, v_)                                                 \
        {                                                                      \
                _Pragma("unroll") for (int q = 0; q < HOT_SLOTS; q++) if (     \
                    q == (j_)) array[q] = (v_);                                \
        }
#define HOT_GET(array, j_, out_)                                               \
        {                                                                      \
                _Pragma("unroll") for (int q = 0; q < HOT_SLOTS; q++) if (     \
                    q == (j_)) out_ = array[q];                                \
        }

                bool starved = true; /* no slot to step last round */
                unsigned int tick = 0; /* round % period */
                for (unsigned int round = 0;; round++) {
                        /* (a) slots coming back from the service CTAs: looked
                         * for every `period` rounds (a global load per slot
                         * that is away), or when the lane has nothing else */
                        const bool look = starved || (tick == 0u);
#pragma unroll
                        for (int j = 0; j < HOT_SLOTS; j++) {
                                if (away[j] && look) {
                                        MirrorSlot m;
                                        m.base = mirror + 32 * j;
                                        const int ph =
                                            hot_phase_of(m.flags_volatile());
                                        if ((ph == PHASE_STEP) ||
                                            (ph == HPHASE_RETIRED)) {
                                                __threadfence();
                                                if (ph == PHASE_STEP) {
                                                        SmemSlot s;
                                                        s.off = mine + 32u * j;
                                                        s.cold = mirror + 32 * j;
#pragma unroll 5
                                                        for (int k = 0;
                                                             k < HOT_RESIDENT_WORDS;
                                                             k++)
                                                                s.stu(k,
                                                                    m.ldu(k));
                                                } else {
                                                        atomicAdd(
                                                            &sh.retired, 1);
                                                }
                                                phase[j] = ph;
                                                away[j] = false;
                                                if (attached == j)
                                                        attached = -1;
                                        }
                                }
                        }

                        /* (b) the end: every slot of the CTA retired */
#if PB_HOT_LOCKSTEP
                        if ((PB_HOT_LOCKSTEP == 1) ||
                            ((round % PB_HOT_LOCKSTEP) == 0u)) {
                                const unsigned int b =
                                    (round / PB_HOT_LOCKSTEP) & 1u;
                                if (threadIdx.x == 0)
                                        *(volatile int *)&sh.stop[b] =
                                            (*(volatile int *)&sh.retired >=
                                                HOT_CTA_SLOTS);
                                named_barrier_sync(1, HOT_THREADS);
                                if (*(volatile int *)&sh.stop[b]) break;
                        }
#else
                        {
                                bool all_retired = true;
#pragma unroll
                                for (int j = 0; j < HOT_SLOTS; j++)
                                        all_retired = all_retired &&
                                            (phase[j] == HPHASE_RETIRED);
                                if (__all_sync(HOT_FULL, all_retired)) break;
                        }
#endif
                        HOT_STAT(HSTAT_ROUNDS, 1)

                        /* (c) every `period` rounds: the pending hard elastic
                         * events of the warp, pumas.c:5393-5402 */
                        tick = (tick + 1u == period) ? 0u : tick + 1u;
                        if (tick == 0u) {
                                int je = -1;
#pragma unroll
                                for (int j = HOT_SLOTS - 1; j >= 0; j--)
                                        if (phase[j] == PHASE_EHS) je = j;
#if PB_HOT_STATS
                                const unsigned int m_ehs =
                                    __ballot_sync(HOT_FULL, je >= 0);
                                if (m_ehs != 0) {
                                        HOT_STAT(HSTAT_EHS_PASSES, 1)
                                        HOT_STAT(HSTAT_EHS_LANES, __popc(m_ehs))
                                }
#endif
                                if (je >= 0) {
                                        SmemSlot s;
                                        s.off = mine + 32u * je;
                                        s.cold = mirror + 32 * je;
                                        if (attached != je) {
                                                const unsigned long long idw =
                                                    s.ldu(HF_ID);
                                                index = (uint32_t)idw;
                                                rng.seek(a.offset +
                                                        (uint64_t)index,
                                                    (uint32_t)(idw >> 32));
                                                attached = je;
                                        }
                                        rng.top_up();
                                        const int ph =
                                            hot_ehs(p, cfg, geo, h, s, rng);
                                        s.stu(HF_ID,
                                            (unsigned long long)index |
                                                ((unsigned long long)rng.count
                                                    << 32));
                                        HOT_SET(phase, je, ph)
                                }
                                __syncwarp();
                        }

                        /* (d) one Monte Carlo step */
                        int js = -1;
#pragma unroll
                        for (int j = HOT_SLOTS - 1; j >= 0; j--)
                                if (phase[j] == PHASE_STEP) js = j;
                        {
                                int pa = -1;
                                HOT_GET(phase, attached, pa)
                                if (pa == PHASE_STEP) js = attached;
                        }
                        HOT_STAT(HSTAT_STEP_LANES,
                            __popc(__ballot_sync(HOT_FULL, js >= 0)))
                        starved = (js < 0);
                        if (js >= 0) {
                                SmemSlot s;
                                s.off = mine + 32u * js;
                                s.cold = mirror + 32 * js;
                                if (attached != js) {
                                        const unsigned long long idw =
                                            s.ldu(HF_ID);
                                        index = (uint32_t)idw;
                                        rng.seek(a.offset + (uint64_t)index,
                                            (uint32_t)(idw >> 32));
                                        attached = js;
                                }
                                rng.top_up();
                                const int ph =
                                    hot_step(p, cfg, geo, h, s, rng);
                                /* (the stream position, for whoever works on
                                 * the slot next) */
                                s.stu(HF_ID, (unsigned long long)index |
                                        ((unsigned long long)rng.count << 32));
                                HOT_SET(phase, js, ph)
                        }
                        __syncwarp();

                        /* (e) hard inelastic events and finished tracks go to
                         * the service CTAs, through the mirror of the slot */
#pragma unroll
                        for (int j = 0; j < HOT_SLOTS; j++) {
                                if (!away[j] &&
                                    ((phase[j] == PHASE_DEL) ||
                                        (phase[j] == PHASE_DONE))) {
                                        SmemSlot s;
                                        s.off = mine + 32u * j;
                                        s.cold = mirror + 32 * j;
                                        MirrorSlot m;
                                        m.base = mirror + 32 * j;
#pragma unroll 5
                                        for (int k = 0; k < HOT_RESIDENT_WORDS;
                                             k++)
                                                m.stu(k, s.ldu(k));
                                        __threadfence();
                                        hot_queue_push(
                                            g, queue, my_id + 32u * j);
                                        away[j] = true;
                                }
                        }
                }
#undef HOT_SET
#undef HOT_GET
        }
#if PB_HOT_STATS
        if (a.stats != nullptr) {
#pragma unroll
                for (int k = 0; k < N_HSTATS; k++) {
                        const unsigned long long v = stat[k];
                        if ((lane == 0) && (v != 0)) atomicAdd(a.stats + k, v);
                }
        }
#endif
#undef HOT_STAT
}

template <class Cfg>
cudaError_t launch_hot(const Physics & p, const Physics & host_view,
    const double * hotK, const double * hotX, const Cfg & cfg,
    const UniformGeometry & geo, uint64_t seed, uint64_t offset, long n,
    State * states, int * events, int * media, int * status,
    unsigned int * n_random, unsigned int * n_steps, void * workspace,
    cudaStream_t stream)
{
        const Tunables & t = tunables();
        const int n_cells = host_view.n_energies - 1;
        const size_t smem = hot_smem_bytes(n_cells);
        if (smem > (size_t)227 * 1024) return cudaErrorInvalidConfiguration;
        cudaError_t e = cudaFuncSetAttribute(transport_hot_kernel<Cfg>,
            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        {
                /* experiment: PUMAS_B200_CARVEOUT = shared memory carveout in
                 * percent of the SM's 228 KB (the rest of 256 KB is L1) */
                static const char * carve = getenv("PUMAS_B200_CARVEOUT");
                if (carve != nullptr)
                        cudaFuncSetAttribute(transport_hot_kernel<Cfg>,
                            cudaFuncAttributePreferredSharedMemoryCarveout,
                            atoi(carve));
        }
        int sms = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms > 160) sms = 160; /* (the mirrors of the workspace) */
        if ((t.grid_cap > 1) && (t.grid_cap < sms)) sms = t.grid_cap;
        /* stepping CTAs for the batch, service CTAs for them. Measured on B200
         * (148 SMs; profiles/r2_experiments.md), tracks/s against the number
         * of service CTAs: headline 14 -> 2.06e6, 16 -> 2.30e6, 18 -> 2.29e6,
         * 21 -> 2.23e6; rock_3km 16 -> 1.36e6, 21 -> 1.41e6; backward
         * (straight.c, three times as many hard inelastic events per step)
         * 16 -> 8.3e6, 21 -> 9.4e6, 28 -> 8.8e6 without scattering and
         * 16 -> 5.6e6, 21 -> 6.7e6, 28 -> 6.8e6 with. Below the optimum the
         * stepping CTAs starve. With the variates of an event generated by
         * the lanes together (hot_service_loop) the service is cheaper:
         * headline 12 -> 2.28e6, 14 -> 2.39e6, 16 -> 2.39e6, 18 -> 2.39e6;
         * rock_3km 16 -> 1.40e6, 21 -> 1.50e6; backward 16 -> 8.3e6, 20 ->
         * 9.9e6, 24 -> 9.6e6 without scattering and 16 -> 5.8e6, 20 -> 6.9e6,
         * 24 -> 6.7e6 with: one SM in 8 forward, one in 7 backward. */
        const long per_cta = (long)HOT_CTA_SLOTS;
        const long want = (n + per_cta - 1) / per_cta;
        int n_service = (t.del_ctas > 0) ?
            t.del_ctas :
            ((Cfg::direction == MODE_FORWARD) ? sms / 8 : sms / 7);
        if (n_service > sms - 1) n_service = sms - 1;
        if (n_service < 1) return cudaErrorInvalidConfiguration;
        int n_step = (want < (long)(sms - n_service)) ? (int)want :
                                                        sms - n_service;
        if (t.del_ctas <= 0) {
                const int need = (n_step + 5) / 6;
                if (need < n_service) n_service = need;
        }
        hot_workspace_init<<<(HOT_QUEUE + 255) / 256, 256, 0, stream>>>(
            static_cast<unsigned char *>(workspace), n_step * HOT_CTA_SLOTS);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (t.stats != nullptr)
                cudaMemsetAsync(t.stats, 0, N_STATS * sizeof *t.stats, stream);
        BatchArgs a;
        a.seed = seed;
        a.offset = offset;
        a.n = n;
        a.states = states;
        a.events = events;
        a.media = media;
        a.status = status;
        a.n_random = n_random;
        a.n_steps = n_steps;
        a.cursor = static_cast<unsigned long long *>(workspace);
        a.workspace = static_cast<unsigned char *>(workspace);
        a.n_del_ctas = n_service;
        a.stats = t.stats;
        a.ehs_period = (t.ehs_period > 0) ? t.ehs_period : HOT_EHS_PERIOD;
        a.del_threshold = (t.del_threshold > 0) ? t.del_threshold : 24;
        a.service_patience = (t.service_patience > 0) ? t.service_patience : 3;
        HotCtx h = hot_context(host_view, p, hotK, hotX, cfg, geo.material,
            geo.density);
#if PB_HOT_SMEM_TABLES_ON
        h.hotK = HOT_TABLES0;
        h.hotX = HOT_TABLES0 + (unsigned int)(n_cells * HOTK_WORDS);
#endif
        last_launch_count() = 2;
        transport_hot_kernel<Cfg><<<n_service + n_step, HOT_THREADS, smem,
            stream>>>(p, cfg, geo, h, a);
        e = cudaGetLastError();
#if defined(PB_HOT_MATH_CHECK)
        {
                /* (debugging build: every division / square root of the step
                 * compared with the IEEE operation, pb_hot_math.h) */
                unsigned long long bad[2] = { 0, 0 };
                cudaStreamSynchronize(stream);
                cudaMemcpyFromSymbol(bad, hot_math_mismatches, sizeof bad);
                fprintf(stderr,
                    "[hot] math check: %llu divisions, %llu square roots "
                    "differ from IEEE (cumulated)\n", bad[0], bad[1]);
        }
#endif
        if ((e == cudaSuccess) && (t.stats != nullptr)) {
                unsigned long long hs[N_STATS];
                cudaMemcpyAsync(hs, t.stats, sizeof hs, cudaMemcpyDeviceToHost,
                    stream);
                cudaStreamSynchronize(stream);
                const double rounds =
                    (double)(hs[HSTAT_ROUNDS] ? hs[HSTAT_ROUNDS] : 1);
                fprintf(stderr,
                    "[hot] grid=%d+%d warps=%d slots/lane=%d period=%d "
                    "rounds=%llu step lanes/round=%.2f ehs: passes=%llu "
                    "lanes/pass=%.2f service: passes=%llu lanes/pass=%.2f\n",
                    n_step, n_service, HOT_WARPS, HOT_SLOTS, a.ehs_period,
                    hs[HSTAT_ROUNDS], (double)hs[HSTAT_STEP_LANES] / rounds,
                    hs[HSTAT_EHS_PASSES],
                    (double)hs[HSTAT_EHS_LANES] /
                        (double)(hs[HSTAT_EHS_PASSES] ? hs[HSTAT_EHS_PASSES] :
                                                        1),
                    hs[HSTAT_SVC_PASSES],
                    (double)hs[HSTAT_SVC_LANES] /
                        (double)(hs[HSTAT_SVC_PASSES] ? hs[HSTAT_SVC_PASSES] :
                                                        1));
        }
        return e;
}

#define PB_DEFINE_HOT_SPECIALISATION(NAME, EL, SC, DI, DE)                     \
        bool NAME(const Physics & p, const Physics & host_view,                \
            const Settings & cfg, const Geometry & geo, uint64_t seed,         \
            uint64_t offset, long n, State * states, int * events,             \
            int * media, int * status, unsigned int * n_random,                \
            unsigned int * n_steps, void * workspace, cudaStream_t stream,     \
            cudaError_t * error)                                               \
        {                                                                      \
                typedef StaticSettings<EL, SC, DI, DE> Cfg;                    \
                UniformGeometry ug;                                            \
                if (!Cfg::matches(cfg) || !uniform_geometry_of(geo, &ug))      \
                        return false;                                          \
                *error = launch_hot(p, host_view, p.hotK, p.hotX, Cfg(cfg),    \
                    ug, seed, offset, n, states, events, media, status,        \
                    n_random, n_steps, workspace, stream);                     \
                return true;                                                   \
        }

} /* namespace pb */

#endif
