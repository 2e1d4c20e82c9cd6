/*
 * resident.cuh -- the track-resident transport kernel (template) and its
 * launcher.
 *
 * A persistent kernel, one CTA per SM. Every lane of a *stepping warp* keeps one
 * track in registers and runs Monte Carlo steps on it (step_transport,
 * pumas.c:6296) round after round: 32 lanes walk the same code on 32 tracks,
 * nothing is loaded or stored between two plain steps. What would break that
 * lockstep is regrouped, the wavefront / event-queue scheme of BASELINE.json's
 * north_star with the queues in shared memory:
 *   - a hard elastic event (EHS, 1 step in ~6; as long as a step): the lane
 *     exchanges its track for one of the warp's *pool* (32 parked tracks in
 *     shared memory) that is ready to step again. Every few rounds -- the
 *     same rounds for all warps -- the warp serves the pending events of its
 *     pool, one lane each: ~26 lanes busy instead of the ~5 that would serve
 *     their event in place;
 *   - a hard inelastic event (DEL, 1 step in ~340 but thousands of divergent
 *     instructions: DCS models, a Brent maximiser) is handed to a *DEL warp*
 *     through a CTA wide parking area and queue; the lane takes back a track
 *     whose event has been served, or a fresh one from the global cursor;
 *   - a finished track is flushed to HBM and replaced in the same way.
 * Queues are rings of 16 bit entry indices; only RS_ENTRIES indices exist, so
 * a ring of that size cannot overflow. When no parking entry is free the lane
 * keeps its track and waits a round: back pressure, no dead lock (the DEL
 * warps never wait for a stepping warp).
 *
 * The stepping warps run in lockstep (one named barrier per round): the hot
 * code is ~75 KB, more than the instruction cache; warps that walk it
 * together share the fetched lines (without the barrier: 2.3x slower).
 *
 * Random numbers: the Philox4x32-10 stream of a track (philox.h) is addressed
 * by the number of variates drawn so far, which is all a parked track keeps.
 * Each thread owns a 16 word ring in shared memory that is topped up at the
 * start of a step by all lanes together (RingRng::top_up), so that the ~100
 * instruction Philox block is executed convergently instead of by the 2-5
 * lanes that happen to run dry inside a step (profiles/r1_v4: 15 % of the warp
 * instructions).
 */
#ifndef PB_RESIDENT_CUH
#define PB_RESIDENT_CUH

#include "device_common.cuh"
#ifdef PB_RS_DEBUG
#include <cassert>
#define RS_ASSERT(x) assert(x)
#else
#define RS_ASSERT(x)
#endif

namespace pb {

#ifndef PB_RS_WARPS
#define PB_RS_WARPS 12
#endif
#ifndef PB_RS_DEL_WARPS
#define PB_RS_DEL_WARPS 1
#endif
/* 1: hard elastic events through the warp pools; 0: served in place */
#ifndef PB_RS_EHS_POOL
#define PB_RS_EHS_POOL 1
#endif
#ifndef PB_RS_LOCKSTEP
#define PB_RS_LOCKSTEP 1
#endif
#define RS_FENCE() __threadfence_block()
/* 1: the hard inelastic events are served by dedicated CTAs (the first
 * `n_del_ctas` of the grid) instead of by the DEL warp of each CTA. The DCS
 * models, the polar angle samplers and the Brent maximiser are ~500 KB of cold
 * code: run on the same SM they evict the stepping loop from the instruction
 * cache, and no_instruction is the first stall reason of this kernel (timing
 * experiment without hard inelastic events: +24 % steps/s). The stepping
 * warps work as before, through the queues and the parking area of their CTA
 * in shared memory; the CTA's last warp becomes a *courier* that moves parked
 * tracks between them and global queues in the workspace, out of the lockstep
 * rounds (a global queue operation is several microseconds of dependent L2
 * round trips). A served track returns to whichever CTA asks first. Set per
 * translation unit (kernels_spec0.cu, kernels_spec1.cu). */
#ifndef PB_RS_OFFLOAD
#define PB_RS_OFFLOAD 0
#endif
#ifndef PB_RS_TIMING
#define PB_RS_TIMING 0
#endif
#if PB_RS_TIMING
#define RS_TICK(k)                                                             \
        {                                                                      \
                const long long now_ = clock64();                              \
                RS_STAT(k, (unsigned long long)(now_ - tick));                 \
                tick = now_;                                                   \
        }
#else
#define RS_TICK(k)
#endif
#ifndef PB_RS_ENTRIES
#define PB_RS_ENTRIES 128
#endif
constexpr int RS_WARPS = PB_RS_WARPS;
constexpr int RS_DEL_WARPS = PB_RS_DEL_WARPS;
constexpr int RS_STEP_WARPS = RS_WARPS - RS_DEL_WARPS;
constexpr int RS_THREADS = 32 * RS_WARPS;
constexpr int RS_ENTRIES = PB_RS_ENTRIES; /* DEL parking entries, power of two */
#ifndef PB_RS_POOL_ENTRIES
#define PB_RS_POOL_ENTRIES 32
#endif
/* pool entries per warp (<= 32: entry i is served by lane i) */
constexpr int RS_POOL = PB_RS_EHS_POOL ? PB_RS_POOL_ENTRIES : 0;
constexpr int RS_POOL_STRIDE = (RS_POOL > 0) ? RS_POOL : 1;
constexpr unsigned int RS_POOL_MASK =
    (RS_POOL >= 32) ? 0xffffffffu : ((1u << RS_POOL) - 1u);
static_assert(RS_POOL <= 32, "one lane per pool entry");
/* A DEL warp starts a pass once that many events are queued (or it has been
 * idle for a while, or the batch is ending). */
constexpr int RS_DEL_THRESHOLD = 24;
/* The pools are served every that many rounds. */
constexpr int RS_EHS_PERIOD = 5;
#ifndef PB_RS_RING_WORDS
#define PB_RS_RING_WORDS 16
#endif
constexpr int RS_RING_WORDS = PB_RS_RING_WORDS; /* random words per thread */
constexpr unsigned int RS_FULL = 0xffffffffu;
constexpr int RS_NONE = 0xffff;
/* Lane phases: PHASE_STEP, PHASE_EHS, PHASE_DEL, PHASE_DONE (pb_transport.h),
 * then */
constexpr int PHASE_EMPTY = 4;
static_assert((RS_ENTRIES & (RS_ENTRIES - 1)) == 0, "power of two");
static_assert(RS_ENTRIES < RS_NONE, "16 bit entry indices");
static_assert(RS_STEP_WARPS > 0, "no stepping warp");
constexpr bool RS_OFFLOAD = (PB_RS_OFFLOAD != 0);
static_assert(!RS_OFFLOAD || (RS_DEL_WARPS == 1), "one courier warp");
static_assert(!RS_OFFLOAD || (PB_RS_LOCKSTEP != 0), "offload needs lockstep");
/* entries of the global parking area (offload), power of two */
constexpr int RS_GENTRIES = 8192;
/* A lane that parks a track without taking a served one in exchange goes on
 * with a fresh track: one more track in flight, and without a limit the parked
 * population grows until every entry, global and shared, holds a track
 * (observed). No fresh track is taken while fewer than that many global
 * entries are free. */
constexpr int RS_INTAKE_LEVEL = RS_GENTRIES / 2;
/* Free entries of a CTA's parking area the stepping lanes leave to the courier
 * for the tracks that come back (it keeps at most 16 waiting). With them a
 * served track can always enter some CTA, a lane with a pending event can
 * always exchange its track for one that waits, and an exchange frees a
 * global entry for the next pending event: the queues cannot dead lock, even
 * with every global entry taken. */
constexpr int RS_INBOUND_RESERVE = 32;
static_assert(RS_GENTRIES < RS_NONE, "16 bit entry indices");

/* Philox stream of one track through a per thread ring in shared memory.
 * `count` words have been consumed, words [count, filled) are in the ring. */
struct RingRng {
        uint32_t key0, key1, trk0, trk1;
        uint32_t count, filled;
        uint32_t ring; /* shared memory address: word w of this thread at
                        * ring + 4 * (w & 15) * RS_THREADS */

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
                    (word & (RS_RING_WORDS - 1)) * (uint32_t)(4 * RS_THREADS);
        }

        __device__ __forceinline__ void attach(uint32_t * base, uint64_t seed)
        {
                ring = (uint32_t)__cvta_generic_to_shared(base);
                key0 = (uint32_t)seed;
                key1 = (uint32_t)(seed >> 32);
                trk0 = trk1 = count = filled = 0;
        }

        /* The track id of the stream (position in the batch + offset). */
        __device__ __forceinline__ uint64_t track() const
        {
                return (uint64_t)trk0 | ((uint64_t)trk1 << 32);
        }

        __device__ __forceinline__ void put(const uint4 & o)
        {
                sts(at(filled + 0), o.x);
                sts(at(filled + 1), o.y);
                sts(at(filled + 2), o.z);
                sts(at(filled + 3), o.w);
                filled += 4;
        }

        /* Position the stream of `track` after `drawn` variates. Nothing is
         * generated here: the ring is left short of `count` (filled - count
         * in -3..0) and the next top_up or uniform fills it. */
        __device__ __forceinline__ void seek(uint64_t track, uint32_t drawn)
        {
                trk0 = (uint32_t)track;
                trk1 = (uint32_t)(track >> 32);
                count = drawn;
                filled = drawn & ~3u;
        }

        /* Called by all lanes together: up to two more blocks. */
        __device__ __forceinline__ void top_up()
        {
#pragma unroll
                for (int r = 0; r < 2; r++) {
                        const uint4 o =
                            philox_block(key0, key1, filled >> 2, trk0, trk1);
                        if ((int)(filled - count) <= (RS_RING_WORDS - 4)) put(o);
                }
        }

        __device__ __forceinline__ double uniform()
        {
                if ((int)(filled - count) <= 0)
                        put(philox_block(key0, key1, filled >> 2, trk0, trk1));
                const uint32_t w = lds(at(count));
                count++;
                return ((double)w + 0.5) * (1.0 / 4294967296.0);
        }
};

/* A track in flight. */
struct Resident {
        State state;
        Track trk;
        Loop lp;
        int medium0; /* first medium (media[0] of the reference) */
};

/* Parked tracks live in shared memory. Accesses are volatile: other warps
 * (or other lanes, later) read what a lane wrote. */
struct ParkMem {
        static __device__ __forceinline__ unsigned long long ld(
            const unsigned long long * p)
        {
                return *(const volatile unsigned long long *)p;
        }
        static __device__ __forceinline__ void st(
            unsigned long long * p, unsigned long long v)
        {
                *(volatile unsigned long long *)p = v;
        }
};

/* A parked track: 31 words of 64 bits, word major (word k of entry e at
 * words[k * STRIDE + e]). */
constexpr int RS_PARK_WORDS = 31;

#define PARK_ST(k, v) ParkMem::st(&words[(k) * STRIDE + e], (v))
#define PARK_W(k) ParkMem::ld(&words[(k) * STRIDE + e])

template <int STRIDE>
__device__ __forceinline__ void park_store(const Resident & s,
    const RingRng & rng, uint64_t offset, unsigned long long * words, int e)
{
        const State & st = s.state;
        PARK_ST(0, pack_d(st.charge));
        PARK_ST(1, pack_d(st.energy));
        PARK_ST(2, pack_d(st.distance));
        PARK_ST(3, pack_d(st.grammage));
        PARK_ST(4, pack_d(st.time));
        PARK_ST(5, pack_d(st.weight));
        PARK_ST(6, pack_d(st.position[0]));
        PARK_ST(7, pack_d(st.position[1]));
        PARK_ST(8, pack_d(st.position[2]));
        PARK_ST(9, pack_d(st.direction[0]));
        PARK_ST(10, pack_d(st.direction[1]));
        PARK_ST(11, pack_d(st.direction[2]));
        const Track & t = s.trk;
        PARK_ST(12, pack_d(t.lifetime));
        PARK_ST(13, pack_d(t.step_X_limit));
        PARK_ST(14, pack_d(t.step_invlb1));
        PARK_ST(15, pack_d(t.randn_next));
        const Loop & l = s.lp;
        PARK_ST(16, pack_d(l.ti));
        PARK_ST(17, pack_d(l.wi));
        PARK_ST(18, pack_d(l.Xi));
        PARK_ST(19, pack_d(l.Xf));
        PARK_ST(20, pack_d(l.dei));
        PARK_ST(21, pack_d(l.grammage_max));
        PARK_ST(22, pack_d(l.step_max_medium));
        PARK_ST(23, pack_d(l.step_max_locals));
        PARK_ST(24, pack_d(l.locals.density));
        PARK_ST(25, pack_d(l.locals.magnet[0]));
        PARK_ST(26, pack_d(l.locals.magnet[1]));
        PARK_ST(27, pack_d(l.locals.magnet[2]));
        /* flags: events are 14 bit masks, media and materials are small */
        const uint32_t f0 = (uint32_t)t.step_event |
            ((uint32_t)t.step_foreseen << 14) |
            ((uint32_t)(t.step_first & 1) << 28) |
            ((uint32_t)(t.randn_done & 1) << 29) |
            ((uint32_t)(st.decayed & 1) << 30) |
            ((uint32_t)(l.locals.magnetized & 1) << 31);
        const uint32_t f1 = (uint32_t)(l.medium & 0xff) |
            ((uint32_t)(l.material & 0xff) << 8) |
            ((uint32_t)(s.medium0 & 0xff) << 16) |
            ((uint32_t)(l.straight & 1) << 24) |
            ((uint32_t)(l.step_max_type & 1) << 25) |
            ((uint32_t)(l.first & 1) << 26) |
            ((uint32_t)(t.error & 0x1f) << 27);
        PARK_ST(28, pack_2(f0, f1));
        PARK_ST(29, pack_2(t.n_steps, rng.count));
        PARK_ST(30, (unsigned long long)(rng.track() - offset) |
            ((unsigned long long)(t.iK & 0xff) << 56));
}

/* Loads a parked track and positions the random stream on it. */
template <int STRIDE>
__device__ __forceinline__ void park_load(Resident & s, RingRng & rng,
    uint64_t offset, const unsigned long long * words, int e)
{
        State & st = s.state;
        st.charge = unpack_d(PARK_W(0));
        st.energy = unpack_d(PARK_W(1));
        st.distance = unpack_d(PARK_W(2));
        st.grammage = unpack_d(PARK_W(3));
        st.time = unpack_d(PARK_W(4));
        st.weight = unpack_d(PARK_W(5));
        st.position[0] = unpack_d(PARK_W(6));
        st.position[1] = unpack_d(PARK_W(7));
        st.position[2] = unpack_d(PARK_W(8));
        st.direction[0] = unpack_d(PARK_W(9));
        st.direction[1] = unpack_d(PARK_W(10));
        st.direction[2] = unpack_d(PARK_W(11));
        Track & t = s.trk;
        t.lifetime = unpack_d(PARK_W(12));
        t.step_X_limit = unpack_d(PARK_W(13));
        t.step_invlb1 = unpack_d(PARK_W(14));
        t.randn_next = unpack_d(PARK_W(15));
        Loop & l = s.lp;
        l.ti = unpack_d(PARK_W(16));
        l.wi = unpack_d(PARK_W(17));
        l.Xi = unpack_d(PARK_W(18));
        l.Xf = unpack_d(PARK_W(19));
        l.dei = unpack_d(PARK_W(20));
        l.grammage_max = unpack_d(PARK_W(21));
        l.step_max_medium = unpack_d(PARK_W(22));
        l.step_max_locals = unpack_d(PARK_W(23));
        l.locals.density = unpack_d(PARK_W(24));
        l.locals.magnet[0] = unpack_d(PARK_W(25));
        l.locals.magnet[1] = unpack_d(PARK_W(26));
        l.locals.magnet[2] = unpack_d(PARK_W(27));
        const unsigned long long f = PARK_W(28);
        const uint32_t f0 = (uint32_t)f, f1 = (uint32_t)(f >> 32);
        t.step_event = (int)(f0 & 0x3fff);
        t.step_foreseen = (int)((f0 >> 14) & 0x3fff);
        t.step_first = (int)((f0 >> 28) & 1);
        t.randn_done = (int)((f0 >> 29) & 1);
        st.decayed = (int)((f0 >> 30) & 1);
        l.locals.magnetized = (int)((f0 >> 31) & 1);
        l.medium = (int)(f1 & 0xff);
        l.material = (int)((f1 >> 8) & 0xff);
        s.medium0 = (int)((f1 >> 16) & 0xff);
        l.straight = (int)((f1 >> 24) & 1);
        l.step_max_type = (int)((f1 >> 25) & 1);
        l.first = (int)((f1 >> 26) & 1);
        t.error = (int)((f1 >> 27) & 0x1f);
        const unsigned long long c = PARK_W(29);
        t.n_steps = (uint32_t)c;
        const unsigned long long ix = PARK_W(30);
        t.iK = (int)(ix >> 56);
        rng.seek(offset + (ix & 0x00ffffffffffffffull), (uint32_t)(c >> 32));
}

/* Unpack the words of a parked track into a lane and position the random
 * stream on it. */
__device__ __forceinline__ void park_unpack(Resident & s, RingRng & rng,
    uint64_t offset, const unsigned long long (&in)[RS_PARK_WORDS])
{
        State & st = s.state;
        Track & t = s.trk;
        Loop & l = s.lp;
        st.charge = unpack_d(in[0]);
        st.energy = unpack_d(in[1]);
        st.distance = unpack_d(in[2]);
        st.grammage = unpack_d(in[3]);
        st.time = unpack_d(in[4]);
        st.weight = unpack_d(in[5]);
        st.position[0] = unpack_d(in[6]);
        st.position[1] = unpack_d(in[7]);
        st.position[2] = unpack_d(in[8]);
        st.direction[0] = unpack_d(in[9]);
        st.direction[1] = unpack_d(in[10]);
        st.direction[2] = unpack_d(in[11]);
        t.lifetime = unpack_d(in[12]);
        t.step_X_limit = unpack_d(in[13]);
        t.step_invlb1 = unpack_d(in[14]);
        t.randn_next = unpack_d(in[15]);
        l.ti = unpack_d(in[16]);
        l.wi = unpack_d(in[17]);
        l.Xi = unpack_d(in[18]);
        l.Xf = unpack_d(in[19]);
        l.dei = unpack_d(in[20]);
        l.grammage_max = unpack_d(in[21]);
        l.step_max_medium = unpack_d(in[22]);
        l.step_max_locals = unpack_d(in[23]);
        l.locals.density = unpack_d(in[24]);
        l.locals.magnet[0] = unpack_d(in[25]);
        l.locals.magnet[1] = unpack_d(in[26]);
        l.locals.magnet[2] = unpack_d(in[27]);
        const unsigned long long f = in[28], c = in[29], ix = in[30];
        const uint32_t g0 = (uint32_t)f, g1 = (uint32_t)(f >> 32);
        t.step_event = (int)(g0 & 0x3fff);
        t.step_foreseen = (int)((g0 >> 14) & 0x3fff);
        t.step_first = (int)((g0 >> 28) & 1);
        t.randn_done = (int)((g0 >> 29) & 1);
        st.decayed = (int)((g0 >> 30) & 1);
        l.locals.magnetized = (int)((g0 >> 31) & 1);
        l.medium = (int)(g1 & 0xff);
        l.material = (int)((g1 >> 8) & 0xff);
        s.medium0 = (int)((g1 >> 16) & 0xff);
        l.straight = (int)((g1 >> 24) & 1);
        l.step_max_type = (int)((g1 >> 25) & 1);
        l.first = (int)((g1 >> 26) & 1);
        t.error = (int)((g1 >> 27) & 0x1f);
        t.n_steps = (uint32_t)c;
        t.iK = (int)(ix >> 56);
        rng.seek(offset + (ix & 0x00ffffffffffffffull), (uint32_t)(c >> 32));
}

/* Exchange the track of a lane with a parked one in place (the lane has a
 * pending event and takes over a track whose event has been served). */
template <int STRIDE>
__device__ __forceinline__ void park_swap(Resident & s, RingRng & rng,
    uint64_t offset, unsigned long long * words, int e)
{
        /* the incoming track: all loads in flight together */
        unsigned long long in[RS_PARK_WORDS];
#pragma unroll
        for (int k = 0; k < RS_PARK_WORDS; k++) in[k] = PARK_W(k);
        park_store<STRIDE>(s, rng, offset, words, e);
        park_unpack(s, rng, offset, in);
}
/* One copy of the load / store / unpack code for the call sites where lanes
 * do different things (the hot code must stay small, see the header): a lane
 * that `takes` the parked track of entry e loads it, a lane that `leaves` its
 * own track there stores it; both = an exchange in place. */
template <int STRIDE>
__device__ __forceinline__ void park_exchange(Resident & s, RingRng & rng,
    uint64_t offset, unsigned long long * words, int e, bool takes, bool leaves)
{
        /* (a lane that only leaves reads the vacant entry for nothing: an
         * unconditional load keeps the words in registers) */
        unsigned long long in[RS_PARK_WORDS];
#pragma unroll
        for (int k = 0; k < RS_PARK_WORDS; k++) in[k] = PARK_W(k);
        if (leaves) park_store<STRIDE>(s, rng, offset, words, e);
        if (takes) park_unpack(s, rng, offset, in);
}
#undef PARK_W
#undef PARK_ST

/* A ring of entry indices in shared memory, many producers, many consumers.
 * The slot arrays hold 2 x ENTRIES slots for at most ENTRIES indices in
 * flight: a producer that wraps around cannot reach a slot whose consumer has
 * claimed it (head advanced) but not read it yet. */
struct Ring {
        unsigned int head, tail;
};

template <int ENTRIES = RS_ENTRIES>
__device__ __forceinline__ void ring_push(
    volatile unsigned short * slots, Ring * r, int e)
{
        const unsigned int pos = atomicAdd(&r->tail, 1u);
        slots[pos & (2 * ENTRIES - 1)] = (unsigned short)e;
}

__device__ __forceinline__ int ring_available(const Ring * r)
{
        const unsigned int h = *(volatile const unsigned int *)&r->head;
        const unsigned int t = *(volatile const unsigned int *)&r->tail;
        return (int)(t - h);
}

/* Pop one index for each lane that wants one, as many as there are. Called by
 * the 32 lanes of a warp together. Returns the index, or -1. */
template <int ENTRIES = RS_ENTRIES>
__device__ __forceinline__ int ring_pop_warp(volatile unsigned short * slots,
    Ring * r, bool want, int lane, int reserve = 0)
{
        const unsigned int m = __ballot_sync(RS_FULL, want);
        if (m == 0) return -1;
        const int k = __popc(m);
        unsigned int h = 0;
        int n = 0;
        if (lane == 0) {
                for (;;) {
                        h = *(volatile unsigned int *)&r->head;
                        const unsigned int t = *(volatile unsigned int *)&r->tail;
                        n = (int)(t - h) - reserve;
                        if (n > k) n = k;
                        if (n <= 0) {
                                n = 0;
                                break;
                        }
                        if (atomicCAS(&r->head, h, h + (unsigned int)n) == h)
                                break;
                }
        }
        h = __shfl_sync(RS_FULL, h, 0);
        n = __shfl_sync(RS_FULL, n, 0);
        const int rank = __popc(m & ((1u << lane) - 1u));
        int e = -1;
        if (want && (rank < n)) {
                const unsigned int at =
                    (h + (unsigned int)rank) & (2 * ENTRIES - 1);
                do {
                        e = slots[at];
                } while (e == RS_NONE);
                slots[at] = (unsigned short)RS_NONE;
        }
        return e;
}

/* Write a finished track back to HBM. */
__device__ __forceinline__ void resident_flush(const Resident & s,
    const RingRng & rng, const Outcome & out, const BatchArgs & a)
{
        const long i = (long)(rng.track() - a.offset);
        const uint32_t drawn = rng.count;
        a.states[i] = s.state;
        if (a.events != nullptr) a.events[i] = out.event;
        if (a.media != nullptr) {
                a.media[2 * i] = out.medium[0];
                a.media[2 * i + 1] = out.medium[1];
        }
        if (out.rc != RETURN_SUCCESS) atomicCAS(a.status, 0, out.rc);
        if (a.n_random != nullptr) a.n_random[i] = drawn;
        if (a.n_steps != nullptr) a.n_steps[i] = s.trk.n_steps;
}

/* Re-alignment of the stepping warps inside a step (pb_transport.h:
 * step_transport, `sync.point(k)`): one mbarrier per alignment point, one
 * arrival per warp and per round, by its first stepping lane; a warp without
 * stepping lanes arrives ahead for the whole round. */
#ifndef PB_RS_SYNC_POINTS
#define PB_RS_SYNC_POINTS 1
#endif
struct MbarSync {
        uint32_t base;     /* shared memory address of the mbarriers */
        uint32_t parity;   /* round & 1 */
        unsigned int mask; /* the stepping lanes of this warp */
        bool leader;

        static __device__ __forceinline__ void arrive(uint32_t address)
        {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(
                                 address)
                             : "memory");
        }

        __device__ __forceinline__ void point(int k) const
        {
                const uint32_t address = base + 8u * (uint32_t)k;
                __syncwarp(mask);
                if (leader) arrive(address);
                asm volatile("{\n"
                             ".reg .pred P1;\n"
                             "RS_WAIT:\n"
                             "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
                             "@P1 bra RS_DONE;\n"
                             "bra RS_WAIT;\n"
                             "RS_DONE:\n"
                             "}" ::"r"(address),
                             "r"(parity), "r"(0x989680u) /* suspend, not spin */
                             : "memory");
        }
};

/* Which kernels re-align inside the step: those whose step has no early
 * return (a uniform medium without locals) and is long enough for the warps
 * to drift apart (mixed scattering; without it the marks cost 13 %). */
template <class Cfg, class Geo> struct StepSync {
        static constexpr bool enabled = false;
};
template <int EL, int DI, int DE>
struct StepSync<StaticSettings<EL, MODE_MIXED, DI, DE>, UniformGeometry> {
        static constexpr bool enabled = (PB_RS_SYNC_POINTS != 0);
};

/* CTA wide shared state. */
struct ResidentShared {
        unsigned long long mbar[N_SYNC_POINTS];
        Ring free_, del_, ret_;
        int live;      /* tracks taken by this CTA and not flushed yet */
        int exhausted; /* the global cursor ran out */
        int stop[2];
        int over; /* (offload) every track of the batch has been flushed */
        int may_intake; /* (offload) the global parking area has room */
};

constexpr size_t RS_POOL_BYTES =
    (size_t)RS_PARK_WORDS * RS_POOL * sizeof(unsigned long long);

/* Scheduler statistics (diagnostics), summed over warps. */
enum {
        RSTAT_ROUNDS = 0,
        RSTAT_STEP_LANES,
        RSTAT_EHS_PASSES,
        RSTAT_EHS_LANES,
        RSTAT_DEL_PASSES,
        RSTAT_DEL_LANES,
        RSTAT_SERVICES,
        RSTAT_WAITS,
        RSTAT_T_POOL, /* clock cycles per section (PB_RS_TIMING builds) */
        RSTAT_T_SERVICE,
        RSTAT_T_BARRIER,
        RSTAT_T_EHS,
        RSTAT_T_STEP,
        N_RSTATS
};
static_assert(N_RSTATS <= N_STATS, "statistics buffer");

/* Shared memory of one CTA: control block, queues, random number rings, the
 * DEL parking area and, when the configuration has hard elastic events
 * (mixed scattering), the warp pools. What is left of the SM's 256 KB is L1
 * for the physics tables and the register spills. */
constexpr size_t RS_PARK_BYTES =
    (size_t)RS_PARK_WORDS * RS_ENTRIES * sizeof(unsigned long long);

/* Layout of the global workspace in offload mode (device_common.cuh:
 * WORKSPACE_BYTES). */
struct OffloadHeader {
        unsigned long long cursor; /* (first: BatchArgs::cursor) */
        int live;                  /* tracks taken and not flushed, all CTAs */
        int pad_;
        Ring free_, del_, ret_;
};
constexpr size_t RS_OFFLOAD_QUEUES = 256; /* byte offset of the queues */
constexpr size_t RS_OFFLOAD_PARK =
    RS_OFFLOAD_QUEUES + 6 * (size_t)RS_GENTRIES * sizeof(unsigned short);
constexpr size_t RS_GPARK_BYTES =
    (size_t)RS_PARK_WORDS * RS_GENTRIES * sizeof(unsigned long long);
static_assert(sizeof(OffloadHeader) <= RS_OFFLOAD_QUEUES, "header");
static_assert(RS_OFFLOAD_PARK + RS_GPARK_BYTES <= WORKSPACE_BYTES,
    "workspace too small");

/* The queues of the global workspace, as a kernel sees them. */
struct OffloadQueues {
        OffloadHeader * h;
        volatile unsigned short * free_q, * del_q, * ret_q;
        unsigned long long * park;

        __device__ __forceinline__ explicit OffloadQueues(unsigned char * w)
        {
                h = reinterpret_cast<OffloadHeader *>(w);
                free_q = reinterpret_cast<volatile unsigned short *>(
                    w + RS_OFFLOAD_QUEUES);
                del_q = free_q + 2 * RS_GENTRIES;
                ret_q = del_q + 2 * RS_GENTRIES;
                park = reinterpret_cast<unsigned long long *>(
                    w + RS_OFFLOAD_PARK);
        }
        /* Changes whenever anything moves (watchdogs). */
        __device__ __forceinline__ unsigned long long progress() const
        {
                return *(volatile unsigned long long *)&h->cursor * 1000003ull +
                    (unsigned long long)(unsigned int)*(volatile int *)&h
                            ->live *
                        7919ull +
                    *(volatile unsigned int *)&h->del_.tail +
                    *(volatile unsigned int *)&h->ret_.tail * 31ull;
        }
        /* Every track has been taken from the cursor ... */
        __device__ __forceinline__ bool all_taken(const BatchArgs & a) const
        {
                return *(volatile unsigned long long *)&h->cursor >=
                    (unsigned long long)a.n;
        }
        /* ... and flushed. (A taker adds to `live` before it advances the
         * cursor: the cursor is read first.) */
        __device__ __forceinline__ bool all_over(const BatchArgs & a) const
        {
                if (!all_taken(a)) return false;
                __threadfence();
                return *(volatile int *)&h->live == 0;
        }
};

/* Prepares the workspace for one launch. */
static __global__ void resident_offload_init(unsigned char * workspace)
{
        OffloadHeader * const h = reinterpret_cast<OffloadHeader *>(workspace);
        unsigned short * const q =
            reinterpret_cast<unsigned short *>(workspace + RS_OFFLOAD_QUEUES);
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i == 0) {
                h->cursor = 0;
                h->live = 0;
                h->free_.head = 0;
                h->free_.tail = RS_GENTRIES;
                h->del_.head = h->del_.tail = 0;
                h->ret_.head = h->ret_.tail = 0;
        }
        if (i < RS_GENTRIES) {
                q[i] = (unsigned short)i;
                q[RS_GENTRIES + i] = (unsigned short)RS_NONE;
                for (int k = 2; k < 6; k++)
                        q[k * RS_GENTRIES + i] = (unsigned short)RS_NONE;
        }
}

/* A value read by lane 0 for the whole warp: the decisions of the queue loops
 * must be warp uniform (their queue operations are warp collectives), and two
 * lanes that poll the same word at different times may see different values. */
template <class F> __device__ __forceinline__ int warp_uniform(F f, int lane)
{
        __syncwarp();
        int v = 0;
        if (lane == 0) v = f();
        return __shfl_sync(RS_FULL, v, 0);
}

/* The courier of a stepping CTA (offload): moves the tracks parked with a
 * pending hard inelastic event from the CTA's shared memory to the global
 * queue of the DEL CTAs, and served tracks back, whole warp, one track per
 * lane. Returns when every track of the batch has been flushed. */
__device__ __forceinline__ void courier_loop(const BatchArgs & a,
    ResidentShared & sh, volatile unsigned short * free_q,
    volatile unsigned short * del_q, volatile unsigned short * ret_q,
    unsigned long long * park, int lane)
{
        const OffloadQueues g(a.workspace);
        long still = 0; /* watchdog: a stuck queue fails loudly */
        int stalled = 0;
        unsigned long long last_sig = 0;
        for (;;) {
                bool moved = false;
                /* outbound: pending events */
                const int out = warp_uniform(
                    [&]() { return ring_available(&sh.del_); }, lane);
                if (out > 0) {
                        /* a global entry first, without waiting for one (a
                         * courier that waited here while every entry holds a
                         * served track nobody fetches would be a deadlock);
                         * then the event: this warp is the only consumer of
                         * the CTA's queue, what it saw is still there */
                        const int eg = ring_pop_warp<RS_GENTRIES>(
                            g.free_q, &g.h->free_, lane < out, lane);
                        const int es = ring_pop_warp(
                            del_q, &sh.del_, eg >= 0, lane);
                        if (eg >= 0) {
                                RS_ASSERT(es >= 0);
                                __threadfence_block();
#pragma unroll
                                for (int k = 0; k < RS_PARK_WORDS; k++)
                                        ParkMem::st(
                                            &g.park[k * RS_GENTRIES + eg],
                                            ParkMem::ld(
                                                &park[k * RS_ENTRIES + es]));
                                __threadfence();
                                ring_push<RS_GENTRIES>(
                                    g.del_q, &g.h->del_, eg);
                                ring_push(free_q, &sh.free_, es);
                                moved = true;
                        }
                }
                /* inbound: served tracks, as long as the CTA has lanes that
                 * may want them (few waiting in shared memory) and room */
                const int in = warp_uniform(
                    [&]() {
                            const int waiting = ring_available(&sh.ret_);
                            const int room = ring_available(&sh.free_);
                            if ((waiting >= 16) || (room <= 0)) return 0;
                            int n = ring_available(&g.h->ret_);
                            if (n > room) n = room;
                            if (n > 16 - waiting) n = 16 - waiting;
                            return n;
                    },
                    lane);
                {
                        if (in > 0) {
                                int eg = ring_pop_warp<RS_GENTRIES>(
                                    g.ret_q, &g.h->ret_, lane < in, lane);
                                const int es = ring_pop_warp(
                                    free_q, &sh.free_, eg >= 0, lane);
                                if ((eg >= 0) && (es < 0)) {
                                        /* the stepping warps took the room:
                                         * give the track back */
                                        ring_push<RS_GENTRIES>(
                                            g.ret_q, &g.h->ret_, eg);
                                        eg = -1;
                                }
                                if (eg >= 0) {
                                        __threadfence();
#pragma unroll
                                        for (int k = 0; k < RS_PARK_WORDS; k++)
                                                ParkMem::st(
                                                    &park[k * RS_ENTRIES + es],
                                                    ParkMem::ld(&g.park[k *
                                                            RS_GENTRIES +
                                                        eg]));
                                        __threadfence_block();
                                        ring_push(ret_q, &sh.ret_, es);
                                        ring_push<RS_GENTRIES>(
                                            g.free_q, &g.h->free_, eg);
                                        moved = true;
                                }
                        }
                }
                /* the end of the batch, for the stepping warps of the CTA */
                const int over = warp_uniform(
                    [&]() {
                            if (lane == 0)
                                    *(volatile int *)&sh.may_intake =
                                        (ring_available(&g.h->free_) >
                                            RS_INTAKE_LEVEL) ?
                                        1 :
                                        0;
                            return g.all_over(a) ? 1 : 0;
                    },
                    lane);
                if (over) {
                        if (lane == 0) *(volatile int *)&sh.over = 1;
                        break;
                }
#ifndef PB_RS_COURIER_SLEEP
#define PB_RS_COURIER_SLEEP 500
#endif
                if (!__any_sync(RS_FULL, moved))
                        __nanosleep(PB_RS_COURIER_SLEEP);
                still = __any_sync(RS_FULL, moved) ? 0 : still + 1;
                if ((still > 0) && ((still & 0x1fffff) == 0)) {
                        /* watchdog, as in the DEL CTAs */
                        const unsigned long long sig = g.progress();
                        stalled = (sig == last_sig) ? stalled + 1 : 0;
                        last_sig = sig;
                        if ((stalled >= 8) && (lane == 0)) {
                                printf("[pumas_b200: courier of CTA %d] "
                                       "stuck: sfree=%d sdel=%d sret=%d "
                                       "gret=%d live=%d cursor=%llu\n",
                                    (int)blockIdx.x, ring_available(&sh.free_),
                                    ring_available(&sh.del_),
                                    ring_available(&sh.ret_),
                                    ring_available(&g.h->ret_),
                                    *(volatile int *)&g.h->live,
                                    *(volatile unsigned long long *)&g.h
                                         ->cursor);
                                __trap();
                        }
                }
        }
}

/* (A template on the CTA size: translation units may be built with different
 * PB_RS_WARPS, see kernels_spec2.cu, and each needs its own host copy.) */
template <int WARPS>
__host__ __device__ constexpr size_t resident_smem_bytes(bool pooled)
{
        return ((sizeof(ResidentShared) + 15) & ~(size_t)15) +
            6 * RS_ENTRIES * sizeof(unsigned short) +
            (size_t)RS_RING_WORDS * (32 * WARPS) * sizeof(uint32_t) +
            (size_t)WARPS * N_RSTATS * sizeof(unsigned long long) +
            RS_PARK_BYTES +
            (pooled ? RS_POOL_BYTES * (WARPS - RS_DEL_WARPS) : 0);
}

/* Do hard elastic events go through the warp pools? */
template <class Cfg> __host__ __device__ inline bool resident_pooled(const Cfg & cfg)
{
        return (RS_POOL > 0) && (cfg.scattering == MODE_MIXED);
}


template <class Cfg, class Geo>
__global__ void __launch_bounds__(RS_THREADS, 1)
    transport_resident_kernel(const __grid_constant__ Physics p,
        const __grid_constant__ Cfg cfg, const __grid_constant__ Geo geo,
        const __grid_constant__ BatchArgs a)
{
        extern __shared__ __align__(16) unsigned char smem_raw[];
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;

        ResidentShared & sh = *reinterpret_cast<ResidentShared *>(smem_raw);
        unsigned char * at =
            smem_raw + ((sizeof(ResidentShared) + 15) & ~(size_t)15);
        volatile unsigned short * const free_q =
            reinterpret_cast<volatile unsigned short *>(at);
        volatile unsigned short * const del_q = free_q + 2 * RS_ENTRIES;
        volatile unsigned short * const ret_q = del_q + 2 * RS_ENTRIES;
        at += 6 * RS_ENTRIES * sizeof(unsigned short);
        uint32_t * const rings = reinterpret_cast<uint32_t *>(at);
        at += (size_t)RS_RING_WORDS * RS_THREADS * sizeof(uint32_t);
        unsigned long long * const stats =
            reinterpret_cast<unsigned long long *>(at);
        at += (size_t)RS_WARPS * N_RSTATS * sizeof(unsigned long long);
        unsigned long long * const park =
            reinterpret_cast<unsigned long long *>(at);
        at += RS_PARK_BYTES;
        unsigned char * const pools = at; /* (absent when not pooled) */
        const bool pooled = resident_pooled(cfg);

        if (threadIdx.x == 0) {
                sh.free_.head = 0;
                sh.free_.tail = RS_ENTRIES;
                sh.del_.head = sh.del_.tail = 0;
                sh.ret_.head = sh.ret_.tail = 0;
                sh.live = 0;
                sh.exhausted = 0;
                sh.over = 0;
                sh.may_intake = 1;
                sh.stop[0] = sh.stop[1] = 0;
                for (int k = 0; k < N_SYNC_POINTS; k++)
                        asm volatile(
                            "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(
                                (uint32_t)__cvta_generic_to_shared(
                                    &sh.mbar[k])),
                            "r"(RS_STEP_WARPS)
                            : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::
                                 : "memory");
        }
        for (int e = threadIdx.x; e < RS_ENTRIES; e += RS_THREADS) {
                free_q[e] = (unsigned short)e;
                free_q[RS_ENTRIES + e] = (unsigned short)RS_NONE;
                del_q[e] = del_q[RS_ENTRIES + e] = (unsigned short)RS_NONE;
                ret_q[e] = ret_q[RS_ENTRIES + e] = (unsigned short)RS_NONE;
        }
        __syncthreads();

        /* scheduler statistics: per warp counters in shared memory (in
         * registers they would cost every thread 26 of its 128) */
        unsigned long long * const wstat = stats + warp * N_RSTATS;
        if (lane < N_RSTATS) wstat[lane] = 0;
        __syncwarp();
#define RS_STAT(k, v)                                                          \
        {                                                                      \
                const unsigned long long v_ = (v); /* (by all lanes) */        \
                if (lane == 0) wstat[k] += v_;                                 \
        }

#if PB_RS_OFFLOAD
        /* the DEL CTAs serve the global queues with all their warps; a
         * stepping CTA has a courier warp instead of a DEL warp */
        const bool del_cta = ((int)blockIdx.x < a.n_del_ctas);
        const OffloadQueues gq(a.workspace);
        volatile unsigned short * const d_free_q = gq.free_q;
        volatile unsigned short * const d_del_q = gq.del_q;
        volatile unsigned short * const d_ret_q = gq.ret_q;
        Ring * const d_free = &gq.h->free_;
        Ring * const d_del = &gq.h->del_;
        Ring * const d_ret = &gq.h->ret_;
        unsigned long long * const d_park = gq.park;
        int * const live = &gq.h->live;
        constexpr int D_ENTRIES = RS_GENTRIES;
#define RS_DEL_FENCE() __threadfence()
        if (!del_cta && (warp >= RS_STEP_WARPS)) {
                courier_loop(a, sh, free_q, del_q, ret_q, park, lane);
        } else if (del_cta) {
#else
        volatile unsigned short * const d_free_q = free_q;
        volatile unsigned short * const d_del_q = del_q;
        volatile unsigned short * const d_ret_q = ret_q;
        Ring * const d_free = &sh.free_;
        Ring * const d_del = &sh.del_;
        Ring * const d_ret = &sh.ret_;
        unsigned long long * const d_park = park;
        int * const live = &sh.live;
        constexpr int D_ENTRIES = RS_ENTRIES;
#define RS_DEL_FENCE() RS_FENCE()
        if (warp >= RS_STEP_WARPS) {
#endif
                /* ---- DEL warp: hard inelastic events of the whole CTA ----
                 * (its own track and stream objects: the out-of-line DCS and
                 * polar angle code takes them by reference, which pins them
                 * to local memory; the stepping warps' must stay in
                 * registers) */
                Resident s;
                RingRng rng;
                rng.attach(rings + threadIdx.x, a.seed);
                int idle = 0;
#if PB_RS_OFFLOAD
                int stalled = 0;
                unsigned long long last_sig = 0;
#endif
                for (;;) {
#if PB_RS_OFFLOAD
                        const int avail = warp_uniform(
                            [&]() { return ring_available(d_del); }, lane);
                        const int ending = warp_uniform(
                            [&]() {
                                    return gq.all_taken(a) ?
                                        (gq.all_over(a) ? 2 : 1) :
                                        0;
                            },
                            lane);
                        const bool over_soon = (ending >= 1);
                        const bool over = (ending >= 2);
#else
                        const int avail = ring_available(d_del);
                        const bool over_soon =
                            (*(volatile int *)&sh.exhausted != 0);
                        const bool over = over_soon &&
                            (*(volatile int *)&sh.live == 0);
#endif
                        if ((avail >= a.del_threshold) ||
                            ((avail > 0) && ((idle >= 64) || over_soon))) {
                                const int e = ring_pop_warp<D_ENTRIES>(
                                    d_del_q, d_del, lane < avail, lane);
                                int flushed = 0;
                                if (e >= 0) {
                                        RS_DEL_FENCE();
                                        park_load<D_ENTRIES>(
                                            s, rng, a.offset, d_park, e);
                                        RS_ASSERT((s.lp.material >= 0) &&
                                            (s.lp.material < 8));
                                        RS_ASSERT((long)(rng.track() -
                                                      a.offset) < a.n);
                                        RS_ASSERT(s.trk.step_event != 0);
                                        rng.top_up();
                                        const int ph = stepping_del(p, cfg, geo,
                                            s.trk, rng, s.state, s.lp);
                                        if (ph == PHASE_DONE) {
                                                Outcome out;
                                                out.medium[0] =
                                                    (s.medium0 == 0xff) ?
                                                    -1 :
                                                    s.medium0;
                                                track_end(s.trk, s.lp, out);
                                                resident_flush(s, rng, out, a);
                                                ring_push<D_ENTRIES>(
                                                    d_free_q, d_free, e);
                                                flushed = 1;
                                        } else {
                                                park_store<D_ENTRIES>(s, rng,
                                                    a.offset, d_park, e);
                                                RS_DEL_FENCE();
                                                ring_push<D_ENTRIES>(
                                                    d_ret_q, d_ret, e);
                                        }
                                }
                                const unsigned int got =
                                    __ballot_sync(RS_FULL, e >= 0);
                                const unsigned int fl =
                                    __ballot_sync(RS_FULL, flushed != 0);
                                if ((lane == 0) && (fl != 0)) {
                                        __threadfence();
                                        atomicSub(live, __popc(fl));
                                }
                                RS_STAT(RSTAT_DEL_PASSES, 1)
                                RS_STAT(RSTAT_DEL_LANES, __popc(got))
                                idle = 0;
                        } else if (over) {
                                break;
                        } else {
                                __nanosleep(1000);
                                idle++;
#if PB_RS_OFFLOAD
                                /* watchdog: ~25 s without an event to serve
                                 * while nothing moves anywhere in the batch
                                 * is a stuck queue: fail loudly */
                                if ((idle & 0xfffff) == 0) {
                                        const unsigned long long sig =
                                            gq.progress();
                                        stalled =
                                            (sig == last_sig) ? stalled + 1 : 0;
                                        last_sig = sig;
                                        if ((stalled >= 8) && (lane == 0) &&
                                            (warp == 0)) {
                                                printf("[pumas_b200: DEL CTA "
                                                       "%d] stuck: live=%d "
                                                       "cursor=%llu gfree=%d "
                                                       "gdel=%d gret=%d\n",
                                                    (int)blockIdx.x,
                                                    *(volatile int *)live,
                                                    *(volatile unsigned long long
                                                            *)a.cursor,
                                                    ring_available(d_free),
                                                    ring_available(d_del),
                                                    ring_available(d_ret));
                                                __trap();
                                        }
                                }
#endif
                        }
                }
        } else {
                /* ---- stepping warp ---------------------------------------- */
                Resident s;
                RingRng rng;
                rng.attach(rings + threadIdx.x, a.seed);
                int phase = PHASE_EMPTY;
                bool exhausted = false;
                const unsigned int lt_mask = (1u << lane) - 1u;
                /* the pool of this warp: entries ready to step again, and
                 * entries with a pending hard elastic event */
                unsigned long long * const pool =
                    reinterpret_cast<unsigned long long *>(
                        pools + RS_POOL_BYTES * warp);
                unsigned int ready = 0, pending = 0;
                const unsigned int period =
                    (a.ehs_period > 0) ? (unsigned int)a.ehs_period : 1u;

#if PB_RS_TIMING
                long long tick = clock64();
#endif
                for (unsigned int round = 0;; round++) {
                        /* (a) a lane with a hard elastic event to serve
                         * exchanges its track for a ready one of the pool, or
                         * parks it in a free entry; an empty lane takes a
                         * ready one */
                        if (pooled) {
                                const bool ehs = (phase == PHASE_EHS);
                                const bool empty = (phase == PHASE_EMPTY);
                                const unsigned int m_ehs =
                                    __ballot_sync(RS_FULL, ehs);
                                const unsigned int m_empty = (ready != 0) ?
                                    __ballot_sync(RS_FULL, empty) :
                                    0u;
                                if ((m_ehs | m_empty) != 0) {
                                        const int n_ready = __popc(ready);
                                        const unsigned int vacant =
                                            ~(ready | pending) & RS_POOL_MASK;
                                        const int n_ehs = __popc(m_ehs);
                                        /* ranks: events first, then empty */
                                        const int r = ehs ?
                                            __popc(m_ehs & lt_mask) :
                                            n_ehs + __popc(m_empty & lt_mask);
                                        int e = -1;
                                        bool swap = false;
                                        if (ehs || empty) {
                                                if (r < n_ready) {
                                                        e = __fns(ready, 0,
                                                            r + 1);
                                                        swap = true;
                                                } else if (ehs &&
                                                    (r - n_ready <
                                                        __popc(vacant))) {
                                                        e = __fns(vacant, 0,
                                                            r - n_ready + 1);
                                                }
                                        }
                                        unsigned int to_pending = 0;
                                        unsigned int to_vacant = 0;
                                        if (e >= 0) {
                                                /* swap && ehs: exchange;
                                                 * swap: take; else: leave */
                                                park_exchange<RS_POOL_STRIDE>(
                                                    s, rng, a.offset, pool, e,
                                                    swap, ehs);
                                                phase = swap ? PHASE_STEP :
                                                               PHASE_EMPTY;
                                                if (ehs)
                                                        to_pending = 1u << e;
                                                else
                                                        to_vacant = 1u << e;
                                        }
                                        RS_STAT(RSTAT_WAITS,
                                            __popc(__ballot_sync(RS_FULL,
                                                ehs && (e < 0))))
                                        to_pending = __reduce_or_sync(
                                            RS_FULL, to_pending);
                                        to_vacant = __reduce_or_sync(
                                            RS_FULL, to_vacant);
                                        ready &= ~(to_pending | to_vacant);
                                        pending |= to_pending;
                                        __syncwarp();
                                }
                        }

                        RS_TICK(RSTAT_T_POOL)
                        /* (b) finished tracks, hard inelastic events, empty
                         * lanes */
                        if (__any_sync(RS_FULL, (phase != PHASE_STEP) &&
                                (!pooled || (phase != PHASE_EHS)))) {
                                RS_STAT(RSTAT_SERVICES, 1)
                                /* finished tracks go back to HBM */
                                const bool done = (phase == PHASE_DONE);
                                if (done) {
                                        Outcome out;
                                        out.medium[0] = (s.medium0 == 0xff) ?
                                            -1 :
                                            s.medium0;
                                        track_end(s.trk, s.lp, out);
                                        resident_flush(s, rng, out, a);
                                        phase = PHASE_EMPTY;
                                }
                                int gone = __popc(__ballot_sync(RS_FULL, done));

                                /* a pending hard inelastic event: the lane
                                 * hands its track over to the DEL warps, in
                                 * exchange for one whose event has been served
                                 * when there is one */
                                const bool del = (phase == PHASE_DEL);
                                int e = ring_pop_warp(ret_q, &sh.ret_, del, lane);
                                if (e >= 0) {
                                        RS_FENCE();
                                        park_swap<RS_ENTRIES>(
                                            s, rng, a.offset, park, e);
                                        RS_FENCE();
                                        ring_push(del_q, &sh.del_, e);
                                        phase = PHASE_STEP;
                                }
                                const bool del_2 = (phase == PHASE_DEL);
#if PB_RS_OFFLOAD
                                /* (the last free entries of the CTA are kept
                                 * for the tracks that come back: see
                                 * RS_INBOUND_RESERVE) */
                                e = ring_pop_warp(free_q, &sh.free_, del_2,
                                    lane, RS_INBOUND_RESERVE);
#else
                                e = ring_pop_warp(
                                    free_q, &sh.free_, del_2, lane);
#endif
                                if (e >= 0) {
                                        park_store<RS_ENTRIES>(
                                            s, rng, a.offset, park, e);
                                        RS_FENCE();
                                        ring_push(del_q, &sh.del_, e);
                                        phase = PHASE_EMPTY;
                                }
                                RS_STAT(RSTAT_WAITS,
                                    __popc(__ballot_sync(
                                        RS_FULL, del_2 && (e < 0))))

                                /* take back tracks whose event was served */
                                e = ring_pop_warp(ret_q, &sh.ret_,
                                    phase == PHASE_EMPTY, lane);
                                if (e >= 0) {
                                        RS_FENCE();
                                        park_load<RS_ENTRIES>(
                                            s, rng, a.offset, park, e);
                                        RS_FENCE();
                                        ring_push(free_q, &sh.free_, e);
                                        phase = PHASE_STEP;
                                }

                                /* then fresh ones from the global cursor */
#if PB_RS_OFFLOAD
                                /* (no new track while the global parking area
                                 * is half full: the lanes wait for served
                                 * ones; the courier watches the level) */
                                const unsigned int m = __ballot_sync(RS_FULL,
                                    (phase == PHASE_EMPTY) &&
                                        (*(volatile int *)&sh.may_intake !=
                                            0));
#else
                                const unsigned int m = __ballot_sync(
                                    RS_FULL, phase == PHASE_EMPTY);
#endif
                                if ((m != 0) && !exhausted) {
                                        const int k = __popc(m);
                                        long base = 0;
                                        if (lane == 0) {
                                                atomicAdd(live, k);
#if PB_RS_OFFLOAD
                                                /* (other CTAs read the two) */
                                                __threadfence();
#endif
                                                base = (long)atomicAdd(a.cursor,
                                                    (unsigned long long)k);
                                        }
                                        base = __shfl_sync(RS_FULL, base, 0);
                                        long got = a.n - base;
                                        if (got > k) got = k;
                                        if (got < 0) got = 0;
                                        if (base + k >= a.n) {
                                                exhausted = true;
                                                if (lane == 0)
                                                        *(volatile int *)&sh
                                                             .exhausted = 1;
                                        }
                                        gone += k - (int)got;
                                        const int rank = __popc(m & lt_mask);
                                        bool early = false;
                                        if ((phase == PHASE_EMPTY) &&
                                            (((m >> lane) & 1u) != 0) &&
                                            (rank < got)) {
                                                const long i = base + rank;
                                                s.state = a.states[i];
                                                rng.seek(
                                                    a.offset + (uint64_t)i, 0);
                                                Outcome out;
                                                phase = track_begin(p, cfg, geo,
                                                    rng, s.state, s.trk, s.lp,
                                                    out);
                                                s.medium0 = out.medium[0] & 0xff;
                                                if (phase == PHASE_DONE) {
                                                        resident_flush(s, rng, out, a);
                                                        phase = PHASE_EMPTY;
                                                        early = true;
                                                }
                                        }
                                        gone += __popc(
                                            __ballot_sync(RS_FULL, early));
                                }
                                if ((gone > 0) && (lane == 0)) {
                                        __threadfence();
                                        atomicSub(live, gone);
                                }
                        }

                        RS_TICK(RSTAT_T_SERVICE)
                        /* the end: no track left in this CTA, none to take */
#if PB_RS_LOCKSTEP
#if PB_RS_OFFLOAD
                        /* (tracks come back from the DEL CTAs to any stepping
                         * CTA: everybody stays until the last one is flushed,
                         * which the courier watches) */
                        if (threadIdx.x == 0)
                                *(volatile int *)&sh.stop[round & 1] =
                                    *(volatile int *)&sh.over;
#else
                        if (threadIdx.x == 0)
                                *(volatile int *)&sh.stop[round & 1] =
                                    (*(volatile int *)&sh.exhausted != 0) &&
                                    (*(volatile int *)&sh.live == 0);
#endif
                        named_barrier_sync(1, 32 * RS_STEP_WARPS);
                        if (*(volatile int *)&sh.stop[round & 1]) break;
#else
                        if ((*(volatile int *)&sh.exhausted != 0) &&
                            (*(volatile int *)&sh.live == 0))
                                break;
                        if (__all_sync(RS_FULL, phase != PHASE_STEP))
                                __nanosleep(200);
#endif
                        RS_STAT(RSTAT_ROUNDS, 1)
                        RS_TICK(RSTAT_T_BARRIER)

                        /* (c) every `period` rounds, all warps together: the
                         * pending hard elastic events of the pool, entry i by
                         * lane i, whose own track goes to the pool meanwhile;
                         * a lane that could not park its event serves it in
                         * place */
                        if (pooled && ((round % period) == period - 1u)) {
                                const bool mine = ((pending >> lane) & 1u) != 0;
                                const bool in_place = (phase == PHASE_EHS);
                                const bool swap = mine && (phase == PHASE_STEP);
                                const bool load = mine && (phase == PHASE_EMPTY);
                                if (swap || load)
                                        park_exchange<RS_POOL_STRIDE>(s, rng,
                                            a.offset, pool, lane, true, swap);
                                const unsigned int m_swap =
                                    __ballot_sync(RS_FULL, swap);
                                const unsigned int m_load =
                                    __ballot_sync(RS_FULL, load);
                                pending &= ~(m_swap | m_load);
                                ready |= m_swap;
                                loop_canonical(p, cfg, geo, s.trk, s.lp);
                                const bool run = swap || load || in_place;
                                const unsigned int m_run =
                                    __ballot_sync(RS_FULL, run);
                                if (m_run != 0) {
                                        RS_STAT(RSTAT_EHS_PASSES, 1)
                                        RS_STAT(RSTAT_EHS_LANES, __popc(m_run))
                                }
                                if (run) {
                                        rng.top_up();
                                        phase = stepping_ehs(p, cfg, geo, s.trk,
                                            rng, s.state, s.lp);
                                }
                                __syncwarp();
                        }

                        RS_TICK(RSTAT_T_EHS)
                        /* (d) one Monte Carlo step for every lane */
                        loop_canonical(p, cfg, geo, s.trk, s.lp);
                        const unsigned int m_step =
                            __ballot_sync(RS_FULL, phase == PHASE_STEP);
                        RS_STAT(RSTAT_STEP_LANES, __popc(m_step))
                        if (StepSync<Cfg, Geo>::enabled) {
                                MbarSync sync;
                                sync.base = (uint32_t)__cvta_generic_to_shared(
                                    &sh.mbar[0]);
                                sync.parity = round & 1u;
                                sync.mask = m_step;
                                sync.leader = (m_step != 0) &&
                                    (lane == __ffs(m_step) - 1);
                                if (m_step == 0) {
                                        if (lane == 0)
                                                for (int k = 0;
                                                     k < N_SYNC_POINTS; k++)
                                                        MbarSync::arrive(
                                                            sync.base + 8u * k);
                                } else if (phase == PHASE_STEP) {
                                        rng.top_up();
                                        phase = stepping_step(p, cfg, geo,
                                            s.trk, rng, s.state, s.lp, sync);
                                }
                        } else if (phase == PHASE_STEP) {
                                rng.top_up();
                                phase = stepping_step(
                                    p, cfg, geo, s.trk, rng, s.state, s.lp);
                        }
                        if (!pooled && (phase == PHASE_EHS)) {
                                rng.top_up();
                                phase = stepping_ehs(
                                    p, cfg, geo, s.trk, rng, s.state, s.lp);
                        }
#ifdef PB_RS_HACK_NODEL
                        /* timing experiment only (wrong physics): drop the
                         * hard inelastic events */
                        if (phase == PHASE_DEL) phase = PHASE_STEP;
#endif
                        __syncwarp();
                        RS_TICK(RSTAT_T_STEP)
                }
        }
        if ((a.stats != nullptr) && (lane == 0)) {
#pragma unroll
                for (int k = 0; k < N_RSTATS; k++)
                        atomicAdd(a.stats + k, wstat[k]);
        }
}

template <class Cfg, class Geo>
cudaError_t launch_resident(const Physics & p, const Cfg & cfg, const Geo & geo,
    uint64_t seed, uint64_t offset, long n, State * states, int * events,
    int * media, int * status, unsigned int * n_random, unsigned int * n_steps,
    void * workspace, cudaStream_t stream)
{
        /* per instantiation and per device: the opted-in size */
        static size_t ready[MAX_DEVICES] = { 0 };
        const Tunables & t = tunables();
        const size_t smem =
            resident_smem_bytes<RS_WARPS>(resident_pooled(cfg));
        cudaError_t e;
        const int dev = current_device();
        if (ready[dev] < smem) {
                e = cudaFuncSetAttribute(transport_resident_kernel<Cfg, Geo>,
                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                ready[dev] = smem;
        }
        const int sms = device_sms(dev);
        /* Persistent grid: one CTA per SM, fewer for small n. */
        const long per_cta = 32L * RS_STEP_WARPS;
        const long want = (n + per_cta - 1) / per_cta;
        int grid = (want < sms) ? (int)want : sms;
        if ((t.grid_cap > 0) && (t.grid_cap < grid)) grid = t.grid_cap;
        /* Offload: one CTA in 15 serves the hard inelastic events of the
         * others (headline workload on B200: 8 to 16 of 148 give the same
         * throughput, 4 starve the stepping CTAs). */
        int n_del_ctas = 0;
        if (RS_OFFLOAD) {
                n_del_ctas = (t.del_ctas > 0) ? t.del_ctas : (grid + 7) / 15;
                if (n_del_ctas < 1) n_del_ctas = 1;
                if (n_del_ctas > 16) n_del_ctas = 16;
                if (grid + n_del_ctas > sms) grid = sms - n_del_ctas;
                if (grid < 1) return cudaErrorInvalidConfiguration;
        }
        unsigned long long * cursor =
            static_cast<unsigned long long *>(workspace);
        if (RS_OFFLOAD) {
                resident_offload_init<<<(RS_GENTRIES + 255) / 256, 256, 0,
                    stream>>>(static_cast<unsigned char *>(workspace));
                e = cudaGetLastError();
        } else {
                e = cudaMemsetAsync(
                    cursor, 0, sizeof(unsigned long long), stream);
        }
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
        a.cursor = cursor;
        a.workspace = static_cast<unsigned char *>(workspace);
        a.n_del_ctas = n_del_ctas;
        a.stats = t.stats;
        a.ehs_period = (t.ehs_period > 0) ? t.ehs_period : RS_EHS_PERIOD;
        a.del_threshold =
            (t.del_threshold > 0) ? t.del_threshold : RS_DEL_THRESHOLD;
        grid += n_del_ctas;
        last_launch_count() = RS_OFFLOAD ? 2 : 1;
        transport_resident_kernel<Cfg, Geo>
            <<<grid, RS_THREADS, smem, stream>>>(p, cfg, geo, a);
        e = cudaGetLastError();
        if ((e == cudaSuccess) && (t.stats != nullptr)) {
                unsigned long long h[N_STATS];
                cudaMemcpyAsync(h, t.stats, sizeof h, cudaMemcpyDeviceToHost,
                    stream);
                cudaStreamSynchronize(stream);
                const double rounds =
                    (double)(h[RSTAT_ROUNDS] ? h[RSTAT_ROUNDS] : 1);
                fprintf(stderr,
                    "[resident] grid=%d warps=%d+%d pool=%d period=%d rounds=%llu "
                    "step lanes/round=%.2f services/round=%.3f waits=%llu "
                    "ehs: passes=%llu lanes/pass=%.2f del: passes=%llu "
                    "lanes/pass=%.2f\n",
                    grid, RS_STEP_WARPS, RS_DEL_WARPS,
                    resident_pooled(cfg) ? RS_POOL : 0, a.ehs_period,
                    h[RSTAT_ROUNDS], (double)h[RSTAT_STEP_LANES] / rounds,
                    (double)h[RSTAT_SERVICES] / rounds, h[RSTAT_WAITS],
                    h[RSTAT_EHS_PASSES],
                    (double)h[RSTAT_EHS_LANES] /
                        (double)(h[RSTAT_EHS_PASSES] ? h[RSTAT_EHS_PASSES] : 1),
                    h[RSTAT_DEL_PASSES],
                    (double)h[RSTAT_DEL_LANES] /
                        (double)(h[RSTAT_DEL_PASSES] ? h[RSTAT_DEL_PASSES] : 1));
#if PB_RS_TIMING
                fprintf(stderr,
                    "[resident] cycles/round: pool=%.0f service=%.0f "
                    "barrier=%.0f ehs=%.0f step=%.0f\n",
                    (double)h[RSTAT_T_POOL] / rounds,
                    (double)h[RSTAT_T_SERVICE] / rounds,
                    (double)h[RSTAT_T_BARRIER] / rounds,
                    (double)h[RSTAT_T_EHS] / rounds,
                    (double)h[RSTAT_T_STEP] / rounds);
#endif
        }
        return e;
}

/* Specialised instantiations (kernels_spec*.cu), tried in turn by
 * launch_transport before the general kernel. Each returns false when the
 * settings / geometry are not its own. */
#define PB_SPECIALISATION(NAME)                                                \
        bool NAME(const Physics & p, const Settings & cfg,                     \
            const Geometry & geo, uint64_t seed, uint64_t offset, long n,      \
            State * states, int * events, int * media, int * status,           \
            unsigned int * n_random, unsigned int * n_steps, void * workspace, \
            cudaStream_t stream, cudaError_t * error)

#define PB_DEFINE_SPECIALISATION(NAME, EL, SC, DI, DE)                         \
        PB_SPECIALISATION(NAME)                                                \
        {                                                                      \
                typedef StaticSettings<EL, SC, DI, DE> Cfg;                    \
                UniformGeometry ug;                                            \
                if (!Cfg::matches(cfg) || !uniform_geometry_of(geo, &ug))      \
                        return false;                                          \
                *error = launch_resident(p, Cfg(cfg), ug, seed, offset, n,     \
                    states, events, media, status, n_random, n_steps,          \
                    workspace, stream);                                        \
                return true;                                                   \
        }

/* The same for any descriptor geometry (layers, locals, magnetic field): only
 * the settings are compile time constants. */
#define PB_DEFINE_SPECIALISATION_GEO(NAME, EL, SC, DI, DE)                     \
        PB_SPECIALISATION(NAME)                                                \
        {                                                                      \
                typedef StaticSettings<EL, SC, DI, DE> Cfg;                    \
                if (!Cfg::matches(cfg)) return false;                          \
                *error = launch_resident(p, Cfg(cfg), geo, seed, offset, n,    \
                    states, events, media, status, n_random, n_steps,          \
                    workspace, stream);                                        \
                return true;                                                   \
        }

} /* namespace pb */

#endif
