// Forward tracing kernel (sm_100a): persistent, dynamic ray fetch, 4-wide quantised BVH walk, warp-cooperative leaf tests,
// sort and compositing.
//
// Replaces gaussiantrace_forward.cu:12-141 of the reference (raygen with 16-hit chunks + any-hit sorted insertion).
//
// Every lane owns one ray and runs a small state machine
//     FETCH -> TRAV -> COMP -> (TRAV for another pass | FETCH)          TRAV -> FULL -> TRAV
//   TRAV  near-first stack walk of the BVH in two alternating sub-phases.  NODE: two 256-bit loads fetch a 64-byte
//         4-wide quantised node (a binary node folded with its two children: up to four grandchildren, 16 bits per
//         plane, conservative), four slab tests; children that are leaves are not tested on the spot but pushed on a
//         small per-lane queue of pending leaves (shared memory), internal children are walked / stacked.
//         LEAF: the pending leaves of the whole warp are flattened into one list and tested 32 (ray, surfel) pairs at a
//         time, one pair per lane whoever owns the ray (two 256-bit loads fetch the 64-byte record; the owner's ray
//         comes through shuffles).  With node and leaf work interleaved per lane, ncu showed the leaf code (27 % of all
//         issued instructions) running at 3.4 of 32 lanes; with every lane draining its own queue, 5.7.
//         The traversal stack lives in shared memory ([entry][lane]: conflict free for any mix of depths).
//         (A variant that fetched each 64-byte record with four lanes and staged it through a swizzled shared-memory
//         tile measured 25 % slower than direct loads and was removed; profiles/r01_notes.md.)
//         A surfel that passes the plane / alpha test is APPENDED (one 16-byte store: t, surfel id, alpha, leaf
//         position; unsorted) to its ray's candidate row in a global scratch buffer (L2 resident).
//   FULL  the row holds KB candidates: the warp sorts it co-operatively.  If the buffered hits alone already push the
//         transmittance below T_min at some entry, nothing behind that entry can ever be composited: the row is
//         trimmed there and the walk continues with its range clipped to that depth.  Otherwise the pass's depth range
//         is SPLIT: the nearest KB/2 candidates are kept, the range is clipped to the last of them and the rest of the
//         ray is left to a following pass (no per-candidate eviction ever happens).
//   COMP  the pass's walk is finished.  The rows of as many finished rays as fit are packed into the 32 lanes (one
//         candidate per lane, a ray's candidates in consecutive lanes); per ray the candidates are ranked through
//         shuffles and permuted into depth order, the transmittance chain is evaluated in the sequential order of the
//         reference, one hit per lane is shaded (SH colour: six 256-bit loads, all hits in flight at once), and one lane
//         per (ray, output channel) adds the weighted terms in depth order to the (pre-zeroed) outputs.  The ordered
//         surfel ids are written with one coalesced store (hit list for the backward replay).
// A lane that finishes its ray pulls the next one from a global counter (refills wait for FETCH_MIN idle lanes); the warp
// leaves the walk to serve the other phases as soon as fewer than MIN_ACTIVE lanes are still walking.
#include "trace_common.cuh"

namespace irgs {

constexpr int KB = 32;             // candidates per row == warp width (one candidate per lane in the COMP phase)
#ifndef IRGS_SSTK
#define IRGS_SSTK 32
#endif
constexpr int SSTK = IRGS_SSTK;    // traversal stack entries kept in shared memory; deeper entries spill to local
#ifndef IRGS_MIN_ACTIVE
#define IRGS_MIN_ACTIVE 24
#endif
#ifndef IRGS_FWD_BLOCKS
#define IRGS_FWD_BLOCKS 6
#endif
constexpr int MIN_ACTIVE = IRGS_MIN_ACTIVE;
#ifndef IRGS_PQ
#define IRGS_PQ 16
#endif
constexpr int PQ = IRGS_PQ;               // pending-leaf queue entries per lane
// Row stride of the queue: a lane pushes into its own column, but the LEAF sub-phase pops the entries of ONE owner with
// consecutive lanes -- with rows of 32 words they all sit in the owner's bank (an m-way conflict for m pending leaves); rows of
// 33 words skew the columns over the banks
#ifndef IRGS_PS
#define IRGS_PS 33
#endif
constexpr int PS = IRGS_PS;
#ifndef IRGS_FETCH_MIN
#define IRGS_FETCH_MIN 4
#endif
constexpr int FETCH_MIN = IRGS_FETCH_MIN;   // idle lanes a refill waits for (1: refill at once)
#ifndef IRGS_LEAF_2STAGE
#define IRGS_LEAF_2STAGE 0
#endif
constexpr bool LEAF_2STAGE = IRGS_LEAF_2STAGE != 0;
#ifndef IRGS_COMP_MIN
#define IRGS_COMP_MIN 24
#endif
constexpr int COMP_MIN = IRGS_COMP_MIN;   // candidates that must be waiting before a packed compositing round is run
#ifndef IRGS_DRAIN_MAX
#define IRGS_DRAIN_MAX 0   // measured (profiles/r01_sweeps.txt): 8 makes a 2^14-ray call 10 % faster, 2^18 rays unchanged, the C3 step 2.5 % slower
#endif
constexpr int DRAIN_MAX = IRGS_DRAIN_MAX;   // drain: at most this many walking lanes per warp -> co-operative wide walk (0: off)
constexpr int DRAIN_G = 4;                  // lanes (= node visits in flight) per walking ray in the wide walk
static_assert(DRAIN_MAX * DRAIN_G <= 32, "one group of DRAIN_G lanes per walking ray");
static_assert(DRAIN_MAX >= 0, "0 compiles the wide walk out");
constexpr int DRAIN_SP_MAX = SSTK - 3 * DRAIN_G - 1;   // deepest stack the wide walk accepts: sp + 3 G + 1 <= SSTK entries afterwards
#ifndef IRGS_FULL_RANK_SMEM
#define IRGS_FULL_RANK_SMEM 1     // 0: the full-row rank loop through 64 shuffles (comparison builds)
#endif
#ifndef IRGS_FULL_CHAIN_SMEM
#define IRGS_FULL_CHAIN_SMEM 1    // 0: the full-row transmittance chain through one shuffle per candidate (comparison builds)
#endif
#ifndef IRGS_LEAF_PIPELINE
#define IRGS_LEAF_PIPELINE 1      // 0: every leaf round locates its own items first (comparison builds)
#endif
#ifndef IRGS_LEAF_PACK
#define IRGS_LEAF_PACK 1          // 0: seven shuffles of per-owner integers and window ids per leaf round instead of two
#endif
enum { PH_FETCH = 0, PH_TRAV = 1, PH_COMP = 2, PH_FULL = 3 };
constexpr int CUR_NONE = INT_MIN;

// Row stride of the accumulation scratch: the lanes that add up one ray's channels read [channel][hit] with the SAME hit index --
// a stride of 32 words put all channels of a ray into one bank (8-way conflicts on every step of the accumulation loop)
#ifndef IRGS_CS
#define IRGS_CS 33
#endif
constexpr int CS = IRGS_CS;
template <int NF>   // feature channels the variant is built for: 0, 4 (base colour + roughness: relight, primary pass) or 12
struct WarpSmem {
    float scratch[(8 + NF) * CS];   // co-operative sort (4 x 32) / accumulation scratch (one row of CS words per output channel)
    int pend[PQ * PS];        // pending leaves [entry][lane], rows of PS words
    int stack[SSTK * 32];     // traversal stack [entry][lane]
    // per-segment table of the packed COMP phase
    int seg_lo[32], seg_nc[32];
    float seg_T[32];
    int64_t seg_ray[32];
};
// GENERATED rays (incident / camera): a per-warp queue of the next 32 rays -- 7 words a slot: origin, direction, ray id (a
// forward call traces fewer than 2^31 rays: checked by the launcher) -- lives in the seven deepest shared-memory entries of
// the traversal stack, which then holds SSTK - 7 entries before it spills to local memory (a dedicated 3.5 KB per block
// measured 2 % slower on every path: the L1 that the shared memory carve-out leaves matters more than seven stack entries).
constexpr int RQ_WORDS = 7;

// One lane's share of a queue fill: generate ray `qr` and store it in slot `lane`.  (Keeping this out of line would keep the
// ray arithmetic out of the persistent loop's register allocation, but ptxas 12.9 crashes on a call inside this kernel.)
__device__ __forceinline__ void fill_ray_queue(const TraceArgs &a, int64_t qr, float *rq, int lane) {
    RayCtx t;
    load_ray(a, qr, t);
    rq[0 * 32 + lane] = t.ox; rq[1 * 32 + lane] = t.oy; rq[2 * 32 + lane] = t.oz;
    rq[3 * 32 + lane] = t.dx; rq[4 * 32 + lane] = t.dy; rq[5 * 32 + lane] = t.dz;
    rq[6 * 32 + lane] = __int_as_float((int)qr);
}

template <int NF, bool STATS, bool GEN>
__global__ void __launch_bounds__(TB, IRGS_FWD_BLOCKS) trace_forward_kernel(const KParams p, uint4 *__restrict__ cand_base) {
    constexpr int SS = GEN ? SSTK - RQ_WORDS : SSTK;   // traversal stack entries in shared memory
    constexpr bool FEAT = NF > 0;
    __shared__ __align__(16) WarpSmem<NF> smem[TB / 32];

    const int tid = threadIdx.x;
    const unsigned lane = tid & 31, lt_mask = (1u << lane) - 1u;
    WarpSmem<NF> &ws = smem[tid >> 5];
    int *stk = ws.stack + lane;
    int *pend = ws.pend + lane;
    int stack_spill[STACK - SS];
    const unsigned FULL = 0xffffffffu;
    const TraceArgs &a = p.a;
    const float alpha_min = a.alpha_min, T_min = a.T_min;
    const int back_culling = a.back_culling;
    uint4 *warp_cand = cand_base + ((size_t)blockIdx.x * TB + (tid & ~31)) * KB;   // rows of this warp's 32 lanes
    unsigned long long st_nodes = 0, st_leaf = 0, st_hits = 0, st_pass = 0, st_graze = 0, st_graze_comp = 0, st_comp = 0, st_full = 0;

    int phase = PH_FETCH;
    bool pool_empty = false;  // warp-uniform: the global counter has run past the last ray (the queue may still hold some)
    int q_head = 0, q_cnt = 0; // warp-uniform: the ray queue of this warp
    int64_t ray = 0;
    RayCtx r;
    float T = 1.f;
    float t_last = -INFINITY, t_lo = 0.f, t_hi = IRGS_T_SCENE_MAX;
    int g_last = -1, g_hi = INT_MAX, total = 0, cnt = 0, sp = 0, cur = CUR_NONE, pn = 0;
    bool more = false;        // this pass's depth range was split: another pass follows unless the ray terminates
    r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = 0.f; r.ax = r.ay = r.az = r.bx = r.by = r.bz = 0.f;

    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes
        // (a refill serves whatever lanes are idle: it waits until FETCH_MIN of them are, or until the walk is short of lanes)
        // Rays come from a per-warp QUEUE of 32: when it runs dry the warp takes the next 32 rays off the global counter and all
        // 32 lanes load -- or, for incident / camera rays, GENERATE -- one ray each into shared memory.  Generating at the
        // refill itself ran the ray arithmetic (point record, table entry, angle addition, three divisions) at 7 of 32 lanes and
        // put its dependent loads in front of every refill: trace_incident was 5.3 % behind trace on the same 2^24 rays.
        const unsigned need = __ballot_sync(FULL, phase == PH_FETCH);
        const bool refill = __popc(need) >= FETCH_MIN ||
                            __popc(__ballot_sync(FULL, phase == PH_TRAV && cur != CUR_NONE)) < MIN_ACTIVE;
        if (GEN) {
            float *rq = reinterpret_cast<float *>(ws.stack + SS * 32);   // [word][slot]
            if (need != 0u && refill && !(pool_empty && q_cnt == 0)) {
                const int want = __popc(need), my_rank = __popc(need & lt_mask);
                int given = 0;
                for (int round = 0; round < 2 && given < want; ++round) {
                    if (q_cnt == 0 && !pool_empty) {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(p.counter, 32ull);
                        base = __shfl_sync(FULL, base, 0);
                        int64_t qr = (int64_t)base + lane;
                        __syncwarp();
                        if (qr < a.n_rays) {
                            if (a.ray_mul != 0) qr = (int64_t)(((unsigned)qr * (unsigned)a.ray_mul) % (unsigned)a.n_rays);   // < 2^32
                            fill_ray_queue(a, qr, rq, (int)lane);
                        }
                        __syncwarp();
                        q_head = 0;
                        q_cnt = (int)min((long long)32, max((long long)0, (long long)a.n_rays - (long long)base));
                        if (base + 32ull >= (unsigned long long)a.n_rays) pool_empty = true;
                    }
                    const int take = min(want - given, q_cnt);
                    if (phase == PH_FETCH && my_rank >= given && my_rank < given + take) {
                        const int slot = q_head + (my_rank - given);
                        r.ox = rq[0 * 32 + slot]; r.oy = rq[1 * 32 + slot]; r.oz = rq[2 * 32 + slot];
                        r.dx = rq[3 * 32 + slot]; r.dy = rq[4 * 32 + slot]; r.dz = rq[5 * 32 + slot];
                        ray = __float_as_int(rq[6 * 32 + slot]);
                        ray_setup(r, p.qframe);
                        T = 1.f;
                        t_last = -INFINITY; g_last = -1; total = 0;
                        cnt = 0; sp = 0; pn = 0; cur = 0; t_lo = 0.f; t_hi = IRGS_T_SCENE_MAX; g_hi = INT_MAX; more = false;
                        phase = PH_TRAV;
                        if (STATS) ++st_pass;
                    }
                    q_head += take; q_cnt -= take; given += take;
                    if (pool_empty && q_cnt == 0) break;
                }
            }
        } else if (need != 0u && !pool_empty && refill) {
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(p.counter, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (phase == PH_FETCH) {
                ray = (int64_t)base + __popc(need & lt_mask);
                if (ray < a.n_rays) {
                    if (a.ray_order != nullptr) ray = __ldg(a.ray_order + ray);
                    else if (a.ray_mul != 0) ray = (int64_t)(((unsigned)ray * (unsigned)a.ray_mul) % (unsigned)a.n_rays);   // < 2^32
                    load_ray_mem(a, ray, r);   // (generated rays never reach this kernel variant: launch_trace_forward)
                    ray_setup(r, p.qframe);
                    T = 1.f;
                    t_last = -INFINITY; g_last = -1; total = 0;
                    cnt = 0; sp = 0; pn = 0; cur = 0; t_lo = 0.f; t_hi = IRGS_T_SCENE_MAX; g_hi = INT_MAX; more = false;
                    phase = PH_TRAV;
                    if (STATS) ++st_pass;
                }
            }
            if (base + (unsigned long long)__popc(need) >= (unsigned long long)a.n_rays) pool_empty = true;
        }
        if (__ballot_sync(FULL, phase != PH_FETCH) == 0u) break;  // no rays left (a refill would have handed some out) and every lane idle

        // ------------------------------------------------------------------ BVH walk, NODE sub-phase
        const int thr = (GEN ? (pool_empty && q_cnt == 0) : pool_empty) ? 1 : MIN_ACTIVE;
        unsigned walking = __ballot_sync(FULL, phase == PH_TRAV && cur != CUR_NONE);
        // (a node visit can queue four leaves: the loop is left for the LEAF sub-phase before any queue could overflow)
        bool wide = false;
        while (__popc(walking) >= thr && walking != 0u && !__any_sync(FULL, pn > PQ - 4)) {
            if (DRAIN_MAX > 0 && pool_empty && __popc(walking) <= DRAIN_MAX &&
                !__any_sync(FULL, phase == PH_TRAV && cur != CUR_NONE && sp > DRAIN_SP_MAX)) {
                wide = true;   // the drain of the launch: few long rays left in this warp, see below
                break;
            }
            if (phase == PH_TRAV && cur != CUR_NONE) {
                // one 4-wide node: two 256-bit loads from one 64-byte line, four slab tests
                uint4 c[4];
                ldg256(&p.nodes4[cur].c[0], c[0], c[1]);
                ldg256(&p.nodes4[cur].c[2], c[2], c[3]);
                if (STATS) ++st_nodes;
                // four slab tests; leaves are queued for the LEAF sub-phase, the nearest internal child is walked next and
                // the other internal children are stacked (written with short predicated bodies: the lanes of a warp
                // rarely agree on which children they hit)
                float tk[4];
                bool inner[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float tn;
                    const bool hit = slab(r, c[k], t_lo, t_hi, tn);
                    const int ref = (int)c[k].w;
                    if (hit && ref < 0) { pend[pn * PS] = ref; ++pn; }
                    inner[k] = hit && ref >= 0;
                    tk[k] = inner[k] ? tn : INFINITY;
                }
                // the nearest internal child (the first one among equals) by a select chain ...
                float tbest = INFINITY;
                int next = CUR_NONE, kbest = 4;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool better = tk[k] < tbest;   // tk is finite only for internal children that were hit
                    tbest = better ? tk[k] : tbest; next = better ? (int)c[k].w : next; kbest = better ? k : kbest;
                }
                // ... and the others onto the stack: predicated stores without bound checks while the shared-memory part of the
                // stack has room for all three (the rule), the two-level push otherwise
                if (sp + 3 <= SS) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool push = inner[k] && k != kbest;
                        if (push) stk[sp * 32] = (int)c[k].w;
                        sp += push ? 1 : 0;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (inner[k] && k != kbest) {
                            if (sp < SS) stk[sp * 32] = (int)c[k].w;
                            else if (sp < STACK) stack_spill[sp - SS] = (int)c[k].w;
                            if (sp < STACK) ++sp;
                        }
                    }
                }
                if (next == CUR_NONE && sp > 0) { --sp; next = sp < SS ? stk[sp * 32] : stack_spill[sp - SS]; }
                cur = next;
            }
            walking = __ballot_sync(FULL, phase == PH_TRAV && cur != CUR_NONE);
        }

        // ------------------------------------------------------------------ BVH walk, co-operative wide step (drain)
        // At the end of a launch (work pool empty) a warp is left with a few long rays -- rays that run inside the surfel
        // layer visit thousands of nodes -- and every visit is a dependent L2 access of ONE lane while the others idle:
        // the launch cannot end before its slowest ray (2^14 rays took 0.93 ms, nearly all of it
        // this tail).  Here the idle lanes help: every walking ray gets a group of DRAIN_G lanes which take its current
        // node and the top DRAIN_G - 1 entries of its stack (shared memory), test their four children each against the
        // owner's ray (constants through shuffles) and push the hits back onto the OWNER's stack / leaf queue at offsets
        // from a prefix sum over the group.  Up to DRAIN_G node fetches of one ray are in flight at once.  The walk is no
        // longer strictly near-first, which does not change results: a pass visits everything inside its depth window
        // whatever the order, and rows are ordered by (t, id) before they are composited.
        while (wide) {
            __syncwarp();   // the owners' stack and queue columns are read and written by their helpers
            const int grp = (int)lane / DRAIN_G, mi = (int)lane % DRAIN_G;
            const int n_own = __popc(walking);
            const int owner = grp < n_own ? (int)__fns(walking, 0, grp + 1) : (int)lane;
            const int o_cur = __shfl_sync(FULL, cur, owner), o_sp = __shfl_sync(FULL, sp, owner);
            const int o_pn = __shfl_sync(FULL, pn, owner);
            // nodes this group takes: the current one + stack entries, as many as the leaf queue has room for (4 leaves each)
            const int k_grp = min(min(DRAIN_G, 1 + o_sp), (PQ - o_pn) >> 2);
            const bool work = grp < n_own && mi < k_grp;
            RayCtx ro;
            ro.ax = __shfl_sync(FULL, r.ax, owner); ro.ay = __shfl_sync(FULL, r.ay, owner); ro.az = __shfl_sync(FULL, r.az, owner);
            ro.bx = __shfl_sync(FULL, r.bx, owner); ro.by = __shfl_sync(FULL, r.by, owner); ro.bz = __shfl_sync(FULL, r.bz, owner);
            const float o_tlo = __shfl_sync(FULL, t_lo, owner), o_thi = __shfl_sync(FULL, t_hi, owner);
            unsigned in_mask = 0u, lf_mask = 0u;
            int ref[4] = {0, 0, 0, 0};
            if (work) {
                const int node = mi == 0 ? o_cur : ws.stack[(o_sp - mi) * 32 + owner];
                uint4 c[4];
                ldg256(&p.nodes4[node].c[0], c[0], c[1]);
                ldg256(&p.nodes4[node].c[2], c[2], c[3]);
                if (STATS) ++st_nodes;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float tn;
                    const bool hit = slab(ro, c[k], o_tlo, o_thi, tn);
                    ref[k] = (int)c[k].w;
                    if (hit && ref[k] >= 0) in_mask |= 1u << k;
                    if (hit && ref[k] < 0) lf_mask |= 1u << k;
                }
            }
            // inclusive prefix sums of the push counts over the DRAIN_G lanes of a group
            int in_incl = __popc(in_mask), lf_incl = __popc(lf_mask);
#pragma unroll
            for (int o = 1; o < DRAIN_G; o <<= 1) {
                const int vi = __shfl_up_sync(FULL, in_incl, o, DRAIN_G), vl = __shfl_up_sync(FULL, lf_incl, o, DRAIN_G);
                if (mi >= o) { in_incl += vi; lf_incl += vl; }
            }
            __syncwarp();   // every helper has read its stack entry before anything is pushed
            if (work) {
                const int s0 = o_sp - (k_grp - 1) + in_incl - __popc(in_mask);   // the group popped k_grp - 1 entries
                const int q0 = o_pn + lf_incl - __popc(lf_mask);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((in_mask >> k) & 1u) ws.stack[(s0 + __popc(in_mask & ((1u << k) - 1u))) * 32 + owner] = ref[k];
                    if ((lf_mask >> k) & 1u) ws.pend[(q0 + __popc(lf_mask & ((1u << k) - 1u))) * PS + owner] = ref[k];
                }
            }
            // the owners take the totals of their group (held by its last lane) and pop the next node
            const bool own = phase == PH_TRAV && cur != CUR_NONE;
            const int my_grp_last = own ? __popc(walking & lt_mask) * DRAIN_G + DRAIN_G - 1 : (int)lane;
            const int tot_in = __shfl_sync(FULL, in_incl, my_grp_last), tot_lf = __shfl_sync(FULL, lf_incl, my_grp_last);
            __syncwarp();   // pushes are visible to the owners
            if (own) {
                const int k_own = min(min(DRAIN_G, 1 + sp), (PQ - pn) >> 2);
                if (k_own >= 1) {
                    sp = sp - (k_own - 1) + tot_in;
                    pn += tot_lf;
                    if (sp > 0) { --sp; cur = stk[sp * 32]; }
                    else cur = CUR_NONE;
                }
            }
            // next wide step unless a leaf queue is about to fill up (LEAF sub-phase first) or the walks are done
            walking = __ballot_sync(FULL, phase == PH_TRAV && cur != CUR_NONE);
            wide = walking != 0u && !__any_sync(FULL, pn > PQ - 4) &&
                   !__any_sync(FULL, phase == PH_TRAV && cur != CUR_NONE && sp > DRAIN_SP_MAX);
        }

        // ------------------------------------------------------------------ BVH walk, LEAF sub-phase
        // The pending leaves of ALL lanes form one warp-wide list (prefix sum of the queue lengths) that is tested 32
        // items at a time, one (ray, surfel) pair per lane whoever owns the ray: item -> owner by a binary search over
        // the prefix sums (shuffles), the leaf straight from the owner's queue column in shared memory, the owner's ray
        // and depth window through shuffles.  (Each lane draining its own queue ran this code at 5.7 of 32 lanes:
        // profiles/r01_fwd_regions_v7q.txt.)  A lane submits at most as many leaves as its candidate row has room
        // for; the rest stay queued (PH_FULL).
        for (;;) {
            const int m = (phase == PH_TRAV) ? min(pn, KB - cnt) : 0;
            if (!__any_sync(FULL, m > 0)) break;
            int incl = m;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if ((int)lane >= o) incl += v;
            }
            const int start = incl - m;
            const int n_items = __shfl_sync(FULL, incl, 31);
            __syncwarp();   // queue columns written in the NODE sub-phase are read by other lanes below
#if IRGS_LEAF_PIPELINE
            // item -> (owner, leaf) is static for the whole batch (queues and prefix sums do not change inside it): the item of the
            // NEXT round is located while this round's record loads are in flight, so that a round starts with its loads instead
            // of seven dependent shuffles and a shared-memory read
            auto locate_item = [&](int idx, int &owner_out, unsigned &pk_out, int &leaf_out) {
                const bool in = idx < n_items;
                int o = 0;
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int v = __shfl_sync(FULL, incl, o + step - 1);
                    if (v <= idx) o += step;
                }
                o = in ? o : (int)lane;
                const unsigned pk = __shfl_sync(FULL, (unsigned)start | ((unsigned)m << 10) | ((unsigned)pn << 16), o);
                owner_out = o; pk_out = pk;
                leaf_out = in ? ~ws.pend[((int)((pk >> 16) & 63u) - 1 - (idx - (int)(pk & 1023u))) * PS + o] : 0;   // popped from the top
            };
            int owner_n, leaf_n; unsigned pk_n;
            locate_item((int)lane, owner_n, pk_n, leaf_n);
#endif
            for (int base = 0; base < n_items; base += 32) {
                const int idx = base + (int)lane;
                const bool has = idx < n_items;
#if IRGS_LEAF_PIPELINE
                const int owner = owner_n, leaf = leaf_n;
                const int o_start = (int)(pk_n & 1023u), o_m = (int)((pk_n >> 10) & 63u);
                float4 q0, q1, q2, q3;
                q0 = q1 = q2 = q3 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has) {
                    ldg256_stream(p.recs + leaf, q0, q1);
                    if (!LEAF_2STAGE) ldg256_stream(&p.recs[leaf].r2, q2, q3);
                    if (STATS) ++st_leaf;
                }
                if (base + 32 < n_items) locate_item(idx + 32, owner_n, pk_n, leaf_n);
                const int o_cnt = __shfl_sync(FULL, cnt, owner);   // (changes from round to round: fetched where it is used)
#else
                int owner = 0;
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int v = __shfl_sync(FULL, incl, owner + step - 1);
                    if (v <= idx) owner += step;
                }
                owner = has ? owner : (int)lane;
#if IRGS_LEAF_PACK
                // (start < 2^10, m / pn / cnt <= 32: one shuffle instead of four)
                const unsigned o_pk = __shfl_sync(FULL, (unsigned)start | ((unsigned)m << 10) | ((unsigned)pn << 16) | ((unsigned)cnt << 22), owner);
                const int o_start = (int)(o_pk & 1023u), o_m = (int)((o_pk >> 10) & 63u);
                const int o_pn = (int)((o_pk >> 16) & 63u), o_cnt = (int)(o_pk >> 22);
#else
                const int o_start = __shfl_sync(FULL, start, owner), o_m = __shfl_sync(FULL, m, owner);
                const int o_pn = __shfl_sync(FULL, pn, owner), o_cnt = __shfl_sync(FULL, cnt, owner);
#endif
                float4 q0, q1, q2, q3;
                q0 = q1 = q2 = q3 = make_float4(0.f, 0.f, 0.f, 0.f);
                int leaf = 0;
                if (has) {
                    leaf = ~ws.pend[(o_pn - 1 - (idx - o_start)) * PS + owner];   // popped from the top
                    ldg256_stream(p.recs + leaf, q0, q1);
                    if (!LEAF_2STAGE) ldg256_stream(&p.recs[leaf].r2, q2, q3);
                    if (STATS) ++st_leaf;
                }
#endif
                RayCtx ro;
                ro.ox = __shfl_sync(FULL, r.ox, owner); ro.oy = __shfl_sync(FULL, r.oy, owner); ro.oz = __shfl_sync(FULL, r.oz, owner);
                ro.dx = __shfl_sync(FULL, r.dx, owner); ro.dy = __shfl_sync(FULL, r.dy, owner); ro.dz = __shfl_sync(FULL, r.dz, owner);
                const float o_tlast = __shfl_sync(FULL, t_last, owner), o_thi = __shfl_sync(FULL, t_hi, owner);
                float t, alpha = 0.f, px, py, pz; int g;
#if IRGS_LEAF_PACK
                // the pass window is a pair of (t, id) keys; the ids decide only when a depth equals a bound bit for bit, so they
                // are fetched (two more shuffles) only in the rounds in which some lane needs them
                bool ok = has && leaf_stage1(ro, q0, q1, back_culling, t, g, px, py, pz) && !(t < o_tlast) && !(o_thi < t);
                if (__any_sync(FULL, ok && (t == o_tlast || t == o_thi))) {
                    const int o_glast = __shfl_sync(FULL, g_last, owner), o_ghi = __shfl_sync(FULL, g_hi, owner);
                    ok = ok && key_less(o_tlast, o_glast, t, g) && key_less(t, g, o_thi, o_ghi);
                }
#else
                const int o_glast = __shfl_sync(FULL, g_last, owner), o_ghi = __shfl_sync(FULL, g_hi, owner);
                bool ok = has && leaf_stage1(ro, q0, q1, back_culling, t, g, px, py, pz) &&
                          key_less(o_tlast, o_glast, t, g) && key_less(t, g, o_thi, o_ghi);
#endif
                if (ok) {
                    // fetching the second half of the record only now saves L1TEX wavefronts but serialises two L2
                    // latencies: measured 2 % slower than issuing both loads up front (profiles/r01_sweeps.txt)
                    if (LEAF_2STAGE) ldg256_stream(&p.recs[leaf].r2, q2, q3);
                    ok = leaf_stage2(q2, q3, px, py, pz, alpha_min, alpha);
                }
                if (STATS && has && !ok) {   // diagnostics: grazing pairs dropped by the hit test (first pass of a ray only)
                    float4 s2 = q2, s3 = q3;
                    if (LEAF_2STAGE) ldg256_stream(&p.recs[leaf].r2, s2, s3);
                    bool comp = false;
                    if (!(o_tlast > -INFINITY) && grazing_dropped(ro, q0, q1, s2, s3, alpha_min, &comp)) { ++st_graze; if (comp) ++st_graze_comp; }
                }
                const unsigned acc = __ballot_sync(FULL, ok);
                // the items of one owner are contiguous in the list: bits [lo, hi) of this round
                if (ok) {
                    const int lo = max(o_start - base, 0), hi = min(o_start + o_m - base, 32);
                    const unsigned rmask = (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
                    warp_cand[(size_t)owner * KB + o_cnt + __popc(acc & rmask & lt_mask)] =
                        make_uint4(__float_as_uint(t), (unsigned)g, __float_as_uint(alpha), (unsigned)leaf);
                }
                {
                    const int lo = min(max(start - base, 0), 32), hi = min(max(start + m - base, 0), 32);
                    const unsigned hm = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
                    const unsigned lm = lo >= 32 ? 0xffffffffu : ((1u << lo) - 1u);
                    cnt += __popc(acc & hm & ~lm);
                }
            }
            pn -= m;
            if (phase == PH_TRAV && cnt == KB) phase = PH_FULL;   // pending leaves stay queued until the row has been trimmed
        }
        // the walk of this pass is complete once the stack, the current node and the leaf queue are all empty
        if (phase == PH_TRAV && cur == CUR_NONE && pn == 0) phase = PH_COMP;

        // ------------------------------------------------------------------ FULL rows: co-operative sort + trim / split
        unsigned work = __ballot_sync(FULL, phase == PH_FULL);
        while (work != 0u) {
            const int L = __ffs(work) - 1;
            work &= work - 1u;
            if (STATS && lane == 0) ++st_full;
            // lane i takes candidate i, ranks it by (t, surfel id), and the candidates are permuted into depth order
            __syncwarp();  // lane L's appended candidates are visible to the whole warp
            // (requesting the next waiting row before this one is sorted measured nothing: profiles/r02_sweeps.txt)
            const uint4 e = __ldcg(&warp_cand[(size_t)L * KB + lane]);
            float my_t = __uint_as_float(e.x), my_a = __uint_as_float(e.z); int my_g = (int)e.y, my_p = (int)e.w;
            // (depths are positive and surfel ids non-negative: the (t, id) order is the order of the packed 64-bit keys)
            const unsigned long long my_key = ((unsigned long long)__float_as_uint(my_t) << 32) | (unsigned)my_g;
            int rank = 0;
#if IRGS_FULL_RANK_SMEM
            {   // all 32 keys through shared memory, two per broadcast load (16 LDS.128 instead of 64 shuffles)
                unsigned long long *s_k = reinterpret_cast<unsigned long long *>(ws.scratch);
                s_k[lane] = my_key;
                __syncwarp();
#pragma unroll
                for (int j = 0; j < KB; j += 2) {
                    const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(s_k + j);
                    rank += (kk.x < my_key) ? 1 : 0;
                    rank += (kk.y < my_key) ? 1 : 0;
                }
                __syncwarp();   // the keys have been read before the scratch rows are reused below
            }
#else
            for (int j = 0; j < KB; ++j) {
                const unsigned tj = __shfl_sync(FULL, __float_as_uint(my_t), j);
                const unsigned gj = __shfl_sync(FULL, (unsigned)my_g, j);
                rank += ((((unsigned long long)tj << 32) | gj) < my_key) ? 1 : 0;
            }
#endif
            {
                float *s_f = ws.scratch;
                int *s_i = reinterpret_cast<int *>(ws.scratch);
                s_f[rank] = my_t; s_i[32 + rank] = my_g; s_f[64 + rank] = my_a; s_i[96 + rank] = my_p;
                __syncwarp();
                my_t = s_f[lane]; my_g = s_i[32 + lane]; my_a = s_f[64 + lane]; my_p = s_i[96 + lane];
#if !IRGS_FULL_CHAIN_SMEM
                __syncwarp();
#endif
            }
            // transmittance chain in the reference's sequential order
            float Tc = __shfl_sync(FULL, T, L);
            int n_comp = KB;
            bool term = false;
#if IRGS_FULL_CHAIN_SMEM
            // (the ordered alphas are still in the scratch row: four per broadcast load instead of one shuffle each)
            for (int i = 0; i < KB && !term; i += 4) {
                const float4 a4 = *reinterpret_cast<const float4 *>(ws.scratch + 64 + i);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!term) {
                        Tc *= (1.f - av[u]);
                        if (Tc < T_min) { n_comp = i + u + 1; term = true; }
                    }
                }
            }
            __syncwarp();   // the scratch row may be overwritten from here on
#else
            for (int i = 0; i < KB; ++i) {
                const float ai = __shfl_sync(FULL, my_a, i);
                Tc *= (1.f - ai);
                if (Tc < T_min) { n_comp = i + 1; term = true; break; }
            }
#endif
            // terminated: keep the composited prefix and clip the walk to it; otherwise split the depth range at the
            // KB/2-th candidate and leave the rest of the ray to a following pass
            // (a row that terminates only at its very last entry cannot be trimmed: it is split like any other)
            const bool split = !(term && n_comp < KB);
            const int keep = split ? KB / 2 : n_comp;
            const float t_end = __shfl_sync(FULL, my_t, keep - 1);
            const int g_end = __shfl_sync(FULL, my_g, keep - 1);
            if ((int)lane < keep)
                warp_cand[(size_t)L * KB + lane] = make_uint4(__float_as_uint(my_t), (unsigned)my_g, __float_as_uint(my_a), (unsigned)my_p);
            if ((int)lane == L) {
                cnt = keep; t_hi = t_end; g_hi = g_end;
                if (split) more = true;
                // resume the walk; if nothing is left to walk the row is composited right away
                phase = (cur == CUR_NONE && pn == 0) ? PH_COMP : PH_TRAV;
            }
        }

        // ------------------------------------------------------------------ lanes whose pass found nothing
        if (phase == PH_COMP && cnt == 0) {
            // outputs are pre-zeroed; a multi-pass ray already added its earlier passes
            if (total > 0 && a.hit_count != nullptr) a.hit_count[ray] = total;
            if (STATS) st_hits += total;
            phase = PH_FETCH;
        }

        // ------------------------------------------------------------------ co-operative sort + composite, several rays at once
        // The finished rows of as many lanes as fit are packed into the 32 lanes (one candidate per lane, a ray's
        // candidates in consecutive lanes = a segment); ranking, the transmittance chain and the accumulation run per
        // segment, all segments side by side.  (One ray at a time, these loops were 30 % of all issued instructions with
        // 12.7 of 32 lanes holding a candidate on average.)
        // Compositing waits until enough candidates have piled up to fill most of a packed round, or until the walk is
        // about to run short of lanes anyway.
        unsigned comp = __ballot_sync(FULL, phase == PH_COMP);
        if (comp != 0u) {
            const int waiting = __reduce_add_sync(FULL, phase == PH_COMP ? cnt : 0);
            const int still_walking = __popc(__ballot_sync(FULL, phase == PH_TRAV && cur != CUR_NONE));
            if (waiting < COMP_MIN && still_walking >= thr) comp = 0u;
        }
        while (comp != 0u) {
            __syncwarp();  // appended candidates are visible to the whole warp
            const bool is_comp = (comp >> lane) & 1u;
            int incl = is_comp ? cnt : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if ((int)lane >= o) incl += v;
            }
            const unsigned gmask = __ballot_sync(FULL, is_comp && incl <= 32);   // a prefix of the waiting lanes; never empty
            comp &= ~gmask;
            const int g_last_lane = 31 - __clz(gmask);
            const int gtotal = __shfl_sync(FULL, incl, g_last_lane);
            incl = min(incl, gtotal);
            const bool in_group = (gmask >> lane) & 1u;
            const int c2 = in_group ? cnt : 0;
            const int start2 = incl - c2;
            const int maxn = __reduce_max_sync(FULL, c2);
            if (STATS && lane == 0) { st_comp += ((unsigned long long)maxn << 32) + (unsigned long long)gtotal; st_full += 1ull << 32; }   // sum of the longest segment | candidates; rounds
            // candidate of this lane: owner = number of lanes whose inclusive count is <= lane
            const bool act = (int)lane < gtotal;
            int owner = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int v = __shfl_sync(FULL, incl, owner + step - 1);
                if (v <= (int)lane) owner += step;
            }
            owner = act ? owner : (int)lane;
            const int o_start2 = __shfl_sync(FULL, start2, owner), o_c2 = __shfl_sync(FULL, c2, owner);
            const int seg_lo = act ? o_start2 : (int)lane;
            const int n_o = act ? o_c2 : 0;
            int k = (int)lane - seg_lo;
            float my_t = INFINITY, my_a = 0.f; int my_g = INT_MAX, my_p = 0;
            if (act) {
                const uint4 e = __ldcg(&warp_cand[(size_t)owner * KB + k]);
                my_t = __uint_as_float(e.x); my_g = (int)e.y; my_a = __uint_as_float(e.z); my_p = (int)e.w;
            }
            const unsigned long long my_key = ((unsigned long long)__float_as_uint(my_t) << 32) | (unsigned)my_g;
            // (the same loop through shared memory -- one 64-bit load per step, addresses differing by segment -- measured 3 % slower
            //  than the two shuffles; only the full-row sort above, whose loads are broadcasts, gains from it.  Ranking by depth alone
            //  -- one shuffle per step, the (t, id) loop only after a tie -- also measured 3 % slower; so did nothing at all a special
            //  case for rounds that hold ONE ray, ranked and chained through broadcast loads like a full row: profiles/r02_sweeps.txt)
            int rank = 0;
            for (int j = 0; j < maxn; ++j) {
                const int src = (seg_lo + j) & 31;
                const unsigned tj = __shfl_sync(FULL, __float_as_uint(my_t), src);
                const unsigned gj = __shfl_sync(FULL, (unsigned)my_g, src);
                rank += (j < n_o && (((unsigned long long)tj << 32) | gj) < my_key) ? 1 : 0;
            }
            {
                float *s_f = ws.scratch;
                int *s_i = reinterpret_cast<int *>(ws.scratch);
                if (act) { s_f[seg_lo + rank] = my_t; s_i[32 + seg_lo + rank] = my_g; s_f[64 + seg_lo + rank] = my_a; s_i[96 + seg_lo + rank] = my_p; }
                __syncwarp();
                if (act) { my_t = s_f[lane]; my_g = s_i[32 + lane]; my_a = s_f[64 + lane]; my_p = s_i[96 + lane]; }
                __syncwarp();
            }
            // transmittance chain of every segment in the reference's sequential order (bit-identical termination decisions)
            float Tc = __shfl_sync(FULL, T, owner);
            float my_w = 0.f;
            int n_comp = n_o;
            bool term = false;
            for (int i = 0; i < maxn; ++i) {
                const float ai = __shfl_sync(FULL, my_a, (seg_lo + i) & 31);
                if (i < n_o && !term) {
                    if (k == i) my_w = Tc * ai;
                    Tc *= (1.f - ai);
                    if (Tc < T_min) { n_comp = i + 1; term = true; }
                }
            }
            const float dxl = __shfl_sync(FULL, r.dx, owner), dyl = __shfl_sync(FULL, r.dy, owner), dzl = __shfl_sync(FULL, r.dz, owner);
            const int64_t ray_o = __shfl_sync(FULL, ray, owner);
            const int total_o = __shfl_sync(FULL, total, owner);
            // shade one hit per lane
            float c0 = 0.f, c1 = 0.f, c2c = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, dd = 0.f, oo = 0.f;
            float f[FEAT ? NF : 1];
#pragma unroll
            for (int j = 0; j < (FEAT ? NF : 1); ++j) f[j] = 0.f;
            const bool mine = act && k < n_comp;
            if (mine) {
                float Y[16];
                sh_basis(a.deg, dxl, dyl, dzl, Y);
                const float4 rn = __ldg(&p.recs[my_p].r1);   // the surfel's normal, from its packed record
                const float nx = rn.x, ny = rn.y, nz = rn.z;
                const float dg = dot3_rn(nx, ny, nz, dxl, dyl, dzl);
                const float m = (-dg > 0.f) ? 1.f : -1.f;
                float c[3];
                sh_color(a.shs, a.K, a.deg, my_g, Y, c);
                c0 = my_w * c[0]; c1 = my_w * c[1]; c2c = my_w * c[2];
                n0 = my_w * m * nx; n1 = my_w * m * ny; n2 = my_w * m * nz;
                dd = my_w * my_t; oo = my_w;
                if (FEAT) {
#pragma unroll
                    for (int j = 0; j < NF; ++j)
                        if (j < a.S) f[j] = my_w * __ldg(a.features + (size_t)my_g * a.S + j);
                }
                if (a.hits != nullptr && total_o + k < a.hit_cap) a.hits[ray_o * a.hit_cap + total_o + k] = my_g;
                if (a.hit_rgb != nullptr && total_o + k < a.rgb_cap) {   // colour cache of the backward replay (streamed: read once)
                    float *hc = a.hit_rgb + ((size_t)ray_o * a.rgb_cap + total_o + k) * 3;
                    __stcs(hc, c[0]); __stcs(hc + 1, c[1]); __stcs(hc + 2, c[2]);
                }
            }
            // Sequential accumulation in depth order, exactly like the reference's loop (and the oracle's): one lane per
            // (ray, output channel) adds that ray's n_comp terms one after the other to the running value in global
            // memory (pre-zeroed by the launcher).  The result is therefore independent of how the hits were grouped
            // into passes and of the acceleration structure's topology.
            {
                float *s_c = ws.scratch;   // [channel][lane]
                s_c[0 * CS + lane] = c0; s_c[1 * CS + lane] = c1; s_c[2 * CS + lane] = c2c;
                s_c[3 * CS + lane] = n0; s_c[4 * CS + lane] = n1; s_c[5 * CS + lane] = n2;
                s_c[6 * CS + lane] = dd; s_c[7 * CS + lane] = oo;
                if (FEAT) {
#pragma unroll
                    for (int j = 0; j < NF; ++j)
                        if (j < a.S) s_c[(8 + j) * CS + lane] = f[j];
                }
                // per-segment table, written by the first lane of each segment
                if (act && k == 0) {
                    const int sidx = __popc(gmask & ((1u << owner) - 1u));
                    // bit 8: nothing of this ray has been composited yet (98 % of the rays that hit anything need one pass)
                    ws.seg_lo[sidx] = seg_lo | (total_o == 0 ? 256 : 0); ws.seg_nc[sidx] = term ? -n_comp : n_comp; ws.seg_T[sidx] = Tc;
                    ws.seg_ray[sidx] = ray_o;
                }
                __syncwarp();
                const int n_ch = 8 + (FEAT ? a.S : 0);
                const int pairs = __popc(gmask) * n_ch;
                for (int pr = (int)lane; pr < pairs; pr += 32) {
                    const int sidx = pr / n_ch, ch = pr - sidx * n_ch;
                    const int slo = ws.seg_lo[sidx], lo = slo & 255, nc = abs(ws.seg_nc[sidx]);
                    const int64_t rr = ws.seg_ray[sidx];
                    float *dst;
                    if (ch < 3) dst = a.color + 3 * rr + ch;
                    else if (ch < 6) dst = a.normal + 3 * rr + (ch - 3);
                    else if (ch == 6) dst = a.depth + rr;
                    else if (ch == 7) dst = a.alpha + rr;
                    else dst = a.feature + rr * a.S + (ch - 8);
                    // the outputs are pre-zeroed: a ray's first pass adds to 0 without reading them back (a dependent L2 access)
                    // (L2-only __ldcg / __stcg here measured 12 % SLOWER on the C3 step: profiles/r01_sweeps.txt)
                    float accv = (slo & 256) ? 0.f : *dst;
                    for (int i = 0; i < nc; ++i) accv += s_c[ch * CS + lo + i];
                    *dst = accv;
                }
            }
            if (in_group) {
                const int sidx = __popc(gmask & lt_mask);
                const int nc_signed = ws.seg_nc[sidx];
                const bool term_L = nc_signed < 0;
                T = ws.seg_T[sidx];
                total += abs(nc_signed);
                if (!term_L && more) {
                    // this pass covered the depth range up to (t_hi, g_hi) completely without terminating: the next
                    // pass continues strictly after it
                    t_last = t_hi; g_last = g_hi;
                    cnt = 0; sp = 0; pn = 0; cur = 0; t_lo = fmaxf(t_last, 0.f); t_hi = IRGS_T_SCENE_MAX; g_hi = INT_MAX;
                    more = false;
                    phase = PH_TRAV;
                    if (STATS) ++st_pass;
                } else {
                    if (a.hit_count != nullptr) a.hit_count[ray] = total;
                    if (STATS) st_hits += total;
                    phase = PH_FETCH;
                }
            }
            __syncwarp();
        }
    }
    if (STATS) {
        atomicAdd(p.stats + 0, st_nodes); atomicAdd(p.stats + 1, st_leaf);
        atomicAdd(p.stats + 2, st_hits); atomicAdd(p.stats + 3, st_pass);
        atomicAdd(p.stats + 4, st_graze); atomicAdd(p.stats + 5, st_graze_comp);
        atomicAdd(p.stats + 6, st_comp); atomicAdd(p.stats + 7, st_full);
    }
}

template <typename Kern>
static int launch_fwd(irgs_tracer *h, int slot, Kern kern, const KParams &p, int64_t n_rays, cudaStream_t s) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TB, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    if (h->fwd_blocks_per_sm > 0 && h->fwd_blocks_per_sm < per_sm) per_sm = h->fwd_blocks_per_sm;   // irgs_set_option: experiments
    int grid = h->sm_count * per_sm;
    // Shared-memory carve-out: exactly what the resident blocks need, so that the rest of the 256 KB stays L1 cache (the top
    // levels of the tree are read by every ray; with the default maximum carve-out the L1 hit rate of the walk was 3.8 %).
    {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess) {
            const double need_kb = per_sm * ((double)fa.sharedSizeBytes + 1024.0) / 1024.0;
            int pct = (int)(100.0 * need_kb / 228.0) + 1;
            if (pct > 100) pct = 100;
            if (h->carveout_pct >= 0) pct = h->carveout_pct;   // irgs_set_option("smem_carveout_pct"): experiments
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
    }
    // candidate scratch: one 32-entry row (512 B) per resident thread of this stream slot; only rows in use are live in L2
    const int64_t threads = (int64_t)grid * TB;
    if (threads > h->cand_threads[slot]) {
        IRGS_CHECK(cudaStreamSynchronize(s));   // earlier launches of this slot (all on this stream) are done with the old block
        if (h->cand[slot]) cudaFree(h->cand[slot]);
        h->cand[slot] = nullptr;
        IRGS_CHECK(cudaMalloc(&h->cand[slot], sizeof(uint4) * KB * (size_t)threads));
        h->cand_threads[slot] = threads;
    }
    const int64_t need = (n_rays + TB - 1) / TB;
    if (need < grid) grid = (int)(need > 0 ? need : 1);
    IRGS_CHECK(cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), s));
    kern<<<grid, TB, 0, s>>>(p, h->cand[slot]);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

// Multiplier of the stride start order of small launches: the i-th ray started is (i * m) mod n.  m is odd, coprime to n
// (a bijection on [0, n)) and small enough that i * m < 2^32 for every i < n <= 2^19 (the kernel multiplies in 32 bits);
// ~0.618 * 2^13, so that consecutive starts are ~20 pixel bundles apart.  0: keep the caller's order.
int64_t stride_multiplier(int64_t n_rays) {
    if (n_rays < 64 || n_rays > ((int64_t)1 << 19)) return 0;
    int64_t m = 5063;
    auto gcd = [](int64_t x, int64_t y) { while (y) { const int64_t t = x % y; x = y; y = t; } return x; };
    while (gcd(m, n_rays) != 1) m += 2;
    return m < 8192 ? m : 0;
}

int launch_trace_forward(irgs_tracer *h, const TraceArgs &a, cudaStream_t s) {
    if (a.n_rays >= ((int64_t)1 << 31)) { set_error("a forward call traces fewer than 2^31 rays: split the batch"); return 1; }
    const int slot = slot_for(h, s);
    if (slot < 0) return 1;
    KParams p;
    p.a = a;
    // Generated rays (incident / camera): by default a small kernel writes them into a scratch block of this stream slot first
    // and the forward kernel reads them like any other rays.  Measured on 2^24 C3 rays, same box: rays from memory 20.96 ms,
    // generated inside the persistent kernel (ray queue, full lanes) 22.5 ms, generated at the refill 23.1 ms -- DRAM is 4 % busy,
    // the 24 B per ray are free, while every instruction and register inside the latency- and issue-bound walk is not.  The
    // caller-visible contract is unchanged (no ray arrays in, 28 B per point); irgs_set_option("gen_in_kernel", 1) keeps the
    // in-kernel generation (no scratch: 24 B per ray of a call less memory).
    if ((a.gen_pos != nullptr || a.cam_W > 0) && !h->gen_in_kernel) {
        if (a.n_rays > h->ray_scratch_cap[slot]) {
            IRGS_CHECK(cudaStreamSynchronize(s));
            if (h->ray_scratch[slot]) cudaFree(h->ray_scratch[slot]);
            h->ray_scratch[slot] = nullptr;
            IRGS_CHECK(cudaMalloc(&h->ray_scratch[slot], sizeof(float) * 6 * (size_t)a.n_rays));
            h->ray_scratch_cap[slot] = a.n_rays;
        }
        float *so = h->ray_scratch[slot], *sd = so + 3 * (size_t)a.n_rays;
        if (launch_generated_rays(a, so, sd, s)) return 1;
        p.a.rays_o = so; p.a.rays_d = sd;
        p.a.gen_pos = nullptr; p.a.cam_W = 0;
    }
    p.nodes = h->qnodes; p.nodes4 = h->qnodes4; p.qframe = h->scene + 12; p.recs = h->recs; p.inv_order = h->inv_order; p.counter = h->counter + slot; p.stats = h->stats;
    // Colour cache for the backward replay (internal.cuh): a call that saves hit lists also leaves the colours of each ray's first
    // color_cache hits in a block of this stream slot.  Failing to get the memory only means the backward gathers SH rows.
    p.a.hit_rgb = nullptr; p.a.rgb_cap = 0;
    if (a.hits != nullptr && a.hit_count != nullptr && h->color_cache > 0) {
        const int cc = h->color_cache < a.hit_cap ? h->color_cache : a.hit_cap;
        const int64_t need = a.n_rays * cc * 3;
        std::lock_guard<std::mutex> lock(h->slot_mutex);
        if (need > h->hit_rgb_floats[slot]) {
            IRGS_CHECK(cudaStreamSynchronize(s));   // earlier users of the old block all ran on this stream
            if (h->hit_rgb[slot]) cudaFree(h->hit_rgb[slot]);
            h->hit_rgb[slot] = nullptr; h->hit_rgb_floats[slot] = 0; h->hit_rgb_key[slot] = nullptr;
            if (cudaMalloc(&h->hit_rgb[slot], sizeof(float) * (size_t)need) == cudaSuccess) h->hit_rgb_floats[slot] = need;
            else { h->hit_rgb[slot] = nullptr; cudaGetLastError(); }
        }
        if (h->hit_rgb[slot] != nullptr && cc > 0) {
            for (int i = 0; i < irgs_tracer::MAX_SLOTS; ++i)
                if (h->hit_rgb_key[i] == a.hits) h->hit_rgb_key[i] = nullptr;   // that list is being overwritten
            h->hit_rgb_key[slot] = a.hits; h->hit_rgb_rays[slot] = a.n_rays; h->hit_rgb_cc[slot] = cc;
            p.a.hit_rgb = h->hit_rgb[slot]; p.a.rgb_cap = cc;
        }
    } else if (a.hits != nullptr) {
        std::lock_guard<std::mutex> lock(h->slot_mutex);
        for (int i = 0; i < irgs_tracer::MAX_SLOTS; ++i)
            if (h->hit_rgb_key[i] == a.hits) h->hit_rgb_key[i] = nullptr;
    }
    const bool stats = h->stats_enabled != 0;
    if (stats) IRGS_CHECK(cudaMemsetAsync(h->stats, 0, 8 * sizeof(unsigned long long), s));
    // (generated rays -- a.gen_pos -- have no ray arrays to take sort keys from and arrive grouped per shading point anyway)
    if (h->sort_rays_min > 0 && a.n_rays >= h->sort_rays_min && a.n_rays < ((int64_t)1 << 31) && a.gen_pos == nullptr &&
        a.rays_o != nullptr && a.rays_d != nullptr) {
        int *order = nullptr;
        if (launch_ray_order(h, slot, a.rays_o, a.rays_d, a.n_rays, &order, s)) return 1;
        p.a.ray_order = order;
    }
    // Small launches are latency bound: the call ends with its slowest warp, and the heavy rays (grazing samples that run
    // inside the surfel layer: ~300 node visits, ~40 hits, several passes) come clustered -- the caller lays rays out
    // bundle by bundle, so a grazing pixel puts many of them into the same warp, where their phases serialise.  Starting
    // the rays in a stride order (the i-th ray started is (i * m) mod n, m odd and coprime to n) spreads them over the
    // warps: 2^12 / 2^14 / 2^16 / 2^18 / 2^19 rays 0.84 / 0.92 / 1.01 / 1.14 / 1.41 ms -> 0.62 / 0.63 / 0.73 / 1.00 / 1.36 ms
    // (scripts/stride_check.py); from 2^20 rays on throughput counts and the orders tie (2.19 ms), so large launches keep
    // the caller's order.  Results are written per ray id either way.
    if (p.a.ray_order == nullptr && a.n_rays <= h->stride_rays_max) p.a.ray_mul = stride_multiplier(a.n_rays);
    // rays that hit nothing are never written by the kernel: the outputs are zero-filled first -- with ONE memset when the caller
    // laid them out back to back (color, normal, feature, depth, alpha, hit_count: what irgs_b200/raytracer.py allocates), which
    // matters for the small calls IRGS itself issues (2^18 rays: six memsets were 0.05 ms of a 0.95 ms call)
    const size_t R = (size_t)a.n_rays;
    {
        char *p0 = reinterpret_cast<char *>(a.color);
        const size_t o_n = 12 * R, o_f = 24 * R, o_d = o_f + 4 * (size_t)a.S * R, o_a = o_d + 4 * R, o_h = o_a + 4 * R;
        // (only on the caller's word -- irgs_set_option("contiguous_outputs", 1): arrays that merely HAPPEN to be adjacent may belong
        // to different allocations, which one memset must not span)
        const bool contiguous = h->contiguous_outputs && reinterpret_cast<char *>(a.normal) == p0 + o_n && (a.S == 0 || reinterpret_cast<char *>(a.feature) == p0 + o_f) &&
                                reinterpret_cast<char *>(a.depth) == p0 + o_d && reinterpret_cast<char *>(a.alpha) == p0 + o_a &&
                                (a.hit_count == nullptr || reinterpret_cast<char *>(a.hit_count) == p0 + o_h);
        if (contiguous) {
            IRGS_CHECK(cudaMemsetAsync(p0, 0, o_h + (a.hit_count ? 4 * R : 0), s));
        } else {
            IRGS_CHECK(cudaMemsetAsync(a.color, 0, sizeof(float) * 3 * R, s));
            IRGS_CHECK(cudaMemsetAsync(a.normal, 0, sizeof(float) * 3 * R, s));
            IRGS_CHECK(cudaMemsetAsync(a.depth, 0, sizeof(float) * R, s));
            IRGS_CHECK(cudaMemsetAsync(a.alpha, 0, sizeof(float) * R, s));
            if (a.S > 0) IRGS_CHECK(cudaMemsetAsync(a.feature, 0, sizeof(float) * (size_t)a.S * R, s));
            if (a.hit_count) IRGS_CHECK(cudaMemsetAsync(a.hit_count, 0, sizeof(int32_t) * R, s));
        }
    }
    const bool gen = p.a.gen_pos != nullptr || p.a.cam_W > 0;   // generated inside the forward kernel: the per-warp ray queue
    // kernel variant: feature channels 0 / <= 4 / <= 12 (shared-memory scratch and registers are sized by it), statistics, ray queue
#define IRGS_FWD_CASE(NF_, ST_, GEN_) \
    if (nf == NF_ && stats == ST_ && gen == GEN_) return launch_fwd(h, slot, trace_forward_kernel<NF_, ST_, GEN_>, p, a.n_rays, s);
    const int nf = a.S == 0 ? 0 : (a.S <= 4 ? 4 : NFMAX);
    IRGS_FWD_CASE(0, false, false) IRGS_FWD_CASE(0, false, true) IRGS_FWD_CASE(0, true, false) IRGS_FWD_CASE(0, true, true)
    IRGS_FWD_CASE(4, false, false) IRGS_FWD_CASE(4, false, true) IRGS_FWD_CASE(4, true, false) IRGS_FWD_CASE(4, true, true)
    IRGS_FWD_CASE(NFMAX, false, false) IRGS_FWD_CASE(NFMAX, false, true) IRGS_FWD_CASE(NFMAX, true, false) IRGS_FWD_CASE(NFMAX, true, true)
#undef IRGS_FWD_CASE
    set_error("no forward kernel variant");
    return 1;
}

}  // namespace irgs
