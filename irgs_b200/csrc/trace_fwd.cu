// Forward tracing kernel (sm_100a): persistent, dynamic ray fetch, warp-cooperative sort + compositing.
//
// Replaces gaussiantrace_forward.cu:12-141 of the reference (raygen with 16-hit chunks + any-hit sorted insertion).
//
// Every lane owns one ray and runs a small state machine
//     FETCH -> TRAV -> COMP -> (TRAV for another pass | FETCH)          TRAV -> FULL -> TRAV
//   TRAV  near-first stack walk of the LBVH, one node or one leaf per iteration.  A surfel that passes the plane /
//         alpha test is APPENDED (O(1), unsorted) to the lane's candidate buffer in shared memory.
//   FULL  the buffer holds KB candidates: the warp sorts it co-operatively and trims it at the entry where the
//         buffered hits alone already push the transmittance below T_min (nothing behind it can ever be composited);
//         the walk then continues with the range clipped to that depth.
//   COMP  the pass's walk is finished: the warp sorts the buffer co-operatively (rank sort through shuffles),
//         the transmittance chain is evaluated in the sequential order of the reference, every lane shades ONE hit
//         (SH colour: twelve 16-byte loads, all hits in flight at once) and the weighted sums are warp-reduced.
//         The ordered surfel ids are written with one coalesced store (saved hit list for the backward replay).
// A lane that finishes its ray pulls the next one from a global counter at once; the warp leaves the walk to serve
// FULL / COMP lanes as soon as fewer than MIN_ACTIVE lanes are still walking, which bounds the divergence of the hot
// loop.  (First version: one thread per ray kept a sorted k-buffer by insertion and composited alone; ncu showed 4 of
// 32 lanes active on average -- profiles/r01_ncu_forward_v1.md.)
#include "trace_common.cuh"

namespace irgs {

constexpr int KB = 32;             // candidate buffer depth per ray
constexpr int KROW = KB + 1;       // padded row: co-operative reads of one row are bank-conflict free
constexpr int MIN_ACTIVE = 20;
enum { PH_FETCH = 0, PH_TRAV = 1, PH_COMP = 2, PH_FULL = 3 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool FEAT, bool STATS>
__global__ void __launch_bounds__(TB, 4) trace_forward_kernel(const KParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_t = reinterpret_cast<float *>(smem_raw);          // [TB][KROW]
    int *s_g = reinterpret_cast<int *>(s_t + TB * KROW);        // [TB][KROW]
    float *s_a = reinterpret_cast<float *>(s_g + TB * KROW);    // [TB][KROW]
    const int tid = threadIdx.x;
    const int warp_row0 = (tid & ~31) * KROW;                   // first row of this warp
    float *bt = s_t + tid * KROW; int *bg = s_g + tid * KROW; float *ba = s_a + tid * KROW;
    int stack_n[STACK]; float stack_t[STACK];
    const unsigned lane = tid & 31, lt_mask = (1u << lane) - 1u;
    const unsigned FULL = 0xffffffffu;
    const TraceArgs &a = p.a;
    const float alpha_min = a.alpha_min, T_min = a.T_min;
    const int back_culling = a.back_culling;
    unsigned long long st_nodes = 0, st_leaf = 0, st_hits = 0, st_pass = 0;

    int phase = PH_FETCH;
    bool pool_empty = false;  // warp-uniform
    int64_t ray = 0;
    RayCtx r;
    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, N0 = 0.f, N1 = 0.f, N2 = 0.f, D = 0.f, O = 0.f;
    float F[FEAT ? NFMAX : 1];
    float t_last = -INFINITY, t_lo = 0.f, t_hi = IRGS_T_SCENE_MAX;
    int g_last = -1, g_hi = INT_MAX, total = 0, cnt = 0, sp = 0, cur = 0;
    bool saturated = false;   // the buffer is full and sorted; t_hi is the capacity bound (another pass may follow)
    r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = 0.f; r.idx = r.idy = r.idz = r.oodx = r.oody = r.oodz = 0.f;
#pragma unroll
    for (int j = 0; j < (FEAT ? NFMAX : 1); ++j) F[j] = 0.f;

    for (;;) {
        // ------------------------------------------------------------------ refill idle lanes
        const unsigned need = __ballot_sync(FULL, phase == PH_FETCH);
        if (need != 0u && !pool_empty) {
            const int leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(p.counter, (unsigned long long)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (phase == PH_FETCH) {
                ray = (int64_t)base + __popc(need & lt_mask);
                if (ray < a.n_rays) {
                    load_ray(a, ray, r);
                    ray_setup(r);
                    T = 1.f; C0 = C1 = C2 = N0 = N1 = N2 = D = O = 0.f;
#pragma unroll
                    for (int j = 0; j < (FEAT ? NFMAX : 1); ++j) F[j] = 0.f;
                    t_last = -INFINITY; g_last = -1; total = 0;
                    cnt = 0; sp = 0; cur = 0; t_lo = 0.f; t_hi = IRGS_T_SCENE_MAX; g_hi = INT_MAX; saturated = false;
                    phase = PH_TRAV;
                    if (STATS) ++st_pass;
                }
            }
            if (base + (unsigned long long)__popc(need) >= (unsigned long long)a.n_rays) pool_empty = true;
        }
        unsigned trav = __ballot_sync(FULL, phase == PH_TRAV);
        if (trav == 0u && __ballot_sync(FULL, phase >= PH_COMP) == 0u) break;  // pool empty and every lane idle

        // ------------------------------------------------------------------ BVH walk
        const int thr = pool_empty ? 1 : MIN_ACTIVE;
        while (__popc(trav) >= thr) {
            if (phase == PH_TRAV) {
                bool pop = true;
                if (cur >= 0) {
                    const Node *nd = p.nodes + cur;
                    const float4 qa = __ldg(&nd->a), qb = __ldg(&nd->b), qc = __ldg(&nd->c);
                    const int4 qd = __ldg(&nd->d);
                    if (STATS) ++st_nodes;
                    float tnL, tnR;
                    const bool hL = slab(r, qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, t_lo, t_hi, tnL);
                    const bool hR = slab(r, qb.z, qb.w, qc.x, qc.y, qc.z, qc.w, t_lo, t_hi, tnR);
                    if (hL && hR) {
                        const bool rightNear = tnR < tnL;
                        if (sp < STACK) { stack_n[sp] = rightNear ? qd.x : qd.y; stack_t[sp] = rightNear ? tnL : tnR; ++sp; }
                        cur = rightNear ? qd.y : qd.x;
                        pop = false;
                    } else if (hL) { cur = qd.x; pop = false; }
                    else if (hR) { cur = qd.y; pop = false; }
                } else {
                    if (STATS) ++st_leaf;
                    float t, alpha; int g;
                    if (leaf_test(r, p.recs + (~cur), alpha_min, back_culling, t, g, alpha) &&
                        key_less(t_last, g_last, t, g) && key_less(t, g, t_hi, g_hi)) {
                        if (!saturated) {
                            bt[cnt] = t; bg[cnt] = g; ba[cnt] = alpha;
                            if (++cnt == KB) phase = PH_FULL;
                        } else {
                            // rare: more than KB candidates and no termination among the nearest KB -- keep the KB
                            // nearest by sorted insertion (the buffer is sorted in this mode), dropping the farthest
                            int i = KB - 1;
                            while (i > 0 && key_less(t, g, bt[i - 1], bg[i - 1])) {
                                bt[i] = bt[i - 1]; bg[i] = bg[i - 1]; ba[i] = ba[i - 1];
                                --i;
                            }
                            bt[i] = t; bg[i] = g; ba[i] = alpha;
                            t_hi = bt[KB - 1]; g_hi = bg[KB - 1];
                        }
                    }
                }
                if (pop && phase == PH_TRAV) {
                    for (;;) {
                        if (sp == 0) { phase = PH_COMP; break; }
                        --sp;
                        if (stack_t[sp] <= t_hi) { cur = stack_n[sp]; break; }
                    }
                }
            }
            trav = __ballot_sync(FULL, phase == PH_TRAV);
            if (trav == 0u) break;
        }

        // ------------------------------------------------------------------ lanes whose pass found nothing
        if (phase == PH_COMP && cnt == 0) {
            a.color[3 * ray] = C0; a.color[3 * ray + 1] = C1; a.color[3 * ray + 2] = C2;
            a.normal[3 * ray] = N0; a.normal[3 * ray + 1] = N1; a.normal[3 * ray + 2] = N2;
            a.depth[ray] = D; a.alpha[ray] = O;
            if (FEAT) {
#pragma unroll
                for (int j = 0; j < NFMAX; ++j)
                    if (j < a.S) a.feature[ray * a.S + j] = F[j];
            }
            if (a.hit_count != nullptr) a.hit_count[ray] = total;
            if (STATS) st_hits += total;
            phase = PH_FETCH;
        }

        // ------------------------------------------------------------------ co-operative sort / trim / composite
        unsigned work = __ballot_sync(FULL, phase >= PH_COMP);
        while (work != 0u) {
            const int L = __ffs(work) - 1;
            work &= work - 1u;
            const int n = __shfl_sync(FULL, cnt, L);
            const bool is_full = __shfl_sync(FULL, phase, L) == PH_FULL;
            const int row = warp_row0 + L * KROW;
            // rank sort of the n candidates by (t, surfel id): entry `lane` counts how many precede it
            float my_t = INFINITY, my_a = 0.f; int my_g = INT_MAX;
            if ((int)lane < n) { my_t = s_t[row + lane]; my_g = s_g[row + lane]; my_a = s_a[row + lane]; }
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const float tj = __shfl_sync(FULL, my_t, j);
                const int gj = __shfl_sync(FULL, my_g, j);
                rank += key_less(tj, gj, my_t, my_g) ? 1 : 0;
            }
            __syncwarp();
            if ((int)lane < n) { s_t[row + rank] = my_t; s_g[row + rank] = my_g; s_a[row + rank] = my_a; }
            __syncwarp();
            if ((int)lane < n) { my_t = s_t[row + lane]; my_g = s_g[row + lane]; my_a = s_a[row + lane]; }
            // transmittance chain in the reference's sequential order (bit-identical termination decisions)
            float Tc = __shfl_sync(FULL, T, L);
            float my_w = 0.f;
            int n_comp = n;
            bool term = false;
            for (int i = 0; i < n; ++i) {
                const float ai = __shfl_sync(FULL, my_a, i);
                if ((int)lane == i) my_w = Tc * ai;
                Tc *= (1.f - ai);
                if (Tc < T_min) { n_comp = i + 1; term = true; break; }
            }
            if (is_full) {
                // trim: nothing behind the terminating entry can be composited; otherwise the capacity bound applies
                const float th = __shfl_sync(FULL, my_t, n_comp - 1);
                const int gh = __shfl_sync(FULL, my_g, n_comp - 1);
                if ((int)lane == L) {
                    cnt = n_comp; t_hi = th; g_hi = gh; saturated = (n_comp == KB);  // full and sorted: insertion mode
                    phase = PH_TRAV;
                    // resume the walk: `cur` is the leaf that filled the buffer, pop the next reachable node
                    for (;;) {
                        if (sp == 0) { phase = PH_COMP; break; }
                        --sp;
                        if (stack_t[sp] <= t_hi) { cur = stack_n[sp]; break; }
                    }
                }
                if (__shfl_sync(FULL, phase, L) == PH_COMP) work |= (1u << L);  // stack ran empty: composite right away
                continue;
            }
            // shade one hit per lane
            const float dx = __shfl_sync(FULL, r.dx, L), dy = __shfl_sync(FULL, r.dy, L), dz = __shfl_sync(FULL, r.dz, L);
            float c0 = 0.f, c1 = 0.f, c2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, dd = 0.f, oo = 0.f;
            float f[FEAT ? NFMAX : 1];
#pragma unroll
            for (int j = 0; j < (FEAT ? NFMAX : 1); ++j) f[j] = 0.f;
            if ((int)lane < n_comp) {
                float Y[16];
                sh_basis(a.deg, dx, dy, dz, Y);
                const float nx = __ldg(a.normals + 3 * (size_t)my_g), ny = __ldg(a.normals + 3 * (size_t)my_g + 1),
                            nz = __ldg(a.normals + 3 * (size_t)my_g + 2);
                const float dg = dot3_rn(nx, ny, nz, dx, dy, dz);
                const float m = (-dg > 0.f) ? 1.f : -1.f;
                float c[3];
                sh_color(a.shs, a.K, a.deg, my_g, Y, c);
                c0 = my_w * c[0]; c1 = my_w * c[1]; c2 = my_w * c[2];
                n0 = my_w * m * nx; n1 = my_w * m * ny; n2 = my_w * m * nz;
                dd = my_w * my_t; oo = my_w;
                if (FEAT) {
#pragma unroll
                    for (int j = 0; j < NFMAX; ++j)
                        if (j < a.S) f[j] = my_w * __ldg(a.features + (size_t)my_g * a.S + j);
                }
            }
            const int64_t ray_L = __shfl_sync(FULL, ray, L);
            const int total_L = __shfl_sync(FULL, total, L);
            if (a.hits != nullptr && (int)lane < n_comp && total_L + (int)lane < a.hit_cap)
                a.hits[ray_L * a.hit_cap + total_L + lane] = my_g;
            c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
            n0 = warp_sum(n0); n1 = warp_sum(n1); n2 = warp_sum(n2);
            dd = warp_sum(dd); oo = warp_sum(oo);
            if (FEAT) {
#pragma unroll
                for (int j = 0; j < NFMAX; ++j)
                    if (j < a.S) f[j] = warp_sum(f[j]);
            }
            const float t_end = __shfl_sync(FULL, my_t, n - 1);
            const int g_end = __shfl_sync(FULL, my_g, n - 1);
            if ((int)lane == L) {
                C0 += c0; C1 += c1; C2 += c2; N0 += n0; N1 += n1; N2 += n2; D += dd; O += oo;
                if (FEAT) {
#pragma unroll
                    for (int j = 0; j < NFMAX; ++j) F[j] += f[j];
                }
                T = Tc;
                total += n_comp;
                if (!term && cnt == KB) {
                    // the nearest KB candidates were all composited without terminating: another pass, strictly
                    // after the last one
                    t_last = t_end; g_last = g_end;
                    cnt = 0; sp = 0; cur = 0; t_lo = fmaxf(t_last, 0.f); t_hi = IRGS_T_SCENE_MAX; g_hi = INT_MAX;
                    saturated = false;
                    phase = PH_TRAV;
                    if (STATS) ++st_pass;
                } else {
                    a.color[3 * ray] = C0; a.color[3 * ray + 1] = C1; a.color[3 * ray + 2] = C2;
                    a.normal[3 * ray] = N0; a.normal[3 * ray + 1] = N1; a.normal[3 * ray + 2] = N2;
                    a.depth[ray] = D; a.alpha[ray] = O;
                    if (FEAT) {
#pragma unroll
                        for (int j = 0; j < NFMAX; ++j)
                            if (j < a.S) a.feature[ray * a.S + j] = F[j];
                    }
                    if (a.hit_count != nullptr) a.hit_count[ray] = total;
                    if (STATS) st_hits += total;
                    phase = PH_FETCH;
                }
            }
        }
    }
    if (STATS) {
        atomicAdd(p.stats + 0, st_nodes); atomicAdd(p.stats + 1, st_leaf);
        atomicAdd(p.stats + 2, st_hits); atomicAdd(p.stats + 3, st_pass);
    }
}

template <typename Kern>
static int launch_fwd(irgs_tracer *h, Kern kern, const KParams &p, int64_t n_rays, cudaStream_t s) {
    const size_t smem = (size_t)TB * KROW * 12;
    IRGS_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IRGS_CHECK(cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), s));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TB, smem) != cudaSuccess || per_sm < 1) per_sm = 2;
    int grid = h->sm_count * per_sm;
    const int64_t need = (n_rays + TB - 1) / TB;
    if (need < grid) grid = (int)(need > 0 ? need : 1);
    kern<<<grid, TB, smem, s>>>(p);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int launch_trace_forward(irgs_tracer *h, const TraceArgs &a, cudaStream_t s) {
    KParams p;
    p.a = a; p.nodes = h->nodes; p.recs = h->recs; p.counter = h->counter; p.stats = h->stats;
    const bool feat = a.S > 0, stats = h->stats_enabled != 0;
    if (stats) IRGS_CHECK(cudaMemsetAsync(h->stats, 0, 4 * sizeof(unsigned long long), s));
    if (feat) return stats ? launch_fwd(h, trace_forward_kernel<true, true>, p, a.n_rays, s)
                           : launch_fwd(h, trace_forward_kernel<true, false>, p, a.n_rays, s);
    return stats ? launch_fwd(h, trace_forward_kernel<false, true>, p, a.n_rays, s)
                 : launch_fwd(h, trace_forward_kernel<false, false>, p, a.n_rays, s);
}

}  // namespace irgs
