// Backward, re-trace, record packing, intersection test and incident-ray kernels of the surfel tracer (sm_100a, compute-only
// BVH traversal: B200 has no RT cores).  The forward kernel is in trace_fwd.cu.
//
// Replaces, of /root/reference/submodules/surfel_tracer/src/optix/:
//   gaussiantrace_backward.cu:11-200  raygen (re-trace, analytic gradients, 74 scalar atomics per hit)
//   gaussiantrace_intersection_test.cu:12-35
// and auxiliary.h:91-143 (backward of the SH colour).
//
// Kernels (DESIGN.md section 3 has the long form):
//   * pack_records_kernel          the caller's five per-surfel arrays -> 64-byte records in leaf order + inverse order
//   * trace_backward_flat_kernel   replay of the hit lists the forward saved, ONE HIT PER LANE: segmented warp scans for the
//                                  sequential quantities, one TMA bulk reduction (256 B) per hit into the fused [N,64] buffer
//   * trace_backward_replay_kernel the same replay, one ray per thread (kept for comparison, bwd_mode 1)
//   * trace_backward_retrace_kernel the reference's scheme -- re-trace in ordered 16-hit passes (sorted k-buffer in shared
//                                  memory, binary quantised nodes) -- for rays without a (complete) saved list
//   * intersection_test_kernel, incident_rays_kernel, incident_backward_kernel, unpack_grads_kernel
// The depth of a hit is computed with the same explicit sequence of IEEE operations as oracle/surfel_oracle.c, so the hit
// order is bit-identical to the oracle's.
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "trace_common.cuh"

namespace irgs {

// ------------------------------------------------------------------------------------------------ pack
// Gather the caller's five per-surfel arrays into 64-byte records in leaf (Morton) order, once per trace call,
// so that one leaf test is one (early reject) or two 32-byte loads from one line; also the inverse map surfel -> leaf
// position, through which the backward replay reaches the same records.
// r0.w is the squared radius of the surfel's support { alpha >= alpha_min } around mu: the in-plane map p -> (ru.p, rv.p)
// has smallest singular value sqrt(lambda_min), so |p|^2 <= 2 ln(opacity / alpha_min) / lambda_min on the support.
__global__ void pack_records_kernel(const int *__restrict__ order, int n, const float *__restrict__ means,
                                    const float *__restrict__ opacity, const float *__restrict__ ru,
                                    const float *__restrict__ rv, const float *__restrict__ normals, float alpha_min,
                                    SurfelRec *__restrict__ recs, int *__restrict__ inv_order) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int g = order[i];
    const float *m = means + 3 * (size_t)g, *nn = normals + 3 * (size_t)g, *a = ru + 3 * (size_t)g, *b = rv + 3 * (size_t)g;
    const float op = opacity[g];
    float rmax2 = -1.f;   // opacity <= alpha_min: never a hit
    if (op > alpha_min) {
        rmax2 = INFINITY;  // degenerate frames: no early reject
        const float n2 = nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2];
        if (n2 > 0.f) {
            const float an = (a[0] * nn[0] + a[1] * nn[1] + a[2] * nn[2]), bn = (b[0] * nn[0] + b[1] * nn[1] + b[2] * nn[2]);
            const float aa = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] - an * an / n2;
            const float dd = b[0] * b[0] + b[1] * b[1] + b[2] * b[2] - bn * bn / n2;
            const float ab = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] - an * bn / n2;
            const float h = 0.5f * (aa - dd);
            const float lmin = 0.5f * (aa + dd) - sqrtf(h * h + ab * ab);
            const float r2 = 2.0f * logf(op / alpha_min);
            // the subtraction loses relative accuracy when the frame is very anisotropic: trust it only when well conditioned
            if (lmin > 1e-3f * (aa + dd)) rmax2 = r2 / lmin * 1.001f + 1e-12f;
        }
    }
    SurfelRec r;
    r.r0 = make_float4(m[0], m[1], m[2], rmax2);
    r.r1 = make_float4(nn[0], nn[1], nn[2], __int_as_float(g));
    r.r2 = make_float4(a[0], a[1], a[2], b[0]);
    r.r3 = make_float4(b[1], b[2], op, 0.f);
    recs[i] = r;
    inv_order[g] = i;
}

// One pass: collect the <= KBUF nearest candidates strictly after (t_last, g_last), ascending, into the shared
// k-buffer columns of this thread.  Returns the number collected.
template <bool STATS>
__device__ __forceinline__ int collect_pass(const KParams &p, const RayCtx &r, float t_last, int g_last, float *bt, int *bg,
                                            float *ba, int *stack_n, float *stack_t, unsigned &n_nodes, unsigned &n_leaf) {
    int cnt = 0;
    float t_hi = IRGS_T_SCENE_MAX;
    const float t_lo = fmaxf(t_last, 0.0f);
    int sp = 0;
    int cur = 0;  // root
    const float alpha_min = p.a.alpha_min;
    const int back_culling = p.a.back_culling;
    for (;;) {
        if (cur >= 0) {
            uint4 wl, wr;
                ldg256(&p.nodes[cur], wl, wr);
            const int2 d = make_int2((int)wl.w, (int)wr.w);
            if (STATS) ++n_nodes;
            float tnL, tnR;
            bool hL = slab(r, wl, t_lo, t_hi, tnL);
            bool hR = slab(r, wr, t_lo, t_hi, tnR);
            if (hL && hR) {
                bool rightNear = tnR < tnL;
                int nearC = rightNear ? d.y : d.x, farC = rightNear ? d.x : d.y;
                float farT = rightNear ? tnL : tnR;
                if (sp < STACK) { stack_n[sp] = farC; stack_t[sp] = farT; ++sp; }
                cur = nearC;
                continue;
            } else if (hL) { cur = d.x; continue; }
            else if (hR) { cur = d.y; continue; }
        } else {
            if (STATS) ++n_leaf;
            float t, alpha; int g;
            if (leaf_test(r, p.recs + (~cur), alpha_min, back_culling, t, g, alpha)) {
                bool after = key_less(t_last, g_last, t, g);
                bool fits = cnt < KBUF || key_less(t, g, bt[(KBUF - 1) * TB], bg[(KBUF - 1) * TB]);
                if (after && fits) {
                    int i = cnt < KBUF ? cnt++ : KBUF - 1;
                    while (i > 0 && key_less(t, g, bt[(i - 1) * TB], bg[(i - 1) * TB])) {
                        bt[i * TB] = bt[(i - 1) * TB]; bg[i * TB] = bg[(i - 1) * TB]; ba[i * TB] = ba[(i - 1) * TB];
                        --i;
                    }
                    bt[i * TB] = t; bg[i * TB] = g; ba[i * TB] = alpha;
                    if (cnt == KBUF) t_hi = bt[(KBUF - 1) * TB];
                }
            }
        }
        // pop, skipping entries that the shrinking range has made unreachable
        for (;;) {
            if (sp == 0) return cnt;
            --sp;
            if (stack_t[sp] <= t_hi) { cur = stack_n[sp]; break; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ backward
template <bool FEAT>
struct BwdState {
    float T, C[3], N[3], D, O;
    float Cf[3], Nf[3], Df, Of;
    float gC[3], gN[3], gD, gO;
    float go[3], gd[3];
    float F[FEAT ? NFMAX : 1], Ff[FEAT ? NFMAX : 1], gF[FEAT ? NFMAX : 1];
};

template <bool FEAT>
__device__ __forceinline__ void bwd_load(const TraceArgs &a, int64_t ray, BwdState<FEAT> &s) {
    const int64_t gr = a.gout_period > 0 ? (a.gout_offset + ray) % a.gout_period : ray;
    s.T = 1.f; s.D = 0.f; s.O = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        s.C[j] = 0.f; s.N[j] = 0.f; s.go[j] = 0.f; s.gd[j] = 0.f;
        s.Cf[j] = a.color[3 * ray + j]; s.Nf[j] = a.normal[3 * ray + j];
        s.gC[j] = __ldg(a.gC + 3 * gr + j); s.gN[j] = __ldg(a.gN + 3 * gr + j);
    }
    s.Df = a.depth[ray]; s.Of = a.alpha[ray];
    s.gD = __ldg(a.gD + gr); s.gO = __ldg(a.gO + gr);
    if (FEAT) {
#pragma unroll
        for (int j = 0; j < NFMAX; ++j) {
            s.F[j] = 0.f;
            s.Ff[j] = j < a.S ? a.feature[ray * a.S + j] : 0.f;
            s.gF[j] = j < a.S ? __ldg(a.gF + gr * a.S + j) : 0.f;
        }
    }
}

// One hit of the backward replay: gaussiantrace_backward.cu:61-166 verbatim in its formulas (no derivative masks
// for min(0.99,.) / max(.,0), no SH-direction term, division by the raw d_g), with the 13+48 scalar atomics of the
// reference replaced by 16-byte vector reductions into the fused per-surfel gradient row.
template <bool FEAT>
__device__ __forceinline__ void bwd_hit(const TraceArgs &a, const RayCtx &r, const float Y[16], int g, BwdState<FEAT> &s) {
    const float *pm = a.means + 3 * (size_t)g, *pn = a.normals + 3 * (size_t)g, *pa = a.ru + 3 * (size_t)g,
                *pb = a.rv + 3 * (size_t)g;
    const float mx = __ldg(pm), my = __ldg(pm + 1), mz = __ldg(pm + 2);
    const float nx = __ldg(pn), ny = __ldg(pn + 1), nz = __ldg(pn + 2);
    const float ax = __ldg(pa), ay = __ldg(pa + 1), az = __ldg(pa + 2);
    const float bx = __ldg(pb), by = __ldg(pb + 1), bz = __ldg(pb + 2);
    const float op = __ldg(a.opacity + g);
    const float relx = __fsub_rn(r.ox, mx), rely = __fsub_rn(r.oy, my), relz = __fsub_rn(r.oz, mz);
    const float og = dot3_rn(nx, ny, nz, relx, rely, relz);
    const float dg = dot3_rn(nx, ny, nz, r.dx, r.dy, r.dz);
    const float den = fmaxf(1e-6f, __fmul_rn(dg, dg));
    const float t = __fdiv_rn(__fmul_rn(-og, dg), den);
    const float m = (-dg > 0.f) ? 1.f : -1.f;
    const float px = __fmaf_rn(t, r.dx, relx), py = __fmaf_rn(t, r.dy, rely), pz = __fmaf_rn(t, r.dz, relz);
    const float pu = dot3_rn(ax, ay, az, px, py, pz), pv = dot3_rn(bx, by, bz, px, py, pz);
    const float G = __expf(__fmul_rn(-0.5f, __fadd_rn(__fmul_rn(pu, pu), __fmul_rn(pv, pv))));
    const float alpha = fminf(0.99f, __fmul_rn(op, G));
    float c[3];
    sh_color(a.shs, a.K, a.deg, g, Y, c);
    const float w = s.T * alpha;
    const float nf[3] = {m * nx, m * ny, m * nz};
#pragma unroll
    for (int j = 0; j < 3; ++j) { s.C[j] += w * c[j]; s.N[j] += w * nf[j]; }
    s.D += w * t; s.O += w;
    float feat[FEAT ? NFMAX : 1];
    if (FEAT) {
#pragma unroll
        for (int j = 0; j < NFMAX; ++j) {
            feat[j] = j < a.S ? __ldg(a.features + (size_t)g * a.S + j) : 0.f;
            s.F[j] += w * feat[j];
        }
    }
    s.T *= (1.f - alpha);
    const float T = s.T;
    float dL_dalpha = s.gD * (T * t - (s.Df - s.D)) + s.gO * (1.f - s.Of);
#pragma unroll
    for (int j = 0; j < 3; ++j)
        dL_dalpha += s.gC[j] * (T * c[j] - (s.Cf[j] - s.C[j])) + s.gN[j] * (T * nf[j] - (s.Nf[j] - s.N[j]));
    if (FEAT) {
#pragma unroll
        for (int j = 0; j < NFMAX; ++j) dL_dalpha += s.gF[j] * (T * feat[j] - (s.Ff[j] - s.F[j]));
    }
    dL_dalpha /= (1.f - alpha);
    const float dL_do = dL_dalpha * G;
    const float dL_dG = dL_dalpha * op;
    const float dpu = -dL_dG * G * pu, dpv = -dL_dG * G * pv;
    const float dposx = dpu * ax + dpv * bx, dposy = dpu * ay + dpv * by, dposz = dpu * az + dpv * bz;
    const float dL_dd = s.gD * w + (dposx * r.dx + dposy * r.dy + dposz * r.dz);
    const float dL_dog = -dL_dd / dg;
    const float dL_ddg = dL_dd * og / fmaxf(1e-6f, dg * dg);
    s.go[0] += dposx + dL_dog * nx; s.go[1] += dposy + dL_dog * ny; s.go[2] += dposz + dL_dog * nz;
    s.gd[0] += t * dposx + dL_ddg * nx; s.gd[1] += t * dposy + dL_ddg * ny; s.gd[2] += t * dposz + dL_ddg * nz;
    const float dnx = m * s.gN[0] * w + dL_ddg * r.dx + dL_dog * relx;
    const float dny = m * s.gN[1] * w + dL_ddg * r.dy + dL_dog * rely;
    const float dnz = m * s.gN[2] * w + dL_ddg * r.dz + dL_dog * relz;
    float4 *row = reinterpret_cast<float4 *>(a.grad_fused + (size_t)g * IRGS_GRAD_STRIDE);
    atomicAdd(row + 0, make_float4(-dposx - dL_dog * nx, -dposy - dL_dog * ny, -dposz - dL_dog * nz, dL_do));
    atomicAdd(row + 1, make_float4(dpu * px, dpu * py, dpu * pz, dpv * px));
    atomicAdd(row + 2, make_float4(dpv * py, dpv * pz, dnx, dny));
    atomicAdd(reinterpret_cast<float *>(row + 3), dnz);
    // SH coefficients: Y_k * dL/dc, dL/dc = grad_color * w  (auxiliary.h:91-143)
    const float gc[3] = {s.gC[0] * w, s.gC[1] * w, s.gC[2] * w};
    const int nvec = ((a.deg + 1) * (a.deg + 1) * 3 + 3) >> 2;
#pragma unroll
    for (int v = 0; v < 12; ++v) {
        if (v < nvec) {
            float4 q;
            q.x = Y[(4 * v) / 3] * gc[(4 * v) % 3];
            q.y = Y[(4 * v + 1) / 3] * gc[(4 * v + 1) % 3];
            q.z = Y[(4 * v + 2) / 3] * gc[(4 * v + 2) % 3];
            q.w = Y[(4 * v + 3) / 3] * gc[(4 * v + 3) % 3];
            atomicAdd(row + 4 + v, q);
        }
    }
    if (FEAT) {
#pragma unroll
        for (int j = 0; j < NFMAX; ++j)
            if (j < a.S) atomicAdd(a.grad_features + (size_t)g * a.S + j, s.gF[j] * w);
    }
}

// Replay of the saved hit lists: one thread per ray, no traversal.  (A persistent variant that compacts the ~30 % of
// rays that have hits into a shared-memory queue and refills lanes individually ran at full warps but was not faster
// -- 8.1 vs 7.2 ms per 6.5 M rays: the kernel is bound by the 16 vector reductions per hit, not by lane occupancy.)
template <bool FEAT>
__global__ void __launch_bounds__(TB) trace_backward_replay_kernel(const KParams p) {
    const TraceArgs &a = p.a;
    const int64_t ray = (int64_t)blockIdx.x * TB + threadIdx.x;
    if (ray >= a.n_rays) return;
    const int cnt = a.hit_count[ray];
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    // gaussiantrace_backward.cu:13-14: rays whose forward alpha is exactly zero contribute nothing
    if (cnt > 0 && cnt <= a.hit_cap && a.alpha[ray] != 0.f) {
        RayCtx r;
        load_ray(a, ray, r);
        float Y[16];
        sh_basis(a.deg, r.dx, r.dy, r.dz, Y);
        BwdState<FEAT> s;
        bwd_load<FEAT>(a, ray, s);
        const int32_t *hl = a.hits + ray * a.hit_cap;
        for (int i = 0; i < cnt; ++i) bwd_hit<FEAT>(a, r, Y, __ldg(hl + i), s);
#pragma unroll
        for (int j = 0; j < 3; ++j) { go[j] = s.go[j]; gd[j] = s.gd[j]; }
    }
    if (cnt <= a.hit_cap) {  // rays with longer lists are written by the re-trace kernel
#pragma unroll
        for (int j = 0; j < 3; ++j) { a.g_rays_o[3 * ray + j] = go[j]; a.g_rays_d[3 * ray + j] = gd[j]; }
    }
}

// Hit-parallel replay: ONE HIT PER LANE.  A warp owns 32 consecutive rays; their saved hit lists are flattened (prefix sum
// of the hit counts) and processed 32 hits at a time whichever ray they belong to:
//   * the hit's ray (origin, direction, final outputs, output gradients) comes from the owning lane through shuffles;
//   * the sequential quantities of gaussiantrace_backward.cu:61-98 become segmented warp scans, a segment being the run
//     of lanes that hold hits of one ray: T is a segmented prefix product of (1 - alpha), and the "what is still to
//     come" terms (C_final - C_i, ...) are segmented exclusive SUFFIX sums of the w * c contributions, which have no
//     cancellation (the reference subtracts two nearly equal running sums);  a ray whose list straddles two rounds
//     carries T and the remainders over in registers (lane 31 of one round -> the first segment of the next);
//   * every lane writes the 64-float gradient row of its hit to shared memory; each row is then added to the fused buffer by
//     one TMA bulk reduction (BULK; 2 cache lines per hit instead of 16 scattered requests), or -- comparison variant -- with
//     one 16-byte reduction per lane, sixteen consecutive lanes covering one 256-byte row.
// ncu on the thread-per-ray replay (profiles/r01_bwd_replay_regions.txt): 7.96 of 32 lanes active, L1TEX 85 % busy
// with the scattered reductions.
// Launch shape (round 2, same-box A/B in profiles/r02_sweeps.txt): ONE warp per block.  The warps of a block are independent
// 32-ray groups of very uneven length (a group's hits range from 0 to > 1000), and a block keeps its registers and shared memory
// until its slowest warp is done: with four warps per block ncu showed 11.6 warps active per SM of 16 possible.  96 registers
// (no spills) and 11.4 KB of shared memory per warp give 18 resident warps per SM.
#ifndef IRGS_BWD_TB
#define IRGS_BWD_TB 32   // threads per block of the hit-parallel replay (a block's warps are independent 32-ray groups)
#endif
#ifndef IRGS_BWD_BLOCKS
#define IRGS_BWD_BLOCKS 20
#endif
constexpr int FTB = IRGS_BWD_TB;
#ifndef IRGS_BWD_UNIFORM_ISSUE
#define IRGS_BWD_UNIFORM_ISSUE 1   // 0: per-lane bulk reductions under `if (act)` (comparison builds)
#endif
// one lane of the (converged) warp, chosen by the hardware: the compiler knows that code under it runs in a single thread
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
constexpr int BROW = 68;   // floats per shared-memory row: 64 gradient floats + pad, 68 % 32 == 4 (conflict-free float4 stores)

template <bool FEAT, bool BULK>
__global__ void __launch_bounds__(FTB, IRGS_BWD_BLOCKS) trace_backward_flat_kernel(const KParams p) {
    __shared__ __align__(16) float s_rows[FTB / 32][32 * BROW];
    // per ray of the warp: origin, direction, final outputs (C, N, D), incoming gradients (C, N, D) and gO * (1 - O_final), the only
    // way the alpha output enters: [field][lane]
    __shared__ float s_ray[FTB / 32][21 * 32];
    const TraceArgs &a = p.a;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float *rows = s_rows[threadIdx.x >> 5];
    float *rayv = s_ray[threadIdx.x >> 5];
    // (one 32-ray group per warp, one launch wave after the other: a grid-stride loop over the groups with a resident-sized grid
    // measured 2.44 instead of 1.84 ms per 2^22 rays -- the hardware's dynamic block scheduling balances the very uneven groups)
    const int64_t ray0 = ((int64_t)blockIdx.x * (FTB / 32) + (threadIdx.x >> 5)) * 32;
    if (ray0 >= a.n_rays) return;
    const int64_t my_ray = ray0 + lane;
    const bool valid = my_ray < a.n_rays;
    const int cnt = valid ? a.hit_count[my_ray] : 0;
    // gaussiantrace_backward.cu:13-14: rays whose forward alpha is exactly zero contribute nothing
    const bool has = valid && cnt > 0 && cnt <= a.hit_cap && a.alpha[my_ray] != 0.f;
    RayCtx r;
    r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = 0.f;
    float fin[8], gout[8];   // C(3) N(3) D O of the forward / their incoming gradients
#pragma unroll
    for (int j = 0; j < 8; ++j) { fin[j] = 0.f; gout[j] = 0.f; }
    int64_t gr = 0;
    if (has) {
        load_ray(a, my_ray, r);
        gr = a.gout_period > 0 ? (a.gout_offset + my_ray) % a.gout_period : my_ray;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            fin[j] = a.color[3 * my_ray + j]; fin[3 + j] = a.normal[3 * my_ray + j];
            gout[j] = __ldg(a.gC + 3 * gr + j); gout[3 + j] = __ldg(a.gN + 3 * gr + j);
        }
        fin[6] = a.depth[my_ray]; fin[7] = a.alpha[my_ray];
        gout[6] = __ldg(a.gD + gr); gout[7] = __ldg(a.gO + gr);
    }
    // The hits of a round reach their ray's data through shared memory ([field][owner]: conflict-free for any mix of owners)
    // instead of 22 shuffles out of 22 registers that would stay live through the whole kernel (128 registers, 18 % warps active).
    rayv[0 * 32 + lane] = r.ox; rayv[1 * 32 + lane] = r.oy; rayv[2 * 32 + lane] = r.oz;
    rayv[3 * 32 + lane] = r.dx; rayv[4 * 32 + lane] = r.dy; rayv[5 * 32 + lane] = r.dz;
#pragma unroll
    for (int j = 0; j < 7; ++j) { rayv[(6 + j) * 32 + lane] = fin[j]; rayv[(13 + j) * 32 + lane] = gout[j]; }
    rayv[20 * 32 + lane] = gout[7] * (1.f - fin[7]);
    __syncwarp();
    const int c_eff = has ? cnt : 0;
    int incl = c_eff;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int start = incl - c_eff;
    const int total = __shfl_sync(FULL, incl, 31);
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    float carryT = 1.f, carry_rem[7], carry_remF[FEAT ? NFMAX : 1];
#pragma unroll
    for (int j = 0; j < 7; ++j) carry_rem[j] = 0.f;
#pragma unroll
    for (int j = 0; j < (FEAT ? NFMAX : 1); ++j) carry_remF[j] = 0.f;
    const int nvec = ((a.deg + 1) * (a.deg + 1) * 3 + 3) >> 2;

    // Software pipeline over the 32-hit rounds: the chain hit id -> inverse order -> record is three dependent L2 accesses; the
    // id of a hit is fetched two rounds ahead and its leaf position one round ahead, so that a round only waits for one level
    // (record and SH row, issued together).  locate(): flat index -> (owning lane, index inside that ray's list).
    auto locate = [&](int idx, int &own, int &kk) {
        int o = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int v = __shfl_sync(FULL, incl, o + step - 1);
            if (v <= idx) o += step;
        }
        const bool in = idx < total;
        o = in ? o : lane;
        const int st = __shfl_sync(FULL, start, o);
        own = o; kk = in ? idx - st : 0;
        return in;
    };
    auto fetch_id = [&](int idx) {
        int own, kk;
        const bool in = locate(idx, own, kk);
        return in ? __ldg(a.hits + (ray0 + own) * a.hit_cap + kk) : 0;
    };
#ifndef IRGS_BWD_PIPELINE
#define IRGS_BWD_PIPELINE 1   // 0: every round fetches its own ids and positions (comparison builds)
#endif
#ifndef IRGS_BWD_LOCATE_ONCE
#define IRGS_BWD_LOCATE_ONCE 1   // 0: a round locates its own hits again although the id fetch two rounds earlier already did
#endif
#if IRGS_BWD_LOCATE_ONCE
    // (owner, index) of a flat position is found once, when its id is requested two rounds ahead, and carried to the round that uses it
    int own_c, kk_c, own_n, kk_n;
    int g_cur, g_nxt;
    { const bool in = locate(lane, own_c, kk_c); g_cur = in ? __ldg(a.hits + (ray0 + own_c) * a.hit_cap + kk_c) : 0; }
    { const bool in = locate(32 + lane, own_n, kk_n); g_nxt = in ? __ldg(a.hits + (ray0 + own_n) * a.hit_cap + kk_n) : 0; }
#else
    int g_cur = fetch_id(lane), g_nxt = fetch_id(32 + lane);
#endif
    int pos_cur = lane < total ? __ldg(p.inv_order + g_cur) : 0;
    for (int base = 0; base < total; base += 32) {
        const int idx = base + lane;
        const bool act = idx < total;
#if IRGS_BWD_PIPELINE
        const int pos_nxt = idx + 32 < total ? __ldg(p.inv_order + g_nxt) : 0;   // next round's leaf positions
#if IRGS_BWD_LOCATE_ONCE
        int own_2, kk_2, g_nxt2;                                                    // ids of the round after next
        { const bool in = locate(idx + 64, own_2, kk_2); g_nxt2 = in ? __ldg(a.hits + (ray0 + own_2) * a.hit_cap + kk_2) : 0; }
#else
        const int g_nxt2 = fetch_id(idx + 64);                                     // ids of the round after next
#endif
#else
        const int pos_nxt = 0, g_nxt2 = 0;
        g_cur = fetch_id(idx);
        pos_cur = act ? __ldg(p.inv_order + g_cur) : 0;
#endif
#if IRGS_BWD_LOCATE_ONCE && IRGS_BWD_PIPELINE
        const int owner = own_c, k_loc = kk_c;
#else
        int owner, k_loc;
        locate(idx, owner, k_loc);
#endif
        const int n_o = __shfl_sync(FULL, c_eff, owner);
        const int k = k_loc;
        const int64_t ray = ray0 + owner;
        RayCtx ro;
        ro.ox = rayv[0 * 32 + owner]; ro.oy = rayv[1 * 32 + owner]; ro.oz = rayv[2 * 32 + owner];
        ro.dx = rayv[3 * 32 + owner]; ro.dy = rayv[4 * 32 + owner]; ro.dz = rayv[5 * 32 + owner];
        float F[7], gO[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) { F[j] = rayv[(6 + j) * 32 + owner]; gO[j] = rayv[(13 + j) * 32 + owner]; }
        const float gO_alpha = rayv[20 * 32 + owner];   // grad_alpha * (1 - alpha_final)
        const int64_t o_gr = __shfl_sync(FULL, gr, owner);
        // segment of this lane's ray inside the round
        const int first = act ? lane - k : lane, last = act ? lane + (n_o - 1 - k) : lane;
        const int seg_lo = max(first, 0), seg_hi = min(last, 31);
        const bool straddles_in = first < 0;

        // ---- the hit itself (same arithmetic as bwd_hit / the forward)
        float t = 0.f, alpha = 0.f, G = 0.f, op = 0.f, og = 0.f, dg = 1.f, m = 1.f, pu = 0.f, pv = 0.f;
        float nx = 0.f, ny = 0.f, nz = 0.f, ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
        float relx = 0.f, rely = 0.f, relz = 0.f, px = 0.f, py = 0.f, pz = 0.f;
        float c[3] = {0.f, 0.f, 0.f};
        float feat[FEAT ? NFMAX : 1], Ff[FEAT ? NFMAX : 1], gF[FEAT ? NFMAX : 1];
#pragma unroll
        for (int j = 0; j < (FEAT ? NFMAX : 1); ++j) { feat[j] = 0.f; Ff[j] = 0.f; gF[j] = 0.f; }
        int g = 0;
        if (act) {
            g = g_cur;
            // the hit's colour: from the forward's colour cache (12 B, three streamed loads next to the neighbouring hits' of the ray)
            // or -- hits beyond the cached entries, or no valid cache -- from the surfel's SH row (192 B gather + 48 fma)
            const bool cached = a.hit_rgb != nullptr && k < a.rgb_cap;
            if (cached) {
                const float *hc = a.hit_rgb + ((size_t)ray * a.rgb_cap + k) * 3;
                c[0] = __ldcs(hc); c[1] = __ldcs(hc + 1); c[2] = __ldcs(hc + 2);
            }
            // the surfel's packed record (re-packed from the saved inputs before this launch): two 32-byte loads
            float4 q0, q1, q2, q3;
            const SurfelRec *rec = p.recs + pos_cur;
            ldg256(&rec->r0, q0, q1);
            ldg256(&rec->r2, q2, q3);
            const float mx = q0.x, my = q0.y, mz = q0.z;
            nx = q1.x; ny = q1.y; nz = q1.z;
            ax = q2.x; ay = q2.y; az = q2.z;
            bx = q2.w; by = q3.x; bz = q3.y;
            op = q3.z;
            relx = __fsub_rn(ro.ox, mx); rely = __fsub_rn(ro.oy, my); relz = __fsub_rn(ro.oz, mz);
            og = dot3_rn(nx, ny, nz, relx, rely, relz);
            dg = dot3_rn(nx, ny, nz, ro.dx, ro.dy, ro.dz);
            const float den = fmaxf(1e-6f, __fmul_rn(dg, dg));
            t = __fdiv_rn(__fmul_rn(-og, dg), den);
            m = (-dg > 0.f) ? 1.f : -1.f;
            px = __fmaf_rn(t, ro.dx, relx); py = __fmaf_rn(t, ro.dy, rely); pz = __fmaf_rn(t, ro.dz, relz);
            pu = dot3_rn(ax, ay, az, px, py, pz); pv = dot3_rn(bx, by, bz, px, py, pz);
            G = __expf(__fmul_rn(-0.5f, __fadd_rn(__fmul_rn(pu, pu), __fmul_rn(pv, pv))));
            alpha = fminf(0.99f, __fmul_rn(op, G));
            if (!cached) {
                float Y[16];   // (only for the colour here: the SH gradient row re-derives the basis when it is written, so that
                               //  sixteen registers are not held across the segmented scans in between)
                sh_basis(a.deg, ro.dx, ro.dy, ro.dz, Y);
                sh_color(a.shs, a.K, a.deg, g, Y, c);
            }
            if (FEAT) {
#pragma unroll
                for (int j = 0; j < NFMAX; ++j)
                    if (j < a.S) {
                        feat[j] = __ldg(a.features + (size_t)g * a.S + j);
                        Ff[j] = a.feature[ray * a.S + j];
                        gF[j] = __ldg(a.gF + o_gr * a.S + j);
                    }
            }
        }
        const float nf[3] = {m * nx, m * ny, m * nz};

        // ---- transmittance: segmented inclusive prefix product of (1 - alpha)
        float incT = act ? (1.f - alpha) : 1.f;
        bool up_ok[5], dn_ok[5];
#pragma unroll
        for (int s5 = 0; s5 < 5; ++s5) { up_ok[s5] = lane - (1 << s5) >= seg_lo; dn_ok[s5] = lane + (1 << s5) <= seg_hi; }
#pragma unroll
        for (int s5 = 0; s5 < 5; ++s5) {
            const float v = __shfl_up_sync(FULL, incT, 1 << s5);
            if (up_ok[s5]) incT *= v;
        }
        float exclT = __shfl_up_sync(FULL, incT, 1);
        if (lane <= seg_lo) exclT = 1.f;
        const float T0 = straddles_in ? carryT : 1.f;
        const float T_before = T0 * exclT, T = T0 * incT;   // T: after this hit, like s.T in bwd_hit
        const float w = act ? T_before * alpha : 0.f;

        // ---- "still to come" terms: segmented exclusive suffix sums of the compositing contributions
        float x[7] = {w * c[0], w * c[1], w * c[2], w * nf[0], w * nf[1], w * nf[2], w * t};
        float R[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            float e = __shfl_down_sync(FULL, x[j], 1);
            if (!dn_ok[0]) e = 0.f;
#pragma unroll
            for (int s5 = 0; s5 < 5; ++s5) {
                const float v = __shfl_down_sync(FULL, e, 1 << s5);
                if (dn_ok[s5]) e += v;
            }
            const float seg_total = __shfl_sync(FULL, e + x[j], seg_lo);
            const float rem_after = (straddles_in ? carry_rem[j] : F[j]) - seg_total;
            R[j] = rem_after + e;
            carry_rem[j] = __shfl_sync(FULL, rem_after, 31);
        }
        float RF[FEAT ? NFMAX : 1];
        if (FEAT) {
#pragma unroll
            for (int j = 0; j < NFMAX; ++j) {
                RF[j] = 0.f;
                if (j < a.S) {
                    const float xf = w * feat[j];
                    float e = __shfl_down_sync(FULL, xf, 1);
                    if (!dn_ok[0]) e = 0.f;
#pragma unroll
                    for (int s5 = 0; s5 < 5; ++s5) {
                        const float v = __shfl_down_sync(FULL, e, 1 << s5);
                        if (dn_ok[s5]) e += v;
                    }
                    const float seg_total = __shfl_sync(FULL, e + xf, seg_lo);
                    const float rem_after = (straddles_in ? carry_remF[j] : Ff[j]) - seg_total;
                    RF[j] = rem_after + e;
                    carry_remF[j] = __shfl_sync(FULL, rem_after, 31);
                }
            }
        }
        carryT = __shfl_sync(FULL, T, 31);

        // ---- gradient of this hit (gaussiantrace_backward.cu:100-166)
        float dL_dalpha = gO[6] * (T * t - R[6]) + gO_alpha;
#pragma unroll
        for (int j = 0; j < 3; ++j) dL_dalpha += gO[j] * (T * c[j] - R[j]) + gO[3 + j] * (T * nf[j] - R[3 + j]);
        if (FEAT) {
#pragma unroll
            for (int j = 0; j < NFMAX; ++j) dL_dalpha += gF[j] * (T * feat[j] - RF[j]);
        }
        dL_dalpha /= (1.f - alpha);
        const float dL_do = dL_dalpha * G;
        const float dL_dG = dL_dalpha * op;
        const float dpu = -dL_dG * G * pu, dpv = -dL_dG * G * pv;
        const float dposx = dpu * ax + dpv * bx, dposy = dpu * ay + dpv * by, dposz = dpu * az + dpv * bz;
        const float dL_dd = gO[6] * w + (dposx * ro.dx + dposy * ro.dy + dposz * ro.dz);
        const float dL_dog = -dL_dd / dg;
        const float dL_ddg = dL_dd * og / fmaxf(1e-6f, dg * dg);
        const float dnx = m * gO[3] * w + dL_ddg * ro.dx + dL_dog * relx;
        const float dny = m * gO[4] * w + dL_ddg * ro.dy + dL_dog * rely;
        const float dnz = m * gO[5] * w + dL_ddg * ro.dz + dL_dog * relz;
        if (BULK) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous round's rows have been read
        __syncwarp();
        {
            float4 *row = reinterpret_cast<float4 *>(rows + lane * BROW);
            if (act) {
                row[0] = make_float4(-dposx - dL_dog * nx, -dposy - dL_dog * ny, -dposz - dL_dog * nz, dL_do);
                row[1] = make_float4(dpu * px, dpu * py, dpu * pz, dpv * px);
                row[2] = make_float4(dpv * py, dpv * pz, dnx, dny);
                row[3] = make_float4(dnz, 0.f, 0.f, 0.f);
#if IRGS_BWD_UNIFORM_ISSUE
                rows[lane * BROW + 64] = __uint_as_float((unsigned)g);
#endif
                const float gc[3] = {gO[0] * w, gO[1] * w, gO[2] * w};
                float Y[16];
                sh_basis(a.deg, ro.dx, ro.dy, ro.dz, Y);
#pragma unroll
                for (int v = 0; v < 12; ++v) {
                    float4 q;
                    q.x = Y[(4 * v) / 3] * gc[(4 * v) % 3];
                    q.y = Y[(4 * v + 1) / 3] * gc[(4 * v + 1) % 3];
                    q.z = Y[(4 * v + 2) / 3] * gc[(4 * v + 2) % 3];
                    q.w = Y[(4 * v + 3) / 3] * gc[(4 * v + 3) % 3];
                    row[4 + v] = q;
                }
                if (FEAT) {
#pragma unroll
                    for (int j = 0; j < NFMAX; ++j)
                        if (j < a.S) atomicAdd(a.grad_features + (size_t)g * a.S + j, gF[j] * w);
                }
            }
        }
        // ---- ray gradients: segmented sums of the per-hit terms, collected by the owning lane
        {
            float y[6] = {dposx + dL_dog * nx, dposy + dL_dog * ny, dposz + dL_dog * nz,
                          t * dposx + dL_ddg * nx, t * dposy + dL_ddg * ny, t * dposz + dL_ddg * nz};
            // where (if anywhere) this lane's OWN ray has hits in this round
            const int my_lo = max(start - base, 0);
            const bool mine_here = has && start < base + 32 && start + c_eff > base;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                float e = act ? y[j] : 0.f;
#pragma unroll
                for (int s5 = 0; s5 < 5; ++s5) {
                    const float v = __shfl_down_sync(FULL, e, 1 << s5);
                    if (dn_ok[s5]) e += v;
                }
                const float tot = __shfl_sync(FULL, e, mine_here ? my_lo : lane);
                if (mine_here) { if (j < 3) go[j] += tot; else gd[j - 3] += tot; }
            }
        }
        if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy row writes -> async proxy
        __syncwarp();
        // ---- per-surfel gradients
        if (BULK) {
            // one bulk reduction (TMA unit, UBLKRED.ADD.F32) per hit adds the whole row to the fused buffer.  ncu on the
            // 16-byte vector reductions: the L1TEX data pipe spends one wavefront per LANE on a reduction however
            // well the addresses coalesce (1.65 M wavefronts per SM and launch = hits x 16), which bounded the kernel.
#ifndef IRGS_DEBUG_SKIP_REDUCE   // timing experiment only: how much of the kernel is the reduction traffic
#if IRGS_BWD_UNIFORM_ISSUE
            // UBLKRED is a uniform-datapath instruction: one per WARP, operands in uniform registers.  Issued by every lane under
            // `if (act)` with per-lane operands, the compiler wraps it in a serial loop over the active lanes (ELECT, three R2UR
            // broadcasts, two predicate updates, the branch: eight DEPENDENT instructions per hit, 15 % of the kernel's stall
            // samples).  Instead one elected lane issues the round's reductions in an unrolled loop: the row indices wait in the pad
            // word of the rows, eight LDS / R2UR / address computations are in flight at a time (-3.4 % on the kernel).
            {
                const int n_here = min(32, total - base);
                const uint32_t src0 = (uint32_t)__cvta_generic_to_shared(rows);
                const uint32_t nbytes = (uint32_t)(4 + nvec) * 16u;
                if (elect_one()) {
#pragma unroll 8
                    for (int i = 0; i < n_here; ++i) {
                        const unsigned gi = __float_as_uint(rows[i * BROW + 64]);
                        float *dst = a.grad_fused + (size_t)gi * IRGS_GRAD_STRIDE;
                        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                     :: "l"(dst), "r"(src0 + (uint32_t)(i * BROW * 4)), "r"(nbytes) : "memory");
                    }
                }
            }
#else
            if (act) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(rows + lane * BROW);
                float *dst = a.grad_fused + (size_t)g * IRGS_GRAD_STRIDE;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             :: "l"(dst), "r"(src), "r"((4 + nvec) * 16) : "memory");
            }
#endif
#endif
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        } else {
            // sixteen consecutive lanes add one 256-byte row with 16-byte reductions
            const int n_here = min(32, total - base);
            const int col = lane & 15;
            for (int j = 0; j < n_here; j += 2) {
                const int rowi = j + (lane >> 4);
                const int g_r = __shfl_sync(FULL, g, rowi & 31);
                if (rowi < n_here && col < 4 + nvec) {
                    const float4 v = *reinterpret_cast<const float4 *>(rows + rowi * BROW + 4 * col);
                    atomicAdd(reinterpret_cast<float4 *>(a.grad_fused + (size_t)g_r * IRGS_GRAD_STRIDE) + col, v);
                }
            }
        }
        g_cur = g_nxt; pos_cur = pos_nxt; g_nxt = g_nxt2;
#if IRGS_BWD_LOCATE_ONCE && IRGS_BWD_PIPELINE
        own_c = own_n; kk_c = kk_n; own_n = own_2; kk_n = kk_2;
#endif
    }
    if (BULK) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory stays valid until it has been read
    if (valid && cnt <= a.hit_cap) {  // rays with longer lists are written by the re-trace kernel
#pragma unroll
        for (int j = 0; j < 3; ++j) { a.g_rays_o[3 * my_ray + j] = go[j]; a.g_rays_d[3 * my_ray + j] = gd[j]; }
    }
}

// One ray of the re-trace backward (the reference's scheme): same ordered passes as the forward, gradients applied
// per hit.
template <bool FEAT>
__device__ __forceinline__ void retrace_ray(const KParams &p, int64_t ray, float *bt, int *bg, float *ba, int *stack_n,
                                            float *stack_t) {
    const TraceArgs &a = p.a;
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    if (a.alpha[ray] != 0.f) {
        RayCtx r;
        load_ray(a, ray, r);
        ray_setup(r, p.qframe);
        float Y[16];
        sh_basis(a.deg, r.dx, r.dy, r.dz, Y);
        BwdState<FEAT> s;
        bwd_load<FEAT>(a, ray, s);
        float t_last = -INFINITY; int g_last = -1;
        for (;;) {
            unsigned nn = 0, nl = 0;
            const int cnt = collect_pass<false>(p, r, t_last, g_last, bt, bg, ba, stack_n, stack_t, nn, nl);
            bool term = false;
            for (int i = 0; i < cnt; ++i) {
                bwd_hit<FEAT>(a, r, Y, bg[i * TB], s);
                if (s.T < a.T_min) { term = true; break; }
            }
            if (term || cnt < KBUF) break;
            t_last = bt[(KBUF - 1) * TB]; g_last = bg[(KBUF - 1) * TB];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) { go[j] = s.go[j]; gd[j] = s.gd[j]; }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) { a.g_rays_o[3 * ray + j] = go[j]; a.g_rays_d[3 * ray + j] = gd[j]; }
}

// Re-trace backward over all rays (no hit list was saved), or -- ONLY_OVERFLOW -- over the few rays whose list did not
// fit in hit_cap: those are first compacted into a per-warp queue so that the re-trace runs with full warps.
template <bool FEAT, bool ONLY_OVERFLOW>
__global__ void __launch_bounds__(TB) trace_backward_retrace_kernel(const KParams p) {
    __shared__ float s_t[KBUF * TB];
    __shared__ int s_g[KBUF * TB];
    __shared__ float s_a[KBUF * TB];
    __shared__ long long s_queue[TB / 32][64];
    float *bt = s_t + threadIdx.x; int *bg = s_g + threadIdx.x; float *ba = s_a + threadIdx.x;
    int stack_n[STACK]; float stack_t[STACK];
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
    const TraceArgs &a = p.a;
    if (!ONLY_OVERFLOW) {
        for (;;) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(p.counter, 32ull);
            base = __shfl_sync(FULL, base, 0);
            if (base >= (unsigned long long)a.n_rays) break;
            const int64_t ray = (int64_t)base + lane;
            if (ray < a.n_rays) retrace_ray<FEAT>(p, ray, bt, bg, ba, stack_n, stack_t);
            __syncwarp();
        }
        return;
    }
    long long *queue = s_queue[threadIdx.x >> 5];
    int qn = 0;
    bool done = false;
    while (!done || qn > 0) {
        if (!done && qn < 32) {
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(p.counter, 1024ull);
            b = __shfl_sync(FULL, b, 0);
            if (b >= (unsigned long long)a.n_rays) done = true;
            else {
                // (nearly every tile holds no overflowed list at all: its 1024 counts are first checked with 32 INDEPENDENT loads per
                //  lane -- the ballot loop below is a chain of 32 dependent ones; 0.21 -> 0.11 ms per 16.4 M rays)
                {
                    int mx = 0;
#pragma unroll 8
                    for (int k = 0; k < 32; ++k) {
                        const long long idx = (long long)b + 32 * k + lane;
                        mx = max(mx, idx < a.n_rays ? __ldg(a.hit_count + idx) : 0);
                    }
                    if (!__any_sync(FULL, mx > a.hit_cap)) continue;
                }
                for (int k = 0; k < 32 && qn < 32; ++k) {   // 32 rays per round; stop early so the queue cannot overflow
                    const long long idx = (long long)b + 32 * k + lane;
                    const bool ok = idx < a.n_rays && __ldg(a.hit_count + idx) > a.hit_cap;
                    const unsigned m = __ballot_sync(FULL, ok);
                    if (ok) queue[qn + __popc(m & lt_mask)] = idx;
                    qn += __popc(m);
                    if (qn >= 32 && k < 31) {
                        // process a full warp's worth right away, then finish scanning the tile
                        __syncwarp();
                        const long long ray = queue[qn - 32 + lane];
                        qn -= 32;
                        retrace_ray<FEAT>(p, ray, bt, bg, ba, stack_n, stack_t);
                        __syncwarp();
                    }
                }
                // the tile has 1024 rays: scan whatever the early exit left
                continue;
            }
        }
        __syncwarp();
        const int take = min(qn, 32);
        if ((int)lane < take) retrace_ray<FEAT>(p, queue[qn - take + lane], bt, bg, ba, stack_n, stack_t);
        qn -= take;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ misc kernels
// gaussiantrace_intersection_test.cu:12-35: any surfel support crossed within (FLT_EPSILON, 100).
__global__ void __launch_bounds__(TB) intersection_test_kernel(const KParams p, uint8_t *__restrict__ out) {
    const TraceArgs &a = p.a;
    const int64_t ray = (int64_t)blockIdx.x * TB + threadIdx.x;
    if (ray >= a.n_rays) return;
    RayCtx r;
    load_ray(a, ray, r);
    ray_setup(r, p.qframe);
    int stack_n[STACK];
    int sp = 0, cur = 0;
    bool found = false;
    for (;;) {
        if (cur >= 0) {
            uint4 wl, wr;
                ldg256(&p.nodes[cur], wl, wr);
            const int2 d = make_int2((int)wl.w, (int)wr.w);
            float tnL, tnR;
            bool hL = slab(r, wl, 0.f, IRGS_T_SCENE_MAX, tnL);
            bool hR = slab(r, wr, 0.f, IRGS_T_SCENE_MAX, tnR);
            if (hL && hR) { if (sp < STACK) stack_n[sp++] = d.y; cur = d.x; continue; }
            else if (hL) { cur = d.x; continue; }
            else if (hR) { cur = d.y; continue; }
        } else {
            float t, alpha; int g;
            if (leaf_test(r, p.recs + (~cur), a.alpha_min, 0, t, g, alpha)) { found = true; break; }
        }
        if (sp == 0) break;
        cur = stack_n[--sp];
    }
    out[ray] = found ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ incident rays
// Per-sample table and per-point records of the generated incident rays (trace_common.cuh).
__global__ void incident_table_kernel(int S, IncTab *__restrict__ tab) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) tab[s] = incident_tab_entry(s, S);
}
__global__ void incident_point_kernel(const float *__restrict__ normals, const float *__restrict__ azimuth, int64_t n_points,
                                      IncPoint *__restrict__ pts) {
    const int64_t pt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= n_points) return;
    pts[pt] = incident_point(__ldg(normals + 3 * pt), __ldg(normals + 3 * pt + 1), __ldg(normals + 3 * pt + 2), azimuth != nullptr,
                             azimuth ? __ldg(azimuth + pt) : 0.f);
}

// The generated rays written out (for callers that also shade with the directions, and for tests): one ray per thread, the
// point record computed on the spot (no handle, hence no scratch, on this entry).
__global__ void incident_rays_kernel(const float *__restrict__ position, const float *__restrict__ normals,
                                     const float *__restrict__ azimuth, int64_t n_rays, int S, float t_min,
                                     const IncTab *__restrict__ tab, float *__restrict__ rays_o, float *__restrict__ rays_d) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= n_rays) return;
    const int64_t pt = ray / S;
    const int s = (int)(ray - pt * S);
    const IncPoint ip = incident_point(__ldg(normals + 3 * pt), __ldg(normals + 3 * pt + 1), __ldg(normals + 3 * pt + 2),
                                       azimuth != nullptr, azimuth ? __ldg(azimuth + pt) : 0.f);
    const IncidentSample q = incident_sample(ip, load_inc_tab(tab + s), azimuth != nullptr);
    const float dx = __fdiv_rn(q.vx, q.len), dy = __fdiv_rn(q.vy, q.len), dz = __fdiv_rn(q.vz, q.len);
    if (rays_o) {
        rays_o[3 * ray] = __fadd_rn(__ldg(position + 3 * pt), __fmul_rn(dx, t_min));
        rays_o[3 * ray + 1] = __fadd_rn(__ldg(position + 3 * pt + 1), __fmul_rn(dy, t_min));
        rays_o[3 * ray + 2] = __fadd_rn(__ldg(position + 3 * pt + 2), __fmul_rn(dz, t_min));
    }
    if (rays_d) { rays_d[3 * ray] = dx; rays_d[3 * ray + 1] = dy; rays_d[3 * ray + 2] = dz; }
}

// The rays of a generator (camera rays here; incident rays have their own entry) exactly as load_ray() hands them to the tracer.
__global__ void generated_rays_kernel(const TraceArgs a, float *__restrict__ rays_o, float *__restrict__ rays_d) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= a.n_rays) return;
    RayCtx r;
    load_ray(a, ray, r);
    if (rays_o) { rays_o[3 * ray] = r.ox; rays_o[3 * ray + 1] = r.oy; rays_o[3 * ray + 2] = r.oz; }
    if (rays_d) { rays_d[3 * ray] = r.dx; rays_d[3 * ray + 1] = r.dy; rays_d[3 * ray + 2] = r.dz; }
}

// Chain rule from the per-ray gradients of the tracer back to the shading point: one warp per point.
//   o = x + t_min d,  d = v / |v|,  v = R(n) zs   =>   dL/dx = sum_s g_o,   dL/dd = g_d + t_min g_o,
//   dL/dv = (dL/dd - d (d . dL/dd)) / |v|,   dL/dR = sum_s dL/dv zs^T   (zero on the constant -identity branch).
// dL/dn follows from dL/dR through rotation_between_z (graphics_utils.py:133-165): with v1 = -n.y, v2 = n.x, c = max(n.z + 1, 1e-7),
//   R = [[1 - v2^2/c, v1 v2/c, v2], [v1 v2/c, 1 - v1^2/c, -v1], [-v2, v1, 1 - (v1^2 + v2^2)/c]].
__global__ void __launch_bounds__(128) incident_backward_kernel(const float *__restrict__ position,
                                                                const float *__restrict__ normals,
                                                                const float *__restrict__ azimuth, int64_t n_points, int S,
                                                                float t_min, const IncTab *__restrict__ tab,
                                                                const float *__restrict__ g_rays_o,
                                                                const float *__restrict__ g_rays_d,
                                                                float *__restrict__ grad_position,
                                                                float *__restrict__ grad_normal) {
    const int64_t pt = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pt >= n_points) return;
    const float nx = __ldg(normals + 3 * pt), ny = __ldg(normals + 3 * pt + 1), nz = __ldg(normals + 3 * pt + 2);
    const IncPoint ip = incident_point(nx, ny, nz, azimuth != nullptr, azimuth ? __ldg(azimuth + pt) : 0.f);
    float acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 0.f;
    for (int s = lane; s < S; s += 32) {
        const int64_t ray = pt * S + s;
        const IncidentSample q = incident_sample(ip, load_inc_tab(tab + s), azimuth != nullptr);
        const float dx = __fdiv_rn(q.vx, q.len), dy = __fdiv_rn(q.vy, q.len), dz = __fdiv_rn(q.vz, q.len);
        const float gox = g_rays_o[3 * ray], goy = g_rays_o[3 * ray + 1], goz = g_rays_o[3 * ray + 2];
        const float gdx = g_rays_d[3 * ray] + t_min * gox, gdy = g_rays_d[3 * ray + 1] + t_min * goy,
                    gdz = g_rays_d[3 * ray + 2] + t_min * goz;
        acc[0] += gox; acc[1] += goy; acc[2] += goz;
        if (q.rotated && q.len > 1e-12f) {
            const float dd = dx * gdx + dy * gdy + dz * gdz;
            const float gvx = (gdx - dx * dd) / q.len, gvy = (gdy - dy * dd) / q.len, gvz = (gdz - dz * dd) / q.len;
            acc[3] += gvx * q.zx; acc[4] += gvx * q.zy; acc[5] += gvx * q.zz;
            acc[6] += gvy * q.zx; acc[7] += gvy * q.zy; acc[8] += gvy * q.zz;
            acc[9] += gvz * q.zx; acc[10] += gvz * q.zy; acc[11] += gvz * q.zz;
        }
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) grad_position[3 * pt + j] = acc[j];
        const float *G = acc + 3;   // dL/dR row-major
        const float v1 = -ny, v2 = nx, c = fmaxf(nz + 1.0f, 1e-7f);
        const float gv1 = (G[1] + G[3]) * v2 / c - (G[4] + G[8]) * 2.0f * v1 / c - G[5] + G[7];
        const float gv2 = -(G[0] + G[8]) * 2.0f * v2 / c + (G[1] + G[3]) * v1 / c + G[2] - G[6];
        const float gc = (G[0] * v2 * v2 - (G[1] + G[3]) * v1 * v2 + G[4] * v1 * v1 + G[8] * (v1 * v1 + v2 * v2)) / (c * c);
        grad_normal[3 * pt] = gv2;
        grad_normal[3 * pt + 1] = -gv1;
        grad_normal[3 * pt + 2] = (nz + 1.0f > 1e-7f) ? gc : 0.0f;
    }
}

__global__ void unpack_grads_kernel(const float *__restrict__ fused, int64_t n, int K, float *__restrict__ gm,
                                    float *__restrict__ go, float *__restrict__ gru, float *__restrict__ grv,
                                    float *__restrict__ gn, float *__restrict__ gsh) {
    // one thread per (surfel, float of the 64-float row): coalesced reads, near-coalesced writes
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * IRGS_GRAD_STRIDE) return;
    const int64_t g = idx / IRGS_GRAD_STRIDE;
    const int f = (int)(idx % IRGS_GRAD_STRIDE);
    const float v = fused[idx];
    if (f < 3) gm[3 * g + f] = v;
    else if (f == 3) go[g] = v;
    else if (f < 7) gru[3 * g + f - 4] = v;
    else if (f < 10) grv[3 * g + f - 7] = v;
    else if (f < 13) gn[3 * g + f - 10] = v;
    else if (f >= 16) {
        int k = (f - 16) / 3;
        if (k < K) gsh[(g * K + k) * 3 + (f - 16) % 3] = v;
    }
}
__global__ void zero_sh_tail_kernel(int64_t n, int K, float *__restrict__ gsh) {
    // coefficients k >= 16 never receive gradient
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int tail = (K - 16) * 3;
    if (idx >= n * tail) return;
    const int64_t g = idx / tail;
    gsh[g * K * 3 + 48 + idx % tail] = 0.f;
}

// ------------------------------------------------------------------------------------------------ launchers
static int persistent_grid(irgs_tracer *h, const void *kernel) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TB, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    return h->sm_count * per_sm;
}

static KParams make_params(irgs_tracer *h, const TraceArgs &a, int slot) {
    KParams p;
    p.a = a;
    p.nodes = h->qnodes;
    p.nodes4 = h->qnodes4;
    p.qframe = h->scene + 12;
    p.recs = h->recs;
    p.inv_order = h->inv_order;
    p.counter = h->counter + slot;
    p.stats = h->stats;
    return p;
}

int launch_pack_records(irgs_tracer *h, const TraceArgs &a, cudaStream_t s) {
    ++h->pack_epoch;
    const int n = (int)h->n;
    pack_records_kernel<<<(n + 255) / 256, 256, 0, s>>>(h->order, n, a.means, a.opacity, a.ru, a.rv, a.normals, a.alpha_min, h->recs,
                                                         h->inv_order);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

template <typename Kern>
static int launch_persistent(irgs_tracer *h, Kern kern, const KParams &p, int64_t n_rays, cudaStream_t s) {
    IRGS_CHECK(cudaMemsetAsync(p.counter, 0, sizeof(unsigned long long), s));
    int grid = persistent_grid(h, reinterpret_cast<const void *>(kern));
    int64_t need = (n_rays + TB - 1) / TB;
    if (need < grid) grid = (int)(need > 0 ? need : 1);
    kern<<<grid, TB, 0, s>>>(p);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

// Shared-memory carve-out hint for the replay kernel (irgs_set_option("bwd_carveout_pct"); -1 = the driver's default, which
// measured best: see launch_fwd in trace_fwd.cu).
template <typename Kern>
static void set_carveout(Kern kern, int pct) {
    if (pct >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
}

int launch_trace_backward(irgs_tracer *h, const TraceArgs &a_in, cudaStream_t s) {
    const int slot = slot_for(h, s);
    if (slot < 0) return 1;
    // colour cache: valid when the last forward that saved hit lists on THIS stream wrote exactly this list (same stream = the
    // kernels run in the order of the calls, so a later forward cannot overwrite the block before this replay has read it)
    TraceArgs a = a_in;
    a.hit_rgb = nullptr; a.rgb_cap = 0;
    if (a.hits != nullptr) {
        std::lock_guard<std::mutex> lock(h->slot_mutex);
        if (h->hit_rgb[slot] != nullptr && h->hit_rgb_key[slot] == a.hits && h->hit_rgb_rays[slot] == a.n_rays && h->color_cache > 0) {
            a.hit_rgb = h->hit_rgb[slot]; a.rgb_cap = h->hit_rgb_cc[slot];
        }
    }
    KParams p = make_params(h, a, slot);
    const bool feat = a.S > 0;
    if (a.hits != nullptr && a.hit_count != nullptr) {
        const unsigned grid = (unsigned)((a.n_rays + TB - 1) / TB);
        auto flat_grid = [&](const void *) { return (unsigned)((a.n_rays + FTB - 1) / FTB); };
        if (h->bwd_mode == 1) {   // thread-per-ray replay (kept for comparison: irgs_set_option("bwd_mode", 1))
            if (feat) trace_backward_replay_kernel<true><<<grid, TB, 0, s>>>(p);
            else trace_backward_replay_kernel<false><<<grid, TB, 0, s>>>(p);
        } else {
            if (h->bwd_mode == 2) {   // 16-byte vector reductions instead of bulk reductions (comparison)
                if (feat) trace_backward_flat_kernel<true, false><<<flat_grid((const void *)trace_backward_flat_kernel<true, false>), FTB, 0, s>>>(p);
                else trace_backward_flat_kernel<false, false><<<flat_grid((const void *)trace_backward_flat_kernel<false, false>), FTB, 0, s>>>(p);
            } else {
                if (feat) {
                    set_carveout(trace_backward_flat_kernel<true, true>, h->bwd_carveout_pct);
                    trace_backward_flat_kernel<true, true><<<flat_grid((const void *)trace_backward_flat_kernel<true, true>), FTB, 0, s>>>(p);
                } else {
                    set_carveout(trace_backward_flat_kernel<false, true>, h->bwd_carveout_pct);
                    trace_backward_flat_kernel<false, true><<<flat_grid((const void *)trace_backward_flat_kernel<false, true>), FTB, 0, s>>>(p);
                }
            }
        }
        count_launch();
        IRGS_CHECK(cudaGetLastError());
        return feat ? launch_persistent(h, trace_backward_retrace_kernel<true, true>, p, a.n_rays, s)
                    : launch_persistent(h, trace_backward_retrace_kernel<false, true>, p, a.n_rays, s);
    }
    return feat ? launch_persistent(h, trace_backward_retrace_kernel<true, false>, p, a.n_rays, s)
                : launch_persistent(h, trace_backward_retrace_kernel<false, false>, p, a.n_rays, s);
}

int launch_intersection_test(irgs_tracer *h, const TraceArgs &a, uint8_t *out, cudaStream_t s) {
    KParams p = make_params(h, a, 0);   // no work counter in this kernel
    const unsigned grid = (unsigned)((a.n_rays + TB - 1) / TB);
    intersection_test_kernel<<<grid, TB, 0, s>>>(p, out);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

// The per-sample table of a sample count: built once per (device, S) and kept for the life of the process (32 bytes per
// sample).  The first request fills it on `s` and waits, so that every later user on any stream finds it complete.
const IncTab *incident_table(int sample_num, cudaStream_t s) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, IncTab *> cache;
    int dev = 0;
    if (!check(cudaGetDevice(&dev), "cudaGetDevice")) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, sample_num});
    if (it != cache.end()) return it->second;
    IncTab *tab = nullptr;
    if (!check(cudaMalloc(&tab, sizeof(IncTab) * (size_t)sample_num), "cudaMalloc")) return nullptr;
    incident_table_kernel<<<(sample_num + 127) / 128, 128, 0, s>>>(sample_num, tab);
    count_launch();
    if (!check(cudaGetLastError(), "incident_table_kernel") || !check(cudaStreamSynchronize(s), "cudaStreamSynchronize")) {
        cudaFree(tab);
        return nullptr;
    }
    cache[{dev, sample_num}] = tab;
    return tab;
}

int launch_incident_prepare(irgs_tracer *h, TraceArgs &a, cudaStream_t s) {
    if (a.gen_pos == nullptr) return 0;
    const int slot = slot_for(h, s);
    if (slot < 0) return 1;
    a.gen_tab = incident_table(a.gen_S, s);
    if (!a.gen_tab) return 1;
    if (a.gen_P > h->inc_cap[slot]) {
        IRGS_CHECK(cudaStreamSynchronize(s));   // earlier users of this slot's block (all on this stream) are done
        if (h->inc_pts[slot]) cudaFree(h->inc_pts[slot]);
        h->inc_pts[slot] = nullptr;
        IRGS_CHECK(cudaMalloc(&h->inc_pts[slot], sizeof(IncPoint) * (size_t)a.gen_P));
        h->inc_cap[slot] = a.gen_P;
    }
    IncPoint *pts = static_cast<IncPoint *>(h->inc_pts[slot]);
    incident_point_kernel<<<(unsigned)((a.gen_P + 255) / 256), 256, 0, s>>>(a.gen_nrm, a.gen_azim, a.gen_P, pts);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    a.gen_pts = pts;
    return 0;
}

int launch_incident_rays(const float *position, const float *normals, const float *azimuth, int64_t n_points, int sample_num,
                         float t_min, float *rays_o, float *rays_d, cudaStream_t s) {
    const int64_t n_rays = n_points * sample_num;
    if (n_rays <= 0) return 0;
    const IncTab *tab = incident_table(sample_num, s);
    if (!tab) return 1;
    incident_rays_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, s>>>(position, normals, azimuth, n_rays, sample_num, t_min,
                                                                          tab, rays_o, rays_d);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int launch_incident_backward(const float *position, const float *normals, const float *azimuth, int64_t n_points,
                             int sample_num, float t_min, const float *g_rays_o, const float *g_rays_d, float *grad_position,
                             float *grad_normal, cudaStream_t s) {
    if (n_points <= 0) return 0;
    const IncTab *tab = incident_table(sample_num, s);
    if (!tab) return 1;
    incident_backward_kernel<<<(unsigned)((n_points * 32 + 127) / 128), 128, 0, s>>>(
        position, normals, azimuth, n_points, sample_num, t_min, tab, g_rays_o, g_rays_d, grad_position, grad_normal);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int launch_generated_rays(const TraceArgs &a, float *rays_o, float *rays_d, cudaStream_t s) {
    if (a.n_rays <= 0) return 0;
    generated_rays_kernel<<<(unsigned)((a.n_rays + 255) / 256), 256, 0, s>>>(a, rays_o, rays_d);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int launch_unpack_grads(const float *fused, int64_t n, int K, float *gm, float *go, float *gru, float *grv, float *gn,
                        float *gsh, cudaStream_t s) {
    const int64_t total = n * IRGS_GRAD_STRIDE;
    unpack_grads_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(fused, n, K, gm, go, gru, grv, gn, gsh);
    count_launch();
    if (K > 16) {
        const int64_t tail = n * (K - 16) * 3;
        zero_sh_tail_kernel<<<(unsigned)((tail + 255) / 256), 256, 0, s>>>(n, K, gsh);
        count_launch();
    }
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace irgs
