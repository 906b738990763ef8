// Device helpers shared by the tracing kernels (sm_100a).  See trace.cu / trace_fwd.cu for the kernels.
#pragma once
#include <cfloat>
#include <climits>

#include "internal.cuh"

namespace irgs {

constexpr int TB = 128;     // threads per block
constexpr int KBUF = 16;    // k-buffer depth (MAX_BUFFER_SIZE, auxiliary.h:10)
constexpr int STACK = 96;   // traversal stack entries: the 4-wide walk stacks up to 3 per level, 31 levels at most (binary depth <= 62)
constexpr float T_EPS = 1.1920929e-07f;  // FLT_EPSILON tmin, gaussiantrace_forward.cu:38
constexpr int NFMAX = IRGS_MAX_FEATURES;

// auxiliary.h:16-33
__device__ constexpr float SH_C0 = 0.28209479177387814f;
__device__ constexpr float SH_C1 = 0.4886025119029199f;
__device__ constexpr float SH_C2_0 = 1.0925484305920792f, SH_C2_1 = -1.0925484305920792f, SH_C2_2 = 0.31539156525252005f,
                           SH_C2_3 = -1.0925484305920792f, SH_C2_4 = 0.5462742152960396f;
__device__ constexpr float SH_C3_0 = -0.5900435899266435f, SH_C3_1 = 2.890611442640554f, SH_C3_2 = -0.4570457994644658f,
                           SH_C3_3 = 0.3731763325901154f, SH_C3_4 = -0.4570457994644658f, SH_C3_5 = 1.445305721320277f,
                           SH_C3_6 = -0.5900435899266435f;

struct KParams {
    TraceArgs a;
    const QNode *nodes;
    const QNode4 *nodes4;  // forward walk
    const float *qframe;   // [6] quantisation frame: lo xyz, extent xyz
    const SurfelRec *recs;
    const int *inv_order;   // surfel id -> leaf position
    unsigned long long *counter;
    unsigned long long *stats;
};

// 256-bit read-only loads (sm_100a: LDG.E.256): a 32-byte quantised node or half a surfel record in ONE request.  The
// walk is bound by the L1TEX data pipe (ncu: l1tex__data_pipe_lsu_wavefronts 73 % of peak), not by DRAM or L2.
#ifndef IRGS_NODE_HINT
#define IRGS_NODE_HINT 1   // 1: tree nodes are loaded with L1::evict_last (C3 step 617 -> 619 M rays/s), 0: no hint
#endif
__device__ __forceinline__ void ldg256(const void *p, uint4 &a, uint4 &b) {
#if IRGS_NODE_HINT
    asm("ld.global.nc.L1::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
        : "l"(p));
#else
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
        : "l"(p));
#endif
}
// The same load for STREAMING data (surfel records, SH rows: read once per use, no reuse worth caching): no L1 allocation, so
// that the small L1 (what the shared-memory carve-out leaves) keeps the top levels of the tree.  IRGS_STREAM_HINT=0: plain.
#ifndef IRGS_STREAM_HINT
#define IRGS_STREAM_HINT 1
#endif
__device__ __forceinline__ void ldg256_stream(const void *p, float4 &a, float4 &b) {
#if IRGS_STREAM_HINT
    asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
#else
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
#endif
}
__device__ __forceinline__ void ldg256(const void *p, float4 &a, float4 &b) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
}

// fixed-order dot product, bit-identical to dot3() of oracle/surfel_oracle.c
__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
}

// SH basis Y_k(d), zero beyond (deg+1)^2  (auxiliary.h:52-89)
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float Y[16]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) Y[k] = 0.f;
    Y[0] = SH_C0;
    if (deg > 0) {
        Y[1] = -SH_C1 * y; Y[2] = SH_C1 * z; Y[3] = -SH_C1 * x;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            Y[4] = SH_C2_0 * xy; Y[5] = SH_C2_1 * yz; Y[6] = SH_C2_2 * (2.0f * zz - xx - yy);
            Y[7] = SH_C2_3 * xz; Y[8] = SH_C2_4 * (xx - yy);
            if (deg > 2) {
                Y[9] = SH_C3_0 * y * (3.0f * xx - yy);
                Y[10] = SH_C3_1 * xy * z;
                Y[11] = SH_C3_2 * y * (4.0f * zz - xx - yy);
                Y[12] = SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                Y[13] = SH_C3_4 * x * (4.0f * zz - xx - yy);
                Y[14] = SH_C3_5 * z * (xx - yy);
                Y[15] = SH_C3_6 * x * (xx - 3.0f * yy);
            }
        }
    }
}

// colour = max(0, 0.5 + sum_k Y_k sh[g,k])   (auxiliary.h:52-89).  K == 16: twelve 16-byte loads at most.
__device__ __forceinline__ void sh_color(const float *__restrict__ shs, int K, int deg, int g, const float Y[16],
                                         float c[3]) {
    const int nb = (deg + 1) * (deg + 1);
    float acc[3] = {0.f, 0.f, 0.f};
    if (K == 16) {
        const float4 *p = reinterpret_cast<const float4 *>(shs + (size_t)g * 48);
        const int nvec = (nb * 3 + 3) >> 2;
        const bool wide = (reinterpret_cast<size_t>(shs) & 31) == 0;   // rows are 192 B: 32-byte aligned iff the base is
#pragma unroll
        for (int v = 0; v < 12; v += 2) {
            if (v < nvec) {
                float4 q, q2;
                if (wide) ldg256_stream(p + v, q, q2);
                else { q = __ldg(p + v); q2 = __ldg(p + v + 1); }
                acc[(4 * v) % 3] += Y[(4 * v) / 3] * q.x;
                acc[(4 * v + 1) % 3] += Y[(4 * v + 1) / 3] * q.y;
                acc[(4 * v + 2) % 3] += Y[(4 * v + 2) / 3] * q.z;
                acc[(4 * v + 3) % 3] += Y[(4 * v + 3) / 3] * q.w;
                if (v + 1 < nvec) {
                    acc[(4 * v + 4) % 3] += Y[(4 * v + 4) / 3] * q2.x;
                    acc[(4 * v + 5) % 3] += Y[(4 * v + 5) / 3] * q2.y;
                    acc[(4 * v + 6) % 3] += Y[(4 * v + 6) / 3] * q2.z;
                    acc[(4 * v + 7) % 3] += Y[(4 * v + 7) / 3] * q2.w;
                }
            }
        }
    } else {
        const float *p = shs + (size_t)g * K * 3;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k < nb) {
                acc[0] += Y[k] * __ldg(p + 3 * k);
                acc[1] += Y[k] * __ldg(p + 3 * k + 1);
                acc[2] += Y[k] * __ldg(p + 3 * k + 2);
            }
        }
    }
    c[0] = fmaxf(acc[0] + 0.5f, 0.f);
    c[1] = fmaxf(acc[1] + 0.5f, 0.f);
    c[2] = fmaxf(acc[2] + 0.5f, 0.f);
}

// ------------------------------------------------------------------------------------------------ traversal
struct RayCtx {
    float ox, oy, oz, dx, dy, dz;
    // slab-test constants for quantised planes: t = fma(2^23 + q, a, b)
    float ax, ay, az, bx, by, bz;
};

// The slab test evaluates t = (plane - o) / d as ONE fma per plane, with ONE byte-permute to decode the plane: a
// plane stored as the 16-bit code q on the frame (lo, cell) is lo + q * cell; the float v = 2^23 + q is assembled
// by PRMT from the code's two bytes and the constant 0x4B000000, and t = v * (cell/d) + (lo - 2^23 cell - o)/d.
// The two rounded constants make the decoded plane uncertain by up to ~one cell in SPACE (2^23 * 2^-24 = half a
// cell each), independent of how small a direction component is (the error in t scales with 1/d exactly like the
// t-extent of a spatial pad does); the build therefore moves every quantised plane two cells outward on top of the
// float pad of the leaf bounds (lbvh.cu), and the interval test needs no slack.  Ray origins are assumed to lie
// within ~16 scene diameters (beyond that the reference's own float32 plane arithmetic is equally fuzzy).
__device__ __forceinline__ void ray_setup(RayCtx &r, const float *__restrict__ qframe) {
    const float tiny = 1e-30f;  // a zero component would give 0 * inf = NaN
    const float sx = fabsf(r.dx) > tiny ? r.dx : copysignf(tiny, r.dx);
    const float sy = fabsf(r.dy) > tiny ? r.dy : copysignf(tiny, r.dy);
    const float sz = fabsf(r.dz) > tiny ? r.dz : copysignf(tiny, r.dz);
    const float ix = 1.0f / sx, iy = 1.0f / sy, iz = 1.0f / sz;
    const float lx = __ldg(qframe), ly = __ldg(qframe + 1), lz = __ldg(qframe + 2);
    const float cx = __ldg(qframe + 3), cy = __ldg(qframe + 4), cz = __ldg(qframe + 5);
    const float two23 = 8388608.0f;
    r.ax = cx * ix; r.ay = cy * iy; r.az = cz * iz;
    r.bx = (lx - two23 * cx - r.ox) * ix; r.by = (ly - two23 * cy - r.oy) * iy; r.bz = (lz - two23 * cz - r.oz) * iz;
}

__device__ __forceinline__ float qlo16(unsigned w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)); }
__device__ __forceinline__ float qhi16(unsigned w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)); }

// one child of a quantised node: w = (lo.x|hi.x<<16, lo.y|hi.y<<16, lo.z|hi.z<<16, ref).  t grows with the plane code when the
// slope a = cell / d is positive: the NEAR plane of an axis is its lo code then, its hi code otherwise -- picked by the byte-
// permute selector (0x7610 = low half, 0x7632 = high half; the far plane's selector is the near one ^ 0x22), so no per-axis
// min / max is needed: 6 PRMT + 6 FFMA + two 3-input min / max chains per child.
__device__ __forceinline__ bool slab(const RayCtx &r, const uint4 w, float t_lo, float t_hi, float &tn) {
    const unsigned nx = 0x7610u + (__float_as_uint(r.ax) >> 31) * 0x22u, ny = 0x7610u + (__float_as_uint(r.ay) >> 31) * 0x22u,
                   nz = 0x7610u + (__float_as_uint(r.az) >> 31) * 0x22u;
    const float x0 = __fmaf_rn(__uint_as_float(__byte_perm(w.x, 0x4B000000u, nx)), r.ax, r.bx);
    const float x1 = __fmaf_rn(__uint_as_float(__byte_perm(w.x, 0x4B000000u, nx ^ 0x22u)), r.ax, r.bx);
    const float y0 = __fmaf_rn(__uint_as_float(__byte_perm(w.y, 0x4B000000u, ny)), r.ay, r.by);
    const float y1 = __fmaf_rn(__uint_as_float(__byte_perm(w.y, 0x4B000000u, ny ^ 0x22u)), r.ay, r.by);
    const float z0 = __fmaf_rn(__uint_as_float(__byte_perm(w.z, 0x4B000000u, nz)), r.az, r.bz);
    const float z1 = __fmaf_rn(__uint_as_float(__byte_perm(w.z, 0x4B000000u, nz ^ 0x22u)), r.az, r.bz);
    tn = fmaxf(fmaxf(x0, y0), fmaxf(z0, t_lo));
    const float tf = fminf(fminf(x1, y1), fminf(z1, t_hi));
    return (int)w.w != IRGS_CHILD_NONE && tn <= tf;
}

// Plane hit of a packed record, in two stages so that the second half of the record is only fetched for pairs that pass
// the first.  Arithmetic order == eval_surfel() of oracle/surfel_oracle.c (gaussiantrace_forward.cu:61-81).
// Stage 1 (r0 = mu, support radius^2; r1 = normal, id): depth, range / facing tests, hit point relative to mu, and the
// conservative early reject |pos|^2 > radius^2 (no alpha >= alpha_min is possible there; see pack_records_kernel).
__device__ __forceinline__ bool leaf_stage1(const RayCtx &r, const float4 r0, const float4 r1, int back_culling,
                                            float &t_out, int &g_out, float &px, float &py, float &pz) {
    float relx = __fsub_rn(r.ox, r0.x), rely = __fsub_rn(r.oy, r0.y), relz = __fsub_rn(r.oz, r0.z);
    float og = dot3_rn(r1.x, r1.y, r1.z, relx, rely, relz);
    float dg = dot3_rn(r1.x, r1.y, r1.z, r.dx, r.dy, r.dz);
    float dg2 = __fmul_rn(dg, dg);
    float den = fmaxf(1e-6f, dg2);
    float t = __fdiv_rn(__fmul_rn(-og, dg), den);
    t_out = t; g_out = __float_as_int(r1.w);
    px = __fmaf_rn(t, r.dx, relx); py = __fmaf_rn(t, r.dy, rely); pz = __fmaf_rn(t, r.dz, relz);
    if (!(dg2 >= 1e-6f)) return false;  // grazing pair: the clamped formula is no longer the geometric hit (see oracle)
    if (!(t > T_EPS && t < IRGS_T_SCENE_MAX)) return false;
    if (back_culling && !(-dg > 0.0f)) return false;
    return !(__fmaf_rn(pz, pz, __fmaf_rn(py, py, __fmul_rn(px, px))) > r0.w);
}
// Stage 2 (r2 = ru, rv.x; r3 = rv.yz, opacity): the Gaussian response.  Returns true for a compositing candidate.
__device__ __forceinline__ bool leaf_stage2(const float4 r2, const float4 r3, float px, float py, float pz, float alpha_min,
                                            float &alpha_out) {
    float pu = dot3_rn(r2.x, r2.y, r2.z, px, py, pz);
    float pv = dot3_rn(r2.w, r3.x, r3.y, px, py, pz);
    float power = __fmul_rn(-0.5f, __fadd_rn(__fmul_rn(pu, pu), __fmul_rn(pv, pv)));
    float alpha = fminf(0.99f, __fmul_rn(r3.z, __expf(power)));
    alpha_out = alpha;
    return alpha >= alpha_min;
}

// Diagnostics (statistics builds of the forward kernel only): is this one of the GRAZING pairs the hit test drops
// (|n.d| < 1e-3, leaf_stage1) although the ray geometrically crosses the surfel's support inside the depth range -- i.e. a
// pair the reference, whose candidates are proxy-triangle hits at the TRUE depth, would have evaluated with its clamped depth
// formula (gaussiantrace_forward.cu:61-81)?  *composites: that clamped evaluation would have passed alpha >= alpha_min.
__device__ __forceinline__ bool grazing_dropped(const RayCtx &r, const float4 r0, const float4 r1, const float4 r2, const float4 r3,
                                                float alpha_min, bool *composites) {
    const float relx = r.ox - r0.x, rely = r.oy - r0.y, relz = r.oz - r0.z;
    const float og = dot3_rn(r1.x, r1.y, r1.z, relx, rely, relz), dg = dot3_rn(r1.x, r1.y, r1.z, r.dx, r.dy, r.dz);
    *composites = false;
    if (dg * dg >= 1e-6f || dg == 0.f) return false;
    const float tt = -og / dg;
    if (!(tt > T_EPS && tt < IRGS_T_SCENE_MAX)) return false;
    const float qx = relx + tt * r.dx, qy = rely + tt * r.dy, qz = relz + tt * r.dz;
    if (qx * qx + qy * qy + qz * qz > r0.w) return false;
    const float tc = (-og * dg) / 1e-6f;   // the reference's clamped depth
    const float px = relx + tc * r.dx, py = rely + tc * r.dy, pz = relz + tc * r.dz;
    float a;
    *composites = leaf_stage2(r2, r3, px, py, pz, alpha_min, a);
    return true;
}

__device__ __forceinline__ bool leaf_test(const RayCtx &r, const SurfelRec *__restrict__ rec, float alpha_min,
                                          int back_culling, float &t_out, int &g_out, float &alpha_out) {
    float4 q0, q1, q2, q3;
    float px, py, pz;
    ldg256(&rec->r0, q0, q1);
    alpha_out = 0.f;
    if (!leaf_stage1(r, q0, q1, back_culling, t_out, g_out, px, py, pz)) return false;
    ldg256(&rec->r2, q2, q3);
    return leaf_stage2(q2, q3, px, py, pz, alpha_min, alpha_out);
}

__device__ __forceinline__ bool key_less(float ta, int ga, float tb, int gb) { return ta < tb || (ta == tb && ga < gb); }

// ------------------------------------------------------------------------------------------------ incident rays
// Sample `s` of the Fibonacci hemisphere around `normal`: the reference's fibonacci_sphere_sampling + rotation_between_z
// (utils/graphics_utils.py:19-47,133-165), restated and pinned in oracle/incident.py; azim is the per-point `rand * 2 pi`
// of the training mode (has_azim = false otherwise).  The work is split by what it depends on, so that nothing is
// recomputed per ray that 256 rays share:
//   * IncTab   per SAMPLE index (a table of S entries, filled once per S by incident_table_kernel): the float32 angle
//              delta = 2.39996.. * s exactly as the reference rounds it, sin / cos of THAT float32 value in double
//              precision, z and the radius sqrt(1 - z^2);
//   * IncPoint per shading POINT (incident_point: one 64-byte record, precomputed for the tracing kernels, computed once per
//              warp by the kernels that own a point per warp): the rotation taking +z to the normal -- the reference's
//              four IEEE divisions -- and sin / cos of the azimuth in double precision;
//   * per RAY  (incident_sample) only the angle addition, the 3x3 product and the normalisation remain.
// The reference evaluates sin / cos of theta = fl(azim + delta), a float32 SUM whose rounding error (up to 6e-5 rad at
// theta ~ 1200) is far above the 4e-7 the directions are compared at; the two-sum below recovers that error exactly and the
// angle addition is corrected to first order (second order: 2e-9), in double precision (half rate on B200), which is
// MORE accurate than sincosf of the rounded sum was: directions agree with the unmodified reference to <= 4e-7.
struct __align__(32) IncTab { double sd, cd; float delta, z, rad, pad; };
struct __align__(32) IncPoint { float r[9]; float az; int rotated; int pad; double saz, caz; };
static_assert(sizeof(IncTab) == 32 && sizeof(IncPoint) == 64, "incident-ray records are 32 / 64 bytes");

__device__ __forceinline__ IncTab incident_tab_entry(int s, int S) {
    IncTab t;
    const float idx = (float)s;
    t.z = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fmul_rn(2.0f, idx), (float)(2 * S - 1))), 0.17364817766693033f);
    t.rad = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(t.z, t.z)));
    t.delta = __fmul_rn(2.399963229728653f, idx);
    sincos((double)t.delta, &t.sd, &t.cd);
    t.pad = 0.f;
    return t;
}

__device__ __forceinline__ IncPoint incident_point(float nx, float ny, float nz, bool has_azim, float azim) {
    IncPoint p;
    p.rotated = __fadd_rn(nz, 1.0f) > 0.0f ? 1 : 0;
    p.pad = 0;
    p.az = has_azim ? azim : 0.f;
    p.saz = 0.0; p.caz = 1.0;
    if (has_azim) sincos((double)azim, &p.saz, &p.caz);
    if (p.rotated) {
        const float v1 = -ny, v2 = nx;
        const float c = fmaxf(__fadd_rn(nz, 1.0f), 1e-7f);
        const float v11 = __fmul_rn(v1, v1), v22 = __fmul_rn(v2, v2), v12 = __fmul_rn(v1, v2);
        p.r[0] = __fadd_rn(1.0f, __fdiv_rn(-v22, c)); p.r[1] = __fdiv_rn(v12, c); p.r[2] = v2;
        p.r[3] = p.r[1]; p.r[4] = __fadd_rn(1.0f, __fdiv_rn(-v11, c)); p.r[5] = -v1;
        p.r[6] = -v2; p.r[7] = v1; p.r[8] = __fadd_rn(1.0f, __fdiv_rn(__fsub_rn(-v22, v11), c));
    } else {   // graphics_utils.py:163-164: -identity
        p.r[0] = -1.f; p.r[1] = 0.f; p.r[2] = 0.f; p.r[3] = 0.f; p.r[4] = -1.f; p.r[5] = 0.f; p.r[6] = 0.f; p.r[7] = 0.f; p.r[8] = -1.f;
    }
    return p;
}

// zs = the sample around +z, v = R zs (not yet normalised), len = max(|v|, 1e-12); rotated = false on the -identity branch.
struct IncidentSample { float zx, zy, zz, vx, vy, vz, len; bool rotated; };
__device__ __forceinline__ IncidentSample incident_sample(const IncPoint &p, const IncTab &t, bool has_azim) {
    IncidentSample o;
    float sn, cs;
    if (has_azim) {
        // theta = fl(az + delta) = az + delta - err (two-sum: exact); sin / cos(az + delta) by angle addition, then - err
        const float th = __fadd_rn(p.az, t.delta);
        const float bv = __fsub_rn(th, p.az), av = __fsub_rn(th, bv);
        const float err = __fadd_rn(__fsub_rn(p.az, av), __fsub_rn(t.delta, bv));   // az + delta = th + err
        const double s0 = fma(p.saz, t.cd, p.caz * t.sd), c0 = fma(p.caz, t.cd, -(p.saz * t.sd));
        const double e = (double)err;
        sn = (float)fma(-e, c0, s0);
        cs = (float)fma(e, s0, c0);
    } else {
        sn = (float)t.sd; cs = (float)t.cd;
    }
    o.zy = __fmul_rn(cs, t.rad);
    o.zx = __fmul_rn(sn, t.rad);
    o.zz = t.z;
    o.rotated = p.rotated != 0;
    if (o.rotated) {
        o.vx = __fmaf_rn(p.r[2], o.zz, __fmaf_rn(p.r[1], o.zy, __fmul_rn(p.r[0], o.zx)));
        o.vy = __fmaf_rn(p.r[5], o.zz, __fmaf_rn(p.r[4], o.zy, __fmul_rn(p.r[3], o.zx)));
        o.vz = __fmaf_rn(p.r[8], o.zz, __fmaf_rn(p.r[7], o.zy, __fmul_rn(p.r[6], o.zx)));
    } else {
        o.vx = -o.zx; o.vy = -o.zy; o.vz = -o.zz;
    }
    const float len = __fsqrt_rn(__fmaf_rn(o.vz, o.vz, __fmaf_rn(o.vy, o.vy, __fmul_rn(o.vx, o.vx))));
    o.len = fmaxf(len, 1e-12f);   // F.normalize: v / max(|v|, eps)
    return o;
}

__device__ __forceinline__ IncPoint load_inc_point(const IncPoint *__restrict__ p) {
    float4 a, b, c, d;
    ldg256(p, a, b);
    ldg256(reinterpret_cast<const char *>(p) + 32, c, d);
    IncPoint q;
    q.r[0] = a.x; q.r[1] = a.y; q.r[2] = a.z; q.r[3] = a.w; q.r[4] = b.x; q.r[5] = b.y; q.r[6] = b.z; q.r[7] = b.w;
    q.r[8] = c.x; q.az = c.y; q.rotated = __float_as_int(c.z); q.pad = 0;
    q.saz = __hiloint2double(__float_as_int(d.y), __float_as_int(d.x));
    q.caz = __hiloint2double(__float_as_int(d.w), __float_as_int(d.z));
    return q;
}
__device__ __forceinline__ IncTab load_inc_tab(const IncTab *__restrict__ t) {
    float4 a, b;
    ldg256(t, a, b);
    IncTab q;
    q.sd = __hiloint2double(__float_as_int(a.y), __float_as_int(a.x));
    q.cd = __hiloint2double(__float_as_int(a.w), __float_as_int(a.z));
    q.delta = b.x; q.z = b.y; q.rad = b.z; q.pad = 0.f;
    return q;
}

__device__ __forceinline__ void load_ray_mem(const TraceArgs &a, int64_t ray, RayCtx &r) {
    r.ox = __ldg(a.rays_o + 3 * ray); r.oy = __ldg(a.rays_o + 3 * ray + 1); r.oz = __ldg(a.rays_o + 3 * ray + 2);
    r.dx = __ldg(a.rays_d + 3 * ray); r.dy = __ldg(a.rays_d + 3 * ray + 1); r.dz = __ldg(a.rays_d + 3 * ray + 2);
}

__device__ __forceinline__ void load_ray(const TraceArgs &a, int64_t ray, RayCtx &r) {
    if (a.cam_W > 0) {
        // scene/cameras.py:87-100 (rays_d_camera @ world_view_transform[:3,:3].T, F.normalize), generated instead of read
        const int v = (int)((unsigned)ray / (unsigned)a.cam_W), u = (int)((unsigned)ray - (unsigned)v * (unsigned)a.cam_W);
        const float cx = __fdiv_rn(__fadd_rn(__fsub_rn((float)u, 0.5f * (float)a.cam_W), 0.5f), a.cam_fx);
        const float cy = __fdiv_rn(__fadd_rn(__fsub_rn((float)v, 0.5f * (float)a.cam_H), 0.5f), a.cam_fy);
        const float wx = __fmaf_rn(a.cam_M[0], cx, __fmaf_rn(a.cam_M[1], cy, a.cam_M[2]));
        const float wy = __fmaf_rn(a.cam_M[3], cx, __fmaf_rn(a.cam_M[4], cy, a.cam_M[5]));
        const float wz = __fmaf_rn(a.cam_M[6], cx, __fmaf_rn(a.cam_M[7], cy, a.cam_M[8]));
        const float len = fmaxf(__fsqrt_rn(__fmaf_rn(wz, wz, __fmaf_rn(wy, wy, __fmul_rn(wx, wx)))), 1e-12f);
        r.dx = __fdiv_rn(wx, len); r.dy = __fdiv_rn(wy, len); r.dz = __fdiv_rn(wz, len);
        r.ox = a.cam_o[0]; r.oy = a.cam_o[1]; r.oz = a.cam_o[2];
        return;
    }
    if (a.gen_pos != nullptr) {
        // gaussian_renderer/__init__.py:376: origin = position + dir * light_t_min, generated instead of read
        // (a call's rays number < 2^31, checked by the launcher: one 32-bit division instead of a 64-bit one)
        const int64_t pt = (int64_t)((unsigned)ray / (unsigned)a.gen_S);
        const int s = (int)((unsigned)ray - (unsigned)pt * (unsigned)a.gen_S);
        const IncidentSample q = incident_sample(load_inc_point(a.gen_pts + pt), load_inc_tab(a.gen_tab + s), a.gen_azim != nullptr);
        r.dx = __fdiv_rn(q.vx, q.len); r.dy = __fdiv_rn(q.vy, q.len); r.dz = __fdiv_rn(q.vz, q.len);
        r.ox = __fadd_rn(__ldg(a.gen_pos + 3 * pt), __fmul_rn(r.dx, a.gen_tmin));
        r.oy = __fadd_rn(__ldg(a.gen_pos + 3 * pt + 1), __fmul_rn(r.dy, a.gen_tmin));
        r.oz = __fadd_rn(__ldg(a.gen_pos + 3 * pt + 2), __fmul_rn(r.dz, a.gen_tmin));
        return;
    }
    load_ray_mem(a, ray, r);
}

}  // namespace irgs
