// Shading epilogue around the incident-ray trace (SURVEY.md 8f rank 1 + rank 3), sm_100a.
//
// The reference evaluates its rendering equation (gaussian_renderer/__init__.py:334-415) with some thirty element-wise
// torch kernels over [P, S, 3] tensors: incident directions, environment lookup (nvdiffrast), visibility blend, GGX
// term, two means over S.  Here ONE kernel per direction does all of it: a warp owns a shading point, regenerates the S
// Fibonacci directions with the device function the tracing kernels use (incident_sample), reads 16 bytes per ray (traced
// colour + alpha), looks the environment map up (lat-long bilinear, L2 resident), and reduces over S with shuffles.
// The backward recomputes the same quantities, writes dL/d(colour, alpha) per ray for the tracer's backward, adds the
// texel gradients with atomics, and chains dL/d(direction) through rotation_between_z to the shading normal.
// Per-sample arithmetic: shade_math.cuh (shared with the host harness of the tests).
//
// Roofline: HBM streaming -- forward reads 16 B per ray, backward reads 16 B and writes 16 B per ray; everything per
// point (56 B in, 64 B out) is amortised over S.  At S = 256 that is ~16.5 / ~32.5 bytes per ray.
#include "internal.cuh"
#include "shade_math.cuh"
#include "trace_common.cuh"

namespace irgs {

struct ShadeArgs {
    const float *normals, *azimuth;        // [P,3], [P] or null
    int64_t n_points;
    int S;
    const float *base_color, *roughness, *viewdirs;   // [P,3], [P], [P,3]
    const float *trace_color, *trace_alpha;           // [P*S,3], [P*S] raw tracer outputs
    float saturate;                                   // 1 - transmittance_min, < 0: no normalisation
    EnvMap env;
    // light_sample_num > 0 (irgs_shade_sampling_t): explicit directions instead of generated ones, mixed-sampling weights
    const IncTab *tab;                                // [S] per-sample table of the generated directions (incident_table)
    const float *dirs;                                // [P*S,3] or null (generated from normals / azimuth)
    MisParams mis;
    float inv_count;                                  // 1 / (samples the means run over)
};

// direction of sample s of point pt: generated (q is filled, q.rotated tells whether it depends on the normal) or read
template <bool MIX>
__device__ __forceinline__ void sample_dir(const ShadeArgs &a, const IncPoint &ip, int64_t ray, int s, IncidentSample &q,
                                           float d[3]) {
    if (MIX && a.dirs != nullptr) {
        d[0] = __ldg(a.dirs + 3 * ray); d[1] = __ldg(a.dirs + 3 * ray + 1); d[2] = __ldg(a.dirs + 3 * ray + 2);
        q.rotated = false; q.len = 1.f; q.zx = q.zy = q.zz = q.vx = q.vy = q.vz = 0.f;
        return;
    }
    q = incident_sample(ip, load_inc_tab(a.tab + s), a.azimuth != nullptr);
    d[0] = __fdiv_rn(q.vx, q.len); d[1] = __fdiv_rn(q.vy, q.len); d[2] = __fdiv_rn(q.vz, q.len);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void load_point(const ShadeArgs &a, int64_t pt, float n[3], ShadePoint &p) {
    float view[3], base[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        n[j] = __ldg(a.normals + 3 * pt + j);
        view[j] = __ldg(a.viewdirs + 3 * pt + j);
        base[j] = __ldg(a.base_color + 3 * pt + j);
    }
    shade_point_setup(n, view, __ldg(a.roughness + pt), base, p);
}

// out [P,16]: 0-2 diffuse, 3-5 specular, 6-8 light_direct, 9 visibility, 10-12 light, 13-15 light_indirect (means over S)
template <bool MIX>   // MIX: light_sample_num > 0 (explicit directions and / or mixed-sampling weights)
__global__ void __launch_bounds__(128) shade_forward_kernel(ShadeArgs a, float *__restrict__ out) {
    const int64_t pt = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pt >= a.n_points) return;
    float n[3];
    ShadePoint p;
    load_point(a, pt, n, p);
    const IncPoint ip = incident_point(n[0], n[1], n[2], a.azimuth != nullptr, a.azimuth ? __ldg(a.azimuth + pt) : 0.f);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    for (int s = lane; s < a.S; s += 32) {
        const int64_t ray = pt * a.S + s;
        IncidentSample q;
        float d[3];
        sample_dir<MIX>(a, ip, ray, s, q, d);
        const float c_raw[3] = {__ldg(a.trace_color + 3 * ray), __ldg(a.trace_color + 3 * ray + 1),
                                __ldg(a.trace_color + 3 * ray + 2)};
        ShadeSample o;
        shade_sample_forward<MIX>(p, a.env, a.mis, d, c_raw, __ldg(a.trace_alpha + ray), a.saturate, o);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            acc[c] += p.fd[c] * o.transport[c];
            acc[3 + c] += o.fs * o.transport[c];
            acc[6 + c] += o.env[c];
            acc[10 + c] += o.Li[c];
            acc[13 + c] += o.local[c];
        }
        acc[9] += o.vis;
    }
    const float inv = a.inv_count;
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float v = warp_sum(acc[j]);
        if (lane == j) mine = v * inv;
    }
    if (lane < 16) out[16 * pt + lane] = mine;    // one 64-byte row per point
}

// g_out [P,16] in the layout of the forward's rows.  Per ray: g_color [P*S,3], g_alpha [P*S] (overwritten).  Per point:
// g_point [P,16]: 0-2 dL/dbase_color, 3 dL/droughness, 4-6 dL/dnormal, 7-9 dL/dviewdirs (overwritten).  Texel gradients
// are ADDED into grad_env [H,W,3] (may be null).
#ifndef IRGS_SHADE_BWD_BLOCKS
#define IRGS_SHADE_BWD_BLOCKS 4   // resident blocks per SM the register budget is set for.  Measured forward+backward per 2^22 rays:
                                  // no cap (135 registers, 3 blocks) 0.668 ms | 4 (128 registers, no spills) 0.614 | 5 (96 registers, spills) 0.634
#endif
template <bool MIX>
__global__ void __launch_bounds__(128, IRGS_SHADE_BWD_BLOCKS) shade_backward_kernel(ShadeArgs a, const float *__restrict__ g_out,
                                                             float *__restrict__ g_color, float *__restrict__ g_alpha,
                                                             float *__restrict__ g_point, float *__restrict__ grad_env) {
    const int64_t pt = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pt >= a.n_points) return;
    float n[3];
    ShadePoint p;
    load_point(a, pt, n, p);
    const IncPoint ip = incident_point(n[0], n[1], n[2], a.azimuth != nullptr, a.azimuth ? __ldg(a.azimuth + pt) : 0.f);
    const float inv = a.inv_count;
    float go = (lane < 16) ? __ldg(g_out + 16 * pt + lane) * inv : 0.f;     // the mean over S folded into the gradients
    float gD[3], gS[3], gE[3], gLi[3], gLocal[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        gD[c] = __shfl_sync(0xffffffffu, go, c);
        gS[c] = __shfl_sync(0xffffffffu, go, 3 + c);
        gE[c] = __shfl_sync(0xffffffffu, go, 6 + c);
        gLi[c] = __shfl_sync(0xffffffffu, go, 10 + c);
        gLocal[c] = __shfl_sync(0xffffffffu, go, 13 + c);
    }
    const float gVis = __shfl_sync(0xffffffffu, go, 9);
    ShadeAcc acc;
    shade_acc_zero(acc);
    float G[9];     // dL/dR of rotation_between_z, row-major
#pragma unroll
    for (int j = 0; j < 9; ++j) G[j] = 0.f;
    for (int s = lane; s < a.S; s += 32) {
        const int64_t ray = pt * a.S + s;
        IncidentSample q;
        float d[3];
        sample_dir<MIX>(a, ip, ray, s, q, d);
        const float c_raw[3] = {__ldg(a.trace_color + 3 * ray), __ldg(a.trace_color + 3 * ray + 1),
                                __ldg(a.trace_color + 3 * ray + 2)};
        float g_c[3], g_a, gd[3];
        shade_sample_backward<MIX>(p, a.env, a.mis, d, c_raw, __ldg(a.trace_alpha + ray), a.saturate, gD, gS, gE, gVis, gLi, gLocal,
                              grad_env, acc, g_c, g_a, gd);
        g_color[3 * ray] = g_c[0]; g_color[3 * ray + 1] = g_c[1]; g_color[3 * ray + 2] = g_c[2];
        g_alpha[ray] = g_a;
        if (q.rotated && q.len > 1e-12f) {   // d = v / |v|, v = R zs  (same chain as incident_backward_kernel, trace.cu)
            const float dd = d[0] * gd[0] + d[1] * gd[1] + d[2] * gd[2];
            const float gvx = (gd[0] - d[0] * dd) / q.len, gvy = (gd[1] - d[1] * dd) / q.len, gvz = (gd[2] - d[2] * dd) / q.len;
            G[0] += gvx * q.zx; G[1] += gvx * q.zy; G[2] += gvx * q.zz;
            G[3] += gvy * q.zx; G[4] += gvy * q.zy; G[5] += gvy * q.zz;
            G[6] += gvz * q.zx; G[7] += gvz * q.zy; G[8] += gvz * q.zz;
        }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        acc.g_base[j] = warp_sum(acc.g_base[j]);
        acc.g_n[j] = warp_sum(acc.g_n[j]);
        acc.g_N[j] = warp_sum(acc.g_N[j]);
        acc.g_V[j] = warp_sum(acc.g_V[j]);
    }
    acc.g_nom1 = warp_sum(acc.g_nom1);
    acc.g_a2 = warp_sum(acc.g_a2);
    acc.g_k = warp_sum(acc.g_k);
#pragma unroll
    for (int j = 0; j < 9; ++j) G[j] = warp_sum(G[j]);
    if (lane == 0) {
        float g_base[3], g_rough, g_normal[3], g_view[3];
        shade_point_finish(p, acc, g_base, g_rough, g_normal, g_view);
        // dL/dn through R(n): v1 = -n.y, v2 = n.x, c = max(n.z + 1, 1e-7)   (graphics_utils.py:133-165)
        const float v1 = -n[1], v2 = n[0], c = fmaxf(n[2] + 1.0f, 1e-7f);
        const float gv1 = (G[1] + G[3]) * v2 / c - (G[4] + G[8]) * 2.0f * v1 / c - G[5] + G[7];
        const float gv2 = -(G[0] + G[8]) * 2.0f * v2 / c + (G[1] + G[3]) * v1 / c + G[2] - G[6];
        const float gc = (G[0] * v2 * v2 - (G[1] + G[3]) * v1 * v2 + G[4] * v1 * v1 + G[8] * (v1 * v1 + v2 * v2)) / (c * c);
        g_normal[0] += gv2;
        g_normal[1] += -gv1;
        g_normal[2] += (n[2] + 1.0f > 1e-7f) ? gc : 0.0f;
        float *o = g_point + 16 * pt;
        o[0] = g_base[0]; o[1] = g_base[1]; o[2] = g_base[2]; o[3] = g_rough;
        o[4] = g_normal[0]; o[5] = g_normal[1]; o[6] = g_normal[2];
        o[7] = g_view[0]; o[8] = g_view[1]; o[9] = g_view[2];
        o[10] = o[11] = o[12] = o[13] = o[14] = o[15] = 0.f;
    }
}

// ---- stand-alone environment lookup (EnvLight.__call__(dirs, mode='pure_env'), light.py:287-297,315): one thread per direction
__global__ void env_lookup_forward_kernel(EnvMap env, const float *__restrict__ dirs, int64_t n, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvTap t;
    float raw[3], val[3];
    env_tap(env, __ldg(dirs + 3 * i), __ldg(dirs + 3 * i + 1), __ldg(dirs + 3 * i + 2), t);
    env_fetch(env, t, raw, val);
    out[3 * i] = val[0]; out[3 * i + 1] = val[1]; out[3 * i + 2] = val[2];
}

__global__ void env_lookup_backward_kernel(EnvMap env, const float *__restrict__ dirs, const float *__restrict__ g_out,
                                           int64_t n, float *__restrict__ g_dirs, float *__restrict__ grad_env) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    EnvTap t;
    float raw[3], val[3], gd[3] = {0.f, 0.f, 0.f};
    const float g[3] = {__ldg(g_out + 3 * i), __ldg(g_out + 3 * i + 1), __ldg(g_out + 3 * i + 2)};
    env_tap(env, __ldg(dirs + 3 * i), __ldg(dirs + 3 * i + 1), __ldg(dirs + 3 * i + 2), t);
    env_fetch(env, t, raw, val);
    env_backward(env, t, raw, val, g, grad_env, gd);
    if (g_dirs) { g_dirs[3 * i] = gd[0]; g_dirs[3 * i + 1] = gd[1]; g_dirs[3 * i + 2] = gd[2]; }
}


// ------------------------------------------------------------------------------------------------ relight branch
// gaussian_renderer/__init__.py:362-381: under novel lighting the radiance arriving from a surface point seen by a secondary
// ray is not the baked SH colour but a split-sum shading of that HIT point: the ray composites the surfels' base colour and
// roughness (S = 4 feature channels) and their normals; the hit point is shaded with the environment's diffuse / specular
// prefilter (the caller's own envmap object: cube-map mips built by nvdiffrec / nvdiffrast, scene/light.py:264-328 -- out of
// scope here) and the FG lookup table.  Two element-wise kernels around those two lookups replace the
// reference's ~25 torch kernels over [P,S,*] tensors:
//   relight_hit_kernel      raw tracer outputs + incident direction -> hit normal, reflected direction, roughness (the
//                           arguments of the two environment lookups) and a packed row for the second kernel
//   relight_combine_kernel  the two environment results + FG table -> local incident radiance and the normalised alpha,
//                           in the form irgs_shade_forward consumes as (trace_color, trace_alpha) with saturate_alpha < 0
__global__ void relight_hit_kernel(int64_t n_rays, const float *__restrict__ dirs, const float *__restrict__ normal,
                                   const float *__restrict__ feature, const float *__restrict__ alpha, float saturate,
                                   float *__restrict__ hit_normal, float *__restrict__ reflected, float *__restrict__ rough_out,
                                   float *__restrict__ pack) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float a = alpha[r];
    const bool sat = saturate >= 0.f && !(a < saturate);     // scene/gaussian_model.py:751-756
    const float an = sat ? 1.0f : a;
    float n[3], f[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) n[k] = sat ? normal[3 * r + k] / a : normal[3 * r + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = (sat ? feature[4 * r + k] / a : feature[4 * r + k]) / fmaxf(an, 1e-6f);   // :367
    const float nl = fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-12f);                              // :368 F.normalize
#pragma unroll
    for (int k = 0; k < 3; ++k) n[k] /= nl;
    const float wi[3] = {-dirs[3 * r], -dirs[3 * r + 1], -dirs[3 * r + 2]};                                      // :371
    const float ndv = n[0] * wi[0] + n[1] * wi[1] + n[2] * wi[2];
    float rf[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) rf[k] = ndv * n[k] * 2.0f - wi[k];                                               // :373
    const float rl = fmaxf(sqrtf(rf[0] * rf[0] + rf[1] * rf[1] + rf[2] * rf[2]), 1e-12f);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        hit_normal[3 * r + k] = n[k];
        reflected[3 * r + k] = rf[k] / rl;
    }
    rough_out[r] = f[3];
    float4 *row = reinterpret_cast<float4 *>(pack + 8 * r);
    row[0] = make_float4(f[0], f[1], f[2], fminf(fmaxf(ndv, 0.f), 1.f));                                          // :374 fg_uv.clamp(0, 1)
    row[1] = make_float4(fminf(fmaxf(f[3], 0.f), 1.f), an, 0.f, 0.f);
}

// nvdiffrast texture(filter 'linear', boundary 'clamp') on the [H, W, 2] FG table, restated like the 'wrap' lookup of
// shade_math.cuh: texel centres at (i + 0.5) / size, neighbour indices clamped to the table.
__global__ void relight_combine_kernel(int64_t n_rays, const float *__restrict__ pack, const float *__restrict__ env_diffuse,
                                       const float *__restrict__ env_specular, const float *__restrict__ fg_lut, int H, int W,
                                       float f0, int wo_indirect, float *__restrict__ local, float *__restrict__ alpha_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float4 p0 = __ldg(reinterpret_cast<const float4 *>(pack + 8 * r)), p1 = __ldg(reinterpret_cast<const float4 *>(pack + 8 * r) + 1);
    const float x = p0.w * (float)W - 0.5f, y = p1.x * (float)H - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float tx = x - fx0, ty = y - fy0;
    const int x0 = min(max((int)fx0, 0), W - 1), x1 = min(max((int)fx0 + 1, 0), W - 1);
    const int y0 = min(max((int)fy0, 0), H - 1), y1 = min(max((int)fy0 + 1, 0), H - 1);
    float fg[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const float a00 = __ldg(fg_lut + ((size_t)y0 * W + x0) * 2 + c), a10 = __ldg(fg_lut + ((size_t)y0 * W + x1) * 2 + c);
        const float a01 = __ldg(fg_lut + ((size_t)y1 * W + x0) * 2 + c), a11 = __ldg(fg_lut + ((size_t)y1 * W + x1) * 2 + c);
        const float top = a00 + (a10 - a00) * tx, bot = a01 + (a11 - a01) * tx;
        fg[c] = top + (bot - top) * ty;
    }
    const float ks = f0 * fg[0] + fg[1];                                                                          // :376
    const float base[3] = {p0.x, p0.y, p0.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float v = (base[k] * env_diffuse[3 * r + k] + env_specular[3 * r + k] * ks) * p1.y;                 // :370,376-377
        local[3 * r + k] = wo_indirect ? 0.f : v;                                                                 // :378-379
    }
    alpha_out[r] = p1.y;
}

static int fail_msg(const char *msg) {
    set_error(msg);
    return 1;
}

static int make_env(const irgs_envmap_t *env, EnvMap &e) {
    if (!env || !env->base) return fail_msg("null environment map");
    if (env->height < 1 || env->width < 1) return fail_msg("environment map must be at least 1 x 1");
    if (env->activation < 0 || env->activation > 2) return fail_msg("environment activation must be 0 (none), 1 (exp) or 2 (sigmoid)");
    e.base = env->base; e.H = env->height; e.W = env->width; e.activation = env->activation;
    e.has_transform = env->has_transform;
    for (int j = 0; j < 9; ++j) e.T[j] = env->transform[j];
    return 0;
}

static int make_args(const irgs_incident_t *gen, const irgs_envmap_t *env, const irgs_shade_sampling_t *smp,
                     const float *base_color, const float *roughness, const float *viewdirs, const float *trace_color,
                     const float *trace_alpha, float saturate_alpha, ShadeArgs &a) {
    if (!gen) return fail_msg("null incident-ray descriptor");
    if (gen->n_points < 0 || gen->sample_num < 1) return fail_msg("incident rays: n_points >= 0 and sample_num >= 1 required");
    if (gen->n_points > 0 && (!gen->normals || !base_color || !roughness || !viewdirs || !trace_color || !trace_alpha))
        return fail_msg("shade: null input array");
    if (make_env(env, a.env)) return 1;
    a.normals = gen->normals; a.azimuth = gen->azimuth; a.n_points = gen->n_points; a.S = gen->sample_num;
    a.base_color = base_color; a.roughness = roughness; a.viewdirs = viewdirs;
    a.trace_color = trace_color; a.trace_alpha = trace_alpha; a.saturate = saturate_alpha;
    a.dirs = nullptr; a.mis.pdf = nullptr; a.mis.p_diffuse = 1.f; a.mis.p_light = 0.f;
    a.inv_count = 1.0f / (float)gen->sample_num;
    a.tab = nullptr;
    if (smp != nullptr) {
        if (smp->total_samples < gen->sample_num) return fail_msg("shade: total_samples must be >= the samples of this call");
        if (smp->pdf != nullptr && !(smp->p_diffuse >= 0.f && smp->p_light >= 0.f))
            return fail_msg("shade: sampling probabilities must be non-negative");
        a.dirs = smp->dirs; a.mis.pdf = smp->pdf; a.mis.p_diffuse = smp->p_diffuse; a.mis.p_light = smp->p_light;
        a.inv_count = 1.0f / (float)smp->total_samples;
    }
    return 0;
}

}  // namespace irgs

using namespace irgs;

extern "C" {

int irgs_shade_forward(const irgs_incident_t *gen, const irgs_envmap_t *env, const irgs_shade_sampling_t *sampling,
                       const float *base_color, const float *roughness, const float *viewdirs, const float *trace_color,
                       const float *trace_alpha, float saturate_alpha, float *out, void *stream) {
    ShadeArgs a;
    if (make_args(gen, env, sampling, base_color, roughness, viewdirs, trace_color, trace_alpha, saturate_alpha, a)) return 1;
    if (a.n_points == 0) return 0;
    if (!out) return fail_msg("shade: null output array");
    if (a.dirs == nullptr && !(a.tab = incident_table(a.S, (cudaStream_t)stream))) return 1;
    const unsigned grid = (unsigned)((a.n_points * 32 + 127) / 128);
    if (sampling != nullptr) shade_forward_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(a, out);
    else shade_forward_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(a, out);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int irgs_shade_backward(const irgs_incident_t *gen, const irgs_envmap_t *env, const irgs_shade_sampling_t *sampling,
                        const float *base_color, const float *roughness, const float *viewdirs, const float *trace_color,
                        const float *trace_alpha, float saturate_alpha, const float *g_out, float *g_trace_color,
                        float *g_trace_alpha, float *g_point, float *grad_env, void *stream) {
    ShadeArgs a;
    if (make_args(gen, env, sampling, base_color, roughness, viewdirs, trace_color, trace_alpha, saturate_alpha, a)) return 1;
    if (a.n_points == 0) return 0;
    if (!g_out || !g_trace_color || !g_trace_alpha || !g_point) return fail_msg("shade backward: null array");
    if (a.dirs == nullptr && !(a.tab = incident_table(a.S, (cudaStream_t)stream))) return 1;
    const unsigned grid = (unsigned)((a.n_points * 32 + 127) / 128);
    if (sampling != nullptr)
        shade_backward_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(a, g_out, g_trace_color, g_trace_alpha, g_point, grad_env);
    else
        shade_backward_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(a, g_out, g_trace_color, g_trace_alpha, g_point, grad_env);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int irgs_env_lookup_forward(const irgs_envmap_t *env, const float *dirs, int64_t n_dirs, float *out, void *stream) {
    EnvMap e;
    if (make_env(env, e)) return 1;
    if (n_dirs < 0) return fail_msg("n_dirs < 0");
    if (n_dirs == 0) return 0;
    if (!dirs || !out) return fail_msg("env lookup: null array");
    env_lookup_forward_kernel<<<(unsigned)((n_dirs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(e, dirs, n_dirs, out);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int irgs_env_lookup_backward(const irgs_envmap_t *env, const float *dirs, const float *g_out, int64_t n_dirs, float *g_dirs,
                             float *grad_env, void *stream) {
    EnvMap e;
    if (make_env(env, e)) return 1;
    if (n_dirs < 0) return fail_msg("n_dirs < 0");
    if (n_dirs == 0) return 0;
    if (!dirs || !g_out) return fail_msg("env lookup backward: null array");
    env_lookup_backward_kernel<<<(unsigned)((n_dirs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(e, dirs, g_out, n_dirs, g_dirs,
                                                                                                   grad_env);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int irgs_relight_hit(int64_t n_rays, const float *dirs, const float *trace_normal, const float *trace_feature,
                     const float *trace_alpha, float saturate_alpha, float *hit_normal, float *reflected, float *roughness,
                     float *pack, void *stream) {
    if (n_rays < 0) return fail_msg("n_rays < 0");
    if (n_rays == 0) return 0;
    if (!dirs || !trace_normal || !trace_feature || !trace_alpha || !hit_normal || !reflected || !roughness || !pack)
        return fail_msg("relight_hit: null array");
    relight_hit_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_rays, dirs, trace_normal, trace_feature, trace_alpha, saturate_alpha, hit_normal, reflected, roughness, pack);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int irgs_relight_combine(int64_t n_rays, const float *pack, const float *env_diffuse, const float *env_specular,
                         const float *fg_lut, int lut_height, int lut_width, float f0, int wo_indirect, float *local_light,
                         float *alpha_out, void *stream) {
    if (n_rays < 0) return fail_msg("n_rays < 0");
    if (n_rays == 0) return 0;
    if (!pack || !env_diffuse || !env_specular || !fg_lut || !local_light || !alpha_out) return fail_msg("relight_combine: null array");
    if (lut_height < 1 || lut_width < 1) return fail_msg("relight_combine: the FG table must be at least 1 x 1");
    relight_combine_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_rays, pack, env_diffuse, env_specular, fg_lut, lut_height, lut_width, f0, wo_indirect ? 1 : 0, local_light, alpha_out);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
