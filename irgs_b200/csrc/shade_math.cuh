// Per-sample arithmetic of the shading epilogue (SURVEY.md 8f rank 1 + rank 3): environment-light lookup, GGX specular
// term, rendering-equation integrand, and their hand-derived backward.  Plain float math without device intrinsics, so
// the very same functions compile for the host: tests/shade_host.cpp runs them on the CPU of the build container against
// the golden vectors of the unmodified reference (tests/golden/ref_shading.npz) before any GPU time is spent.
//
// Reference being restated (files under /root/reference):
//   gaussian_renderer/__init__.py:334-415  rendering_equation (the diffuse_sample_num > 0, light_sample_num == 0, non-relight path)
//   gaussian_renderer/__init__.py:417-457  GGX_specular
//   scene/light.py:287-297,315            EnvLight.__call__(mode='pure_env'): lat-long uv, dr.texture(linear), activation, clamp_min(0)
//   scene/gaussian_model.py:748-752       GaussianModel.trace: colour / alpha normalisation of saturated rays
//   nvdiffrast 0.3.x texture.cu indexTextureLinear / TextureFwdKernel / TextureGradKernel (filter 'linear',
//     boundary 'wrap'; the package is not in this image -- its published bilinear scheme is restated: texel centres at
//     (i + 0.5) / size, wrap-around neighbours, uv gradient = finite difference of the four taps times the texture size)
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define IRGS_HD __host__ __device__ __forceinline__
#else
#define IRGS_HD inline
#endif

namespace irgs {

constexpr float SH_PI = 3.14159265358979323846f;
constexpr float SH_TWO_PI = 6.28318530717958647692f;
constexpr float SH_LN2 = 0.69314718055994530942f;
constexpr float SH_FRESNEL = 0.04f;   // rendering_equation calls GGX_specular(..., fresnel=0.04), __init__.py:393

enum { ENV_ACT_NONE = 0, ENV_ACT_EXP = 1, ENV_ACT_SIGMOID = 2 };

struct EnvMap {
    const float *base;   // [H, W, 3] pre-activation texels (EnvLight.base)
    int H, W;
    int activation;      // ENV_ACT_*
    int has_transform;
    float T[9];          // EnvLight.transform, row-major: l = d @ T^T
};

// ------------------------------------------------------------------------------------------------ environment lookup
struct EnvTap {
    int i00, i10, i01, i11;   // texel indices (u0,v0) (u1,v0) (u0,v1) (u1,v1)
    float fu, fv;             // bilinear weights
    float lx, ly, lz;         // direction in the map's frame
    float u, v;               // lat-long coordinates in [0, 1] (before the wrap of the bilinear lookup)
    bool u_open, v_open, y_open;   // the clamps that were NOT active (gradient passes)
};

IRGS_HD float sh_clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

IRGS_HD void env_tap(const EnvMap &e, float dx, float dy, float dz, EnvTap &t) {
    if (e.has_transform) {
        t.lx = e.T[0] * dx + e.T[1] * dy + e.T[2] * dz;
        t.ly = e.T[3] * dx + e.T[4] * dy + e.T[5] * dz;
        t.lz = e.T[6] * dx + e.T[7] * dy + e.T[8] * dz;
    } else {
        t.lx = dx; t.ly = dy; t.lz = dz;
    }
    // light.py:291-294
    float u = atan2f(t.lx, -t.lz);
    if (u != u) u = 0.f;                                   // nan_to_num
    u = u / SH_TWO_PI + 0.5f;
    const float yc = sh_clampf(t.ly, -1.0f + 1e-6f, 1.0f - 1e-6f);
    t.y_open = (t.ly >= -1.0f + 1e-6f) && (t.ly <= 1.0f - 1e-6f);
    float v = acosf(yc) / SH_PI;
    t.u_open = (u >= 0.f) && (u <= 1.f);
    t.v_open = (v >= 0.f) && (v <= 1.f);
    u = sh_clampf(u, 0.f, 1.f);
    v = sh_clampf(v, 0.f, 1.f);
    t.u = u; t.v = v;
    // nvdiffrast indexTextureLinear, boundary mode 'wrap'
    u = u - floorf(u);
    v = v - floorf(v);
    u = u * (float)e.W - 0.5f;
    v = v * (float)e.H - 0.5f;
    int iu0 = (int)floorf(u), iv0 = (int)floorf(v);
    int iu1 = iu0 + 1, iv1 = iv0 + 1;
    t.fu = u - (float)iu0;
    t.fv = v - (float)iv0;
    if (iu0 < 0) iu0 += e.W;
    if (iv0 < 0) iv0 += e.H;
    if (iu1 >= e.W) iu1 -= e.W;
    if (iv1 >= e.H) iv1 -= e.H;
    t.i00 = iu0 + e.W * iv0; t.i10 = iu1 + e.W * iv0;
    t.i01 = iu0 + e.W * iv1; t.i11 = iu1 + e.W * iv1;
}

IRGS_HD float env_bilerp(float a00, float a10, float a01, float a11, float fu, float fv) {
    const float top = a00 + (a10 - a00) * fu, bot = a01 + (a11 - a01) * fu;
    return top + (bot - top) * fv;
}

// raw[c] = filtered pre-activation texel, env[c] = clamp_min(activation(raw), 0)
IRGS_HD void env_fetch(const EnvMap &e, const EnvTap &t, float raw[3], float env[3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float r = env_bilerp(e.base[3 * t.i00 + c], e.base[3 * t.i10 + c], e.base[3 * t.i01 + c], e.base[3 * t.i11 + c],
                                   t.fu, t.fv);
        raw[c] = r;
        float y = r;
        if (e.activation == ENV_ACT_EXP) y = expf(r);
        else if (e.activation == ENV_ACT_SIGMOID) y = 1.0f / (1.0f + expf(-r));
        env[c] = fmaxf(y, 0.f);
    }
}

#ifdef __CUDA_ARCH__
#define IRGS_ENV_ADD(p, v) atomicAdd((p), (v))
#else
#define IRGS_ENV_ADD(p, v) (*(p) += (v))
#endif

// Backward of env_fetch + env_tap: g_env = dL/denv[3].  Adds the texel gradients into grad_base (may be null) and returns
// dL/d(direction) in gd[3] (added).
IRGS_HD void env_backward(const EnvMap &e, const EnvTap &t, const float raw[3], const float /*env*/[3], const float g_env[3],
                          float *grad_base, float gd[3]) {
    float g_u = 0.f, g_v = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float y = raw[c], dy = 1.0f;                                 // activation and its derivative
        if (e.activation == ENV_ACT_EXP) { y = expf(raw[c]); dy = y; }
        else if (e.activation == ENV_ACT_SIGMOID) { y = 1.0f / (1.0f + expf(-raw[c])); dy = y * (1.0f - y); }
        const float g = (y >= 0.f) ? g_env[c] * dy : 0.f;           // clamp_min(0): gradient where activation >= 0
        if (g == 0.f) continue;
        const float a00 = e.base[3 * t.i00 + c], a10 = e.base[3 * t.i10 + c], a01 = e.base[3 * t.i01 + c],
                    a11 = e.base[3 * t.i11 + c];
        if (grad_base) {
            IRGS_ENV_ADD(grad_base + 3 * t.i00 + c, g * (1.0f - t.fu) * (1.0f - t.fv));
            IRGS_ENV_ADD(grad_base + 3 * t.i10 + c, g * t.fu * (1.0f - t.fv));
            IRGS_ENV_ADD(grad_base + 3 * t.i01 + c, g * (1.0f - t.fu) * t.fv);
            IRGS_ENV_ADD(grad_base + 3 * t.i11 + c, g * t.fu * t.fv);
        }
        const float ad = a11 + a00 - a10 - a01;
        g_u += g * ((a10 - a00) + t.fv * ad);
        g_v += g * ((a01 - a00) + t.fu * ad);
    }
    g_u *= (float)e.W;
    g_v *= (float)e.H;
    if (!t.u_open) g_u = 0.f;
    if (!t.v_open) g_v = 0.f;
    // u = atan2(lx, -lz) / 2 pi + 0.5 ;  v = acos(clamp(ly)) / pi
    const float q = t.lx * t.lx + t.lz * t.lz;
    float glx = 0.f, gly = 0.f, glz = 0.f;
    if (q > 0.f) {
        glx = g_u * (-t.lz / q) / SH_TWO_PI;
        glz = g_u * (t.lx / q) / SH_TWO_PI;
    }
    if (t.y_open) gly = -g_v / (sqrtf(fmaxf(1.0f - t.ly * t.ly, 1e-30f)) * SH_PI);
    if (e.has_transform) {
        gd[0] += e.T[0] * glx + e.T[3] * gly + e.T[6] * glz;
        gd[1] += e.T[1] * glx + e.T[4] * gly + e.T[7] * glz;
        gd[2] += e.T[2] * glx + e.T[5] * gly + e.T[8] * glz;
    } else {
        gd[0] += glx; gd[1] += gly; gd[2] += glz;
    }
}

// ------------------------------------------------------------------------------------------------ light-importance sampling
// The light_sample_num > 0 branch of rendering_equation (gaussian_renderer/__init__.py:340-357) mixes the Fibonacci samples
// with directions drawn from the environment map (EnvLight.sample_light_directions, scene/light.py:181-205) and weights
// every sample -- of either kind -- by 1 / clamp_min(p_diffuse / (2 pi) + p_light * light_pdf(dir), 1e-6), light_pdf being
// EnvLight.light_pdf (light.py:207-223): the texel's probability times H W / (2 pi^2 sin(v pi)).
struct MisParams {
    const float *pdf;          // [H, W] texel probabilities (EnvLight._pdf) or null: pure Fibonacci sampling, area = 2 pi
    float p_diffuse, p_light;
};
struct MisTerms { float area, mix, tex, weight, sinv; };

IRGS_HD void mis_area(const EnvMap &e, const MisParams &m, const EnvTap &t, MisTerms &o) {
    if (m.pdf == nullptr) { o.area = SH_TWO_PI; o.mix = 1.0f / SH_TWO_PI; o.tex = o.weight = 0.f; o.sinv = 1.f; return; }
    int ui = (int)(t.u * (float)e.W), vi = (int)(t.v * (float)e.H);      // .long(): truncation, then clamp(0, size - 1)
    ui = ui < 0 ? 0 : (ui > e.W - 1 ? e.W - 1 : ui);
    vi = vi < 0 ? 0 : (vi > e.H - 1 ? e.H - 1 : vi);
    o.tex = m.pdf[ui + vi * e.W];
    o.sinv = sinf(t.v * SH_PI);
    o.weight = (float)e.H * (float)e.W / (2.0f * SH_PI * SH_PI * fmaxf(o.sinv, 1e-6f));
    o.mix = m.p_diffuse / SH_TWO_PI + o.tex * o.weight * m.p_light;
    o.area = 1.0f / fmaxf(o.mix, 1e-6f);
}

// g_area = dL/d area of this sample; adds dL/d direction (through the 1 / sin(v pi) of the weight) to gd[3]
IRGS_HD void mis_backward(const EnvMap &e, const MisParams &m, const EnvTap &t, const MisTerms &o, float g_area, float gd[3]) {
    if (m.pdf == nullptr || g_area == 0.f) return;
    if (!(o.mix >= 1e-6f)) return;                                       // clamp_min(1e-6)
    const float g_mix = -g_area * o.area * o.area;
    const float g_weight = g_mix * m.p_light * o.tex;
    if (!(o.sinv >= 1e-6f) || !t.y_open) return;                         // clamp_min(1e-6) of the sine; clamp of l.y
    const float g_sin = -g_weight * o.weight / o.sinv;
    const float g_v = g_sin * cosf(t.v * SH_PI) * SH_PI;
    const float gly = -g_v / (sqrtf(fmaxf(1.0f - t.ly * t.ly, 1e-30f)) * SH_PI);   // v = acos(l.y) / pi
    if (e.has_transform) { gd[0] += e.T[3] * gly; gd[1] += e.T[4] * gly; gd[2] += e.T[5] * gly; }
    else gd[1] += gly;
}

// ------------------------------------------------------------------------------------------------ shading point
struct ShadePoint {
    float n[3];        // shading normal as given (n_d_i uses it as is, __init__.py:391)
    float N[3];        // GGX: normalize(n) * sign(V . normalize(n))   (__init__.py:427-431)
    float V[3];        // normalize(viewdir)
    float n_len, v_len, sgn;
    float NoV_raw, NoV;
    float r, a2, k, nom1;
    float fd[3];       // base_color / pi
};

IRGS_HD void shade_point_setup(const float n[3], const float view[3], float rough, const float base[3], ShadePoint &p) {
    p.n_len = fmaxf(sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1e-12f);
    p.v_len = fmaxf(sqrtf(view[0] * view[0] + view[1] * view[1] + view[2] * view[2]), 1e-12f);
    float Nn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { p.n[j] = n[j]; Nn[j] = n[j] / p.n_len; p.V[j] = view[j] / p.v_len; }
    const float nov0 = p.V[0] * Nn[0] + p.V[1] * Nn[1] + p.V[2] * Nn[2];
    p.sgn = (nov0 > 0.f) ? 1.0f : ((nov0 < 0.f) ? -1.0f : 0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) p.N[j] = Nn[j] * p.sgn;
    p.NoV_raw = p.N[0] * p.V[0] + p.N[1] * p.V[1] + p.N[2] * p.V[2];
    p.NoV = sh_clampf(p.NoV_raw, 1e-6f, 1.0f);
    p.r = rough;
    const float a = rough * rough;
    p.a2 = a * a;
    p.k = (a + 2.0f * rough + 1.0f) / 8.0f;
    p.nom1 = p.NoV * (1.0f - p.k) + p.k;
#pragma unroll
    for (int j = 0; j < 3; ++j) p.fd[j] = base[j] / SH_PI;
}

struct GgxTerms {   // per-sample intermediates of GGX_specular kept for the backward
    float L[3], H[3], d_len, h_len;
    float NoL_raw, NoH_raw, VoH_raw, NoL, NoH, VoH;
    float pow2, frac0, nom0, nom2, nom_raw, nom, fs;
};

IRGS_HD void ggx_forward(const ShadePoint &p, const float d[3], GgxTerms &g) {
    g.d_len = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-12f);
    float h[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { g.L[j] = d[j] / g.d_len; h[j] = (g.L[j] + p.V[j]) / 2.0f; }
    g.h_len = fmaxf(sqrtf(h[0] * h[0] + h[1] * h[1] + h[2] * h[2]), 1e-12f);
#pragma unroll
    for (int j = 0; j < 3; ++j) g.H[j] = h[j] / g.h_len;
    g.NoL_raw = p.N[0] * g.L[0] + p.N[1] * g.L[1] + p.N[2] * g.L[2];
    g.NoH_raw = p.N[0] * g.H[0] + p.N[1] * g.H[1] + p.N[2] * g.H[2];
    g.VoH_raw = p.V[0] * g.H[0] + p.V[1] * g.H[1] + p.V[2] * g.H[2];
    g.NoL = sh_clampf(g.NoL_raw, 1e-6f, 1.0f);
    g.NoH = sh_clampf(g.NoH_raw, 1e-6f, 1.0f);
    g.VoH = sh_clampf(g.VoH_raw, 1e-6f, 1.0f);
    const float fmi = (-5.55473f * g.VoH - 6.98316f) * g.VoH;
    g.pow2 = exp2f(fmi);
    g.frac0 = SH_FRESNEL + (1.0f - SH_FRESNEL) * g.pow2;
    const float frac = g.frac0 * p.a2;
    g.nom0 = g.NoH * g.NoH * (p.a2 - 1.0f) + 1.0f;
    g.nom2 = g.NoL * (1.0f - p.k) + p.k;
    g.nom_raw = 4.0f * SH_PI * g.nom0 * g.nom0 * p.nom1 * g.nom2;
    g.nom = sh_clampf(g.nom_raw, 1e-6f, 4.0f * SH_PI);
    g.fs = frac / g.nom;
}

IRGS_HD bool sh_open(float x, float lo, float hi) { return x >= lo && x <= hi; }

// What the samples of one shading point accumulate in the backward (reduced over the warp, then shade_point_finish).
struct ShadeAcc {
    float g_base[3];   // dL/dbase_color
    float g_n[3];      // dL/dnormal through n_d_i
    float g_N[3];      // dL/dN (GGX normal)
    float g_V[3];      // dL/dV
    float g_nom1, g_a2, g_k;
};
IRGS_HD void shade_acc_zero(ShadeAcc &a) {
#pragma unroll
    for (int j = 0; j < 3; ++j) a.g_base[j] = a.g_n[j] = a.g_N[j] = a.g_V[j] = 0.f;
    a.g_nom1 = a.g_a2 = a.g_k = 0.f;
}

// g_fs = dL/df_s of this sample; adds to the point accumulators and to gd[3] (dL/d direction).
IRGS_HD void ggx_backward(const ShadePoint &p, const GgxTerms &g, float g_fs, ShadeAcc &acc, float gd[3]) {
    if (g_fs == 0.f) return;
    const float frac = g.frac0 * p.a2;
    const float g_frac = g_fs / g.nom;
    const float g_nom = sh_open(g.nom_raw, 1e-6f, 4.0f * SH_PI) ? -g_fs * frac / (g.nom * g.nom) : 0.f;
    const float c4 = 4.0f * SH_PI;
    const float g_nom0 = g_nom * c4 * 2.0f * g.nom0 * p.nom1 * g.nom2;
    acc.g_nom1 += g_nom * c4 * g.nom0 * g.nom0 * g.nom2;
    const float g_nom2 = g_nom * c4 * g.nom0 * g.nom0 * p.nom1;
    float g_NoH = g_nom0 * 2.0f * g.NoH * (p.a2 - 1.0f);
    acc.g_a2 += g_nom0 * g.NoH * g.NoH + g_frac * g.frac0;
    float g_NoL = g_nom2 * (1.0f - p.k);
    acc.g_k += g_nom2 * (1.0f - g.NoL);
    const float g_frac0 = g_frac * p.a2;
    const float g_fmi = g_frac0 * (1.0f - SH_FRESNEL) * SH_LN2 * g.pow2;
    float g_VoH = g_fmi * (-2.0f * 5.55473f * g.VoH - 6.98316f);
    if (!sh_open(g.NoL_raw, 1e-6f, 1.0f)) g_NoL = 0.f;
    if (!sh_open(g.NoH_raw, 1e-6f, 1.0f)) g_NoH = 0.f;
    if (!sh_open(g.VoH_raw, 1e-6f, 1.0f)) g_VoH = 0.f;
    float gL[3], gH[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        acc.g_N[j] += g_NoL * g.L[j] + g_NoH * g.H[j];
        acc.g_V[j] += g_VoH * g.H[j];
        gL[j] = g_NoL * p.N[j];
        gH[j] = g_NoH * p.N[j] + g_VoH * p.V[j];
    }
    // H = h / |h|, h = (L + V) / 2
    const float hh = g.H[0] * gH[0] + g.H[1] * gH[1] + g.H[2] * gH[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float gh = (gH[j] - g.H[j] * hh) / g.h_len * 0.5f;
        gL[j] += gh;
        acc.g_V[j] += gh;
    }
    // L = d / |d|
    const float ll = g.L[0] * gL[0] + g.L[1] * gL[1] + g.L[2] * gL[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) gd[j] += (gL[j] - g.L[j] * ll) / g.d_len;
}

// Point-level tail of the backward: from the reduced accumulators to dL/d(base_color, roughness, normal, viewdir).
// g_normal receives only the DIRECT dependence (n_d_i and the GGX normal); the dependence through the sampled
// directions is chained by the caller (dL/dR of rotation_between_z, like incident_backward_kernel).
IRGS_HD void shade_point_finish(const ShadePoint &p, const ShadeAcc &a, float g_base[3], float &g_rough, float g_normal[3],
                                float g_view[3]) {
    float gN[3], gV[3];
    // nom1 = NoV (1 - k) + k
    float g_NoV = a.g_nom1 * (1.0f - p.k);
    const float g_k = a.g_k + a.g_nom1 * (1.0f - p.NoV);
    if (!sh_open(p.NoV_raw, 1e-6f, 1.0f)) g_NoV = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) { gN[j] = a.g_N[j] + g_NoV * p.V[j]; gV[j] = a.g_V[j] + g_NoV * p.N[j]; }
    // N = sgn * n / |n|   (sign() has zero gradient)
    float Nn[3], gNn[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { Nn[j] = p.n[j] / p.n_len; gNn[j] = p.sgn * gN[j]; }
    const float nn = Nn[0] * gNn[0] + Nn[1] * gNn[1] + Nn[2] * gNn[2];
    const float vv = p.V[0] * gV[0] + p.V[1] * gV[1] + p.V[2] * gV[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        g_normal[j] = a.g_n[j] + (gNn[j] - Nn[j] * nn) / p.n_len;
        g_view[j] = (gV[j] - p.V[j] * vv) / p.v_len;
        g_base[j] = a.g_base[j] / SH_PI;
    }
    // a2 = r^4, k = (r^2 + 2 r + 1) / 8
    g_rough = a.g_a2 * 4.0f * p.r * p.r * p.r + g_k * (2.0f * p.r + 2.0f) / 8.0f;
}

// ------------------------------------------------------------------------------------------------ one incident sample
// Traced radiance of the ray as GaussianModel.trace hands it to rendering_equation (gaussian_model.py:748-752): rays whose
// accumulated alpha reached 1 - transmittance_min are normalised (colour / alpha, alpha = 1).  sat < 0 disables this.
IRGS_HD void trace_normalise(const float c_raw[3], float a_raw, float sat, float c[3], float &a, bool &saturated) {
    saturated = (sat >= 0.f) && !(a_raw < sat);
    if (saturated) {
        c[0] = c_raw[0] / a_raw; c[1] = c_raw[1] / a_raw; c[2] = c_raw[2] / a_raw;
        a = 1.0f;
    } else {
        c[0] = c_raw[0]; c[1] = c_raw[1]; c[2] = c_raw[2];
        a = a_raw;
    }
}

struct ShadeSample {
    float env[3];        // global_incident_lights
    float local[3];      // local_incident_lights (normalised traced colour)
    float vis;           // incident_visibility = 1 - alpha
    float Li[3];         // incident_lights
    float ndi;
    float fs;
    float transport[3];  // Li * area * n_d_i, area = 2 pi (graphics_utils.py:43) or the mixed-sampling weight (mis_area)
};

// MIX = false: pure Fibonacci sampling (area 2 pi), the mixed-sampling code is compiled out (it cost the default path 19 %)
template <bool MIX>
IRGS_HD void shade_sample_forward(const ShadePoint &p, const EnvMap &e, const MisParams &m, const float d[3],
                                  const float c_raw[3], float a_raw, float sat, ShadeSample &o) {
    EnvTap t;
    float raw[3], a;
    bool saturated;
    MisTerms mt;
    env_tap(e, d[0], d[1], d[2], t);
    env_fetch(e, t, raw, o.env);
    if (MIX) mis_area(e, m, t, mt);
    else mt.area = SH_TWO_PI;
    trace_normalise(c_raw, a_raw, sat, o.local, a, saturated);
    o.vis = 1.0f - a;
    GgxTerms g;
    ggx_forward(p, d, g);
    o.fs = g.fs;
    o.ndi = fmaxf(p.n[0] * d[0] + p.n[1] * d[1] + p.n[2] * d[2], 0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        o.Li[c] = o.vis * o.env[c] + o.local[c];
        o.transport[c] = o.Li[c] * mt.area * o.ndi;
    }
}

// gD, gS, gE = dL/d(diffuse, specular, light_direct) of the point ALREADY divided by the sample count (the mean over S);
// gVis, gLi, gLocal likewise for the evaluation-mode outputs (visibility, light, light_indirect), zero in training.
// Outputs: g_c_raw[3], g_a_raw (gradients of the RAW traced colour / alpha of this ray), gd[3] (dL/d direction, overwritten);
// texel gradients are added into grad_base; point-level terms into acc.
template <bool MIX>
IRGS_HD void shade_sample_backward(const ShadePoint &p, const EnvMap &e, const MisParams &m, const float d[3],
                                   const float c_raw[3], float a_raw, float sat, const float gD[3], const float gS[3],
                                   const float gE[3], float gVis,
                                   const float gLi[3], const float gLocal[3], float *grad_base, ShadeAcc &acc,
                                   float g_c_raw[3], float &g_a_raw, float gd[3]) {
    EnvTap t;
    float raw[3], env[3], local[3], a;
    bool saturated;
    MisTerms mt;
    env_tap(e, d[0], d[1], d[2], t);
    env_fetch(e, t, raw, env);
    if (MIX) mis_area(e, m, t, mt);
    else mt.area = SH_TWO_PI;
    trace_normalise(c_raw, a_raw, sat, local, a, saturated);
    const float vis = 1.0f - a;
    GgxTerms g;
    ggx_forward(p, d, g);
    const float ndi_raw = p.n[0] * d[0] + p.n[1] * d[1] + p.n[2] * d[2];
    const float ndi = fmaxf(ndi_raw, 0.f);
    float g_fs = 0.f, g_ndi = 0.f, g_area = 0.f, g_vis = gVis, g_env[3], g_local[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float Li = vis * env[c] + local[c];
        const float transport = Li * mt.area * ndi;
        const float g_tr = p.fd[c] * gD[c] + g.fs * gS[c];
        acc.g_base[c] += transport * gD[c];
        g_fs += transport * gS[c];
        const float g_Li = g_tr * mt.area * ndi + gLi[c];
        g_ndi += g_tr * Li * mt.area;
        if (MIX) g_area += g_tr * Li * ndi;
        g_env[c] = g_Li * vis + gE[c];
        g_vis += g_Li * env[c];
        g_local[c] = g_Li + gLocal[c];
    }
    gd[0] = gd[1] = gd[2] = 0.f;
    if (ndi_raw >= 0.f) {   // clamp(min=0)
#pragma unroll
        for (int j = 0; j < 3; ++j) { acc.g_n[j] += g_ndi * d[j]; gd[j] += g_ndi * p.n[j]; }
    }
    ggx_backward(p, g, g_fs, acc, gd);
    env_backward(e, t, raw, env, g_env, grad_base, gd);
    if (MIX) mis_backward(e, m, t, mt, g_area, gd);
    // through GaussianModel.trace's normalisation
    const float g_a = -g_vis;
    if (saturated) {
        g_a_raw = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g_c_raw[c] = g_local[c] / a_raw;
            g_a_raw -= g_local[c] * c_raw[c] / (a_raw * a_raw);
        }
    } else {
        g_a_raw = g_a;
        g_c_raw[0] = g_local[0]; g_c_raw[1] = g_local[1]; g_c_raw[2] = g_local[2];
    }
}

}  // namespace irgs
