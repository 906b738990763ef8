// Host harness of the shading epilogue's per-sample arithmetic -- TEST INFRASTRUCTURE (built by tests/test_shading_cpu.py
// with g++ into tests/_shade_host.so).  It runs the very functions of irgs_b200/csrc/shade_math.cuh that the CUDA kernels
// of shade.cu call, sequentially on the CPU, so that the hand-derived backward can be checked against the reference's
// golden gradients in the build container (no GPU).  The incident directions are an input here (on the GPU they come
// from incident_sample, which is pinned separately by tests/test_gpu_incident.py).
#include <cstdint>
#include <cstring>

#include "../irgs_b200/csrc/shade_math.cuh"

using namespace irgs;

static EnvMap make_env(const float *base, int H, int W, int activation, const float *transform) {
    EnvMap e;
    e.base = base; e.H = H; e.W = W; e.activation = activation; e.has_transform = transform != nullptr;
    for (int j = 0; j < 9; ++j) e.T[j] = transform ? transform[j] : 0.f;
    return e;
}

extern "C" {

// out [P,16] in the layout of irgs_shade_forward
void shade_host_forward(int64_t P, int S, const float *normals, const float *viewdirs, const float *roughness,
                        const float *base_color, const float *dirs, const float *c_raw, const float *a_raw, float sat,
                        const float *env_base, int H, int W, int activation, const float *transform, const float *pdf,
                        float p_diffuse, float p_light, int total_samples, float *out) {
    const EnvMap e = make_env(env_base, H, W, activation, transform);
    const MisParams m = {pdf, p_diffuse, p_light};
    for (int64_t pt = 0; pt < P; ++pt) {
        ShadePoint p;
        shade_point_setup(normals + 3 * pt, viewdirs + 3 * pt, roughness[pt], base_color + 3 * pt, p);
        float acc[16] = {0};
        for (int s = 0; s < S; ++s) {
            const int64_t ray = pt * S + s;
            ShadeSample o;
            if (pdf) shade_sample_forward<true>(p, e, m, dirs + 3 * ray, c_raw + 3 * ray, a_raw[ray], sat, o);
            else shade_sample_forward<false>(p, e, m, dirs + 3 * ray, c_raw + 3 * ray, a_raw[ray], sat, o);
            for (int c = 0; c < 3; ++c) {
                acc[c] += p.fd[c] * o.transport[c];
                acc[3 + c] += o.fs * o.transport[c];
                acc[6 + c] += o.env[c];
                acc[10 + c] += o.Li[c];
                acc[13 + c] += o.local[c];
            }
            acc[9] += o.vis;
        }
        for (int j = 0; j < 16; ++j) out[16 * pt + j] = acc[j] / (float)total_samples;
    }
}

// g_point [P,16]: 0-2 base_color, 3 roughness, 4-6 normal (DIRECT dependence only), 7-9 viewdirs; g_dirs [P*S,3]
void shade_host_backward(int64_t P, int S, const float *normals, const float *viewdirs, const float *roughness,
                         const float *base_color, const float *dirs, const float *c_raw, const float *a_raw, float sat,
                         const float *env_base, int H, int W, int activation, const float *transform, const float *pdf,
                         float p_diffuse, float p_light, int total_samples, const float *g_out,
                         float *g_c_raw, float *g_a_raw, float *g_dirs, float *g_point, float *grad_env) {
    const EnvMap e = make_env(env_base, H, W, activation, transform);
    const MisParams m = {pdf, p_diffuse, p_light};
    for (int64_t pt = 0; pt < P; ++pt) {
        ShadePoint p;
        shade_point_setup(normals + 3 * pt, viewdirs + 3 * pt, roughness[pt], base_color + 3 * pt, p);
        const float inv = 1.0f / (float)total_samples;
        const float *go = g_out + 16 * pt;
        float gD[3], gS[3], gE[3], gLi[3], gLocal[3];
        for (int c = 0; c < 3; ++c) {
            gD[c] = go[c] * inv; gS[c] = go[3 + c] * inv; gE[c] = go[6 + c] * inv;
            gLi[c] = go[10 + c] * inv; gLocal[c] = go[13 + c] * inv;
        }
        const float gVis = go[9] * inv;
        ShadeAcc acc;
        shade_acc_zero(acc);
        for (int s = 0; s < S; ++s) {
            const int64_t ray = pt * S + s;
            if (pdf)
                shade_sample_backward<true>(p, e, m, dirs + 3 * ray, c_raw + 3 * ray, a_raw[ray], sat, gD, gS, gE, gVis, gLi,
                                            gLocal, grad_env, acc, g_c_raw + 3 * ray, g_a_raw[ray], g_dirs + 3 * ray);
            else
                shade_sample_backward<false>(p, e, m, dirs + 3 * ray, c_raw + 3 * ray, a_raw[ray], sat, gD, gS, gE, gVis, gLi,
                                             gLocal, grad_env, acc, g_c_raw + 3 * ray, g_a_raw[ray], g_dirs + 3 * ray);
        }
        float *o = g_point + 16 * pt;
        std::memset(o, 0, 16 * sizeof(float));
        shade_point_finish(p, acc, o, o[3], o + 4, o + 7);
    }
}

void env_host_forward(int64_t n, const float *dirs, const float *env_base, int H, int W, int activation,
                      const float *transform, float *out) {
    const EnvMap e = make_env(env_base, H, W, activation, transform);
    for (int64_t i = 0; i < n; ++i) {
        EnvTap t;
        float raw[3];
        env_tap(e, dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], t);
        env_fetch(e, t, raw, out + 3 * i);
    }
}

}  // extern "C"
