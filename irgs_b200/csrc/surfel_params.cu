// Parameter-level entry of the tracer, sm_100a: the caller glue of the reference between the surfel
// PARAMETERS and the tracer's inputs / outputs, as kernels with hand-derived backward instead of ~25 element-wise torch ops.
//
// Replaces, of /root/reference:
//   scene/gaussian_model.py:733-747   s = 1 / scaling; R = build_rotation(rotation); ru = R[:,:,0] s_u; rv = R[:,:,1] s_v;
//                                     normals = safe_normalize(flip_align_view(R[:,:,2], means - camera_center))
//   utils/general_utils.py:78-99      build_rotation (quaternion (w, x, y, z), normalised first)
//   utils/general_utils.py:135-146    safe_normalize (eps 1e-20), flip_align_view (sign(dot) has no gradient)
//   scene/gaussian_model.py:751-756   outputs of saturated rays (alpha >= 1 - T_min) divided by alpha, their alpha set to 1
// and the autograd chain of all of it.  One thread per surfel (frames) / per ray (normalisation): HBM streaming, a few
// dozen bytes per item; not a hot kernel, but it removes the torch launches around every trace call.
#include "internal.cuh"

namespace irgs {

struct Frame {
    float q[4], inv_len;     // normalised quaternion, 1 / |raw quaternion|
    float c0[3], c1[3], c2[3];   // columns of R
    float sign, c2_len;      // flip towards the camera, max(|c2|, 1e-20)
};

__device__ __forceinline__ Frame make_frame(const float *__restrict__ rot, const float *__restrict__ mean, float cx, float cy,
                                            float cz, int has_cam) {
    Frame f;
    const float w0 = rot[0], x0 = rot[1], y0 = rot[2], z0 = rot[3];
    const float len = sqrtf(w0 * w0 + x0 * x0 + y0 * y0 + z0 * z0);
    f.inv_len = 1.0f / len;
    const float r = w0 / len, x = x0 / len, y = y0 / len, z = z0 / len;
    f.q[0] = r; f.q[1] = x; f.q[2] = y; f.q[3] = z;
    f.c0[0] = 1.f - 2.f * (y * y + z * z); f.c0[1] = 2.f * (x * y + r * z); f.c0[2] = 2.f * (x * z - r * y);
    f.c1[0] = 2.f * (x * y - r * z); f.c1[1] = 1.f - 2.f * (x * x + z * z); f.c1[2] = 2.f * (y * z + r * x);
    f.c2[0] = 2.f * (x * z + r * y); f.c2[1] = 2.f * (y * z - r * x); f.c2[2] = 1.f - 2.f * (x * x + y * y);
    f.sign = 1.f;
    if (has_cam) {   // flip_align_view(normal, means - camera_center): dot(normal, -(mean - cam)) >= 0 keeps the normal
        const float d = f.c2[0] * -(mean[0] - cx) + f.c2[1] * -(mean[1] - cy) + f.c2[2] * -(mean[2] - cz);
        f.sign = d >= 0.f ? 1.f : -1.f;
    }
    f.c2_len = fmaxf(sqrtf(f.c2[0] * f.c2[0] + f.c2[1] * f.c2[1] + f.c2[2] * f.c2[2]), 1e-20f);
    return f;
}

__global__ void frames_forward_kernel(int64_t n, const float *__restrict__ means, const float *__restrict__ scales,
                                      const float *__restrict__ rotations, float cx, float cy, float cz, int has_cam,
                                      float *__restrict__ ru, float *__restrict__ rv, float *__restrict__ normals) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const Frame f = make_frame(rotations + 4 * g, means + 3 * g, cx, cy, cz, has_cam);
    const float su = 1.0f / scales[2 * g], sv = 1.0f / scales[2 * g + 1];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ru[3 * g + k] = f.c0[k] * su;
        rv[3 * g + k] = f.c1[k] * sv;
        normals[3 * g + k] = (f.c2[k] * f.sign) / f.c2_len;
    }
}

// Fused row [N,64] -> gradients of the PARAMETERS: rows 0-2 / 3 / 16.. are copied out (means, opacity, SH: one thread per
// float, coalesced), rows 4-12 (d/dru, d/drv, d/dnormals) go through the chain rule of make_frame (one thread per surfel).
__global__ void unpack_params_kernel(const float *__restrict__ fused, int64_t n, int K, const float *__restrict__ means,
                                     const float *__restrict__ scales, const float *__restrict__ rotations, float cx, float cy,
                                     float cz, int has_cam, float *__restrict__ gm, float *__restrict__ go,
                                     float *__restrict__ gscales, float *__restrict__ grot, float *__restrict__ gsh) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * IRGS_GRAD_STRIDE) return;
    const int64_t g = idx / IRGS_GRAD_STRIDE;
    const int fi = (int)(idx % IRGS_GRAD_STRIDE);
    const float v = fused[idx];
    if (fi < 3) gm[3 * g + fi] = v;
    else if (fi == 3) go[g] = v;
    else if (fi >= 16) {
        const int k = (fi - 16) / 3;
        if (k < K) gsh[(g * K + k) * 3 + (fi - 16) % 3] = v;
    } else if (fi == 4) {
        const float *row = fused + g * IRGS_GRAD_STRIDE;
        const float g_ru[3] = {row[4], row[5], row[6]}, g_rv[3] = {row[7], row[8], row[9]}, g_n[3] = {row[10], row[11], row[12]};
        const Frame f = make_frame(rotations + 4 * g, means + 3 * g, cx, cy, cz, has_cam);
        const float su = 1.0f / scales[2 * g], sv = 1.0f / scales[2 * g + 1];
        // ru = c0 / s_u: d/dc0 = g_ru / s_u, d/ds_u = -(g_ru . c0) / s_u^2   (likewise rv)
        float A[3], B[3], C[3];
        float dsu = 0.f, dsv = 0.f, ndot = 0.f;
        float nrm[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            A[k] = g_ru[k] * su; B[k] = g_rv[k] * sv;
            dsu -= g_ru[k] * f.c0[k]; dsv -= g_rv[k] * f.c1[k];
            nrm[k] = f.c2[k] / f.c2_len;          // the unflipped unit normal
            ndot += nrm[k] * (g_n[k] * f.sign);
        }
        gscales[2 * g] = dsu * su * su;
        gscales[2 * g + 1] = dsv * sv * sv;
        // normals = sign * c2 / max(|c2|, eps): d/dc2 = sign * (g_n - n (n . g_n)) / |c2|  (the clamp is inactive for unit quaternions)
#pragma unroll
        for (int k = 0; k < 3; ++k) C[k] = (g_n[k] * f.sign - nrm[k] * ndot) / f.c2_len;
        // R(q), q = (r, x, y, z) normalised (general_utils.py:90-98): gradient with respect to q
        const float r = f.q[0], x = f.q[1], y = f.q[2], z = f.q[3];
        float gq[4];
        gq[0] = 2.f * (A[1] * z - A[2] * y - B[0] * z + B[2] * x + C[0] * y - C[1] * x);
        gq[1] = 2.f * (A[1] * y + A[2] * z + B[0] * y - 2.f * B[1] * x + B[2] * r + C[0] * z - C[1] * r - 2.f * C[2] * x);
        gq[2] = 2.f * (-2.f * A[0] * y + A[1] * x - A[2] * r + B[0] * x + B[2] * z + C[0] * r + C[1] * z - 2.f * C[2] * y);
        gq[3] = 2.f * (-2.f * A[0] * z + A[1] * r + A[2] * x - B[0] * r - 2.f * B[1] * z + B[2] * y + C[0] * x + C[1] * y);
        // q = raw / |raw|: d/draw = (gq - q (q . gq)) / |raw|
        const float qd = gq[0] * r + gq[1] * x + gq[2] * y + gq[3] * z;
#pragma unroll
        for (int k = 0; k < 4; ++k) grot[4 * g + k] = (gq[k] - f.q[k] * qd) * f.inv_len;
    }
}

// scene/gaussian_model.py:751-756, out of place (the tracer's backward needs the raw accumulations).
__global__ void normalize_outputs_kernel(int64_t n_rays, int S, float threshold, const float *__restrict__ color,
                                         const float *__restrict__ normal, const float *__restrict__ feature,
                                         const float *__restrict__ depth, const float *__restrict__ alpha,
                                         float *__restrict__ o_color, float *__restrict__ o_normal, float *__restrict__ o_feature,
                                         float *__restrict__ o_depth, float *__restrict__ o_alpha) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float a = alpha[r];
    const bool sat = !(a < threshold);     // torch.where(alpha < 1 - T_min, x, x / alpha)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o_color[3 * r + k] = sat ? color[3 * r + k] / a : color[3 * r + k];
        o_normal[3 * r + k] = sat ? normal[3 * r + k] / a : normal[3 * r + k];
    }
    for (int j = 0; j < S; ++j) o_feature[r * S + j] = sat ? feature[r * S + j] / a : feature[r * S + j];
    o_depth[r] = sat ? depth[r] / a : depth[r];
    o_alpha[r] = sat ? 1.0f : a;
}

// Backward of the above: g_* arrive for the normalised outputs and are rewritten IN PLACE as gradients of the raw ones.
__global__ void normalize_outputs_backward_kernel(int64_t n_rays, int S, float threshold, const float *__restrict__ color,
                                                  const float *__restrict__ normal, const float *__restrict__ feature,
                                                  const float *__restrict__ depth, const float *__restrict__ alpha,
                                                  float *__restrict__ g_color, float *__restrict__ g_normal,
                                                  float *__restrict__ g_feature, float *__restrict__ g_depth,
                                                  float *__restrict__ g_alpha) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    const float a = alpha[r];
    if (a < threshold) return;             // pass-through
    const float inv = 1.0f / a;
    float ga = 0.f;                        // d(x / a)/da = -x / a^2; the output alpha is the constant 1
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float gc = g_color[3 * r + k], gn = g_normal[3 * r + k];
        ga -= (gc * color[3 * r + k] + gn * normal[3 * r + k]) * inv * inv;
        g_color[3 * r + k] = gc * inv;
        g_normal[3 * r + k] = gn * inv;
    }
    for (int j = 0; j < S; ++j) {
        const float gf = g_feature[r * S + j];
        ga -= gf * feature[r * S + j] * inv * inv;
        g_feature[r * S + j] = gf * inv;
    }
    const float gd = g_depth[r];
    ga -= gd * depth[r] * inv * inv;
    g_depth[r] = gd * inv;
    g_alpha[r] = ga;
}

static int bad(const char *msg) {
    set_error(msg);
    return 1;
}

}  // namespace irgs

using namespace irgs;

extern "C" {

int irgs_surfel_frames(const float *means, const float *scales, const float *rotations, const float *camera_center_host,
                       int64_t n, float *ru, float *rv, float *normals, void *stream) {
    if (n < 0) return bad("n < 0");
    if (n == 0) return 0;
    if (!means || !scales || !rotations || !ru || !rv || !normals) return bad("surfel_frames: null array");
    const float *c = camera_center_host;
    frames_forward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, means, scales, rotations, c ? c[0] : 0.f, c ? c[1] : 0.f, c ? c[2] : 0.f, c ? 1 : 0, ru, rv, normals);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int irgs_unpack_grads_params(const float *grad_fused, int64_t n, int K, const float *means, const float *scales,
                             const float *rotations, const float *camera_center_host, float *grad_means, float *grad_opacity,
                             float *grad_scales, float *grad_rotations, float *grad_shs, void *stream) {
    if (n <= 0) return 0;
    if (K < 1) return bad("K must be positive");
    if (!grad_fused || !means || !scales || !rotations || !grad_means || !grad_opacity || !grad_scales || !grad_rotations || !grad_shs)
        return bad("unpack_grads_params: null array");
    const float *c = camera_center_host;
    const int64_t total = n * IRGS_GRAD_STRIDE;
    unpack_params_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        grad_fused, n, K, means, scales, rotations, c ? c[0] : 0.f, c ? c[1] : 0.f, c ? c[2] : 0.f, c ? 1 : 0, grad_means,
        grad_opacity, grad_scales, grad_rotations, grad_shs);
    count_launch();
    if (K > 16) IRGS_CHECK(cudaMemset2DAsync(grad_shs + 48, sizeof(float) * 3 * K, 0, sizeof(float) * 3 * (K - 16), (size_t)n,
                                             (cudaStream_t)stream));   // coefficients k >= 16 never receive gradient
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int irgs_normalize_outputs(int64_t n_rays, int S, float threshold, const float *color, const float *normal, const float *feature,
                           const float *depth, const float *alpha, float *out_color, float *out_normal, float *out_feature,
                           float *out_depth, float *out_alpha, void *stream) {
    if (n_rays <= 0) return 0;
    if (S < 0 || S > IRGS_MAX_FEATURES) return bad("feature channels S must be in [0, 12]");
    normalize_outputs_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_rays, S, threshold, color, normal, feature, depth, alpha, out_color, out_normal, out_feature, out_depth, out_alpha);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int irgs_normalize_outputs_backward(int64_t n_rays, int S, float threshold, const float *color, const float *normal,
                                    const float *feature, const float *depth, const float *alpha, float *g_color,
                                    float *g_normal, float *g_feature, float *g_depth, float *g_alpha, void *stream) {
    if (n_rays <= 0) return 0;
    if (S < 0 || S > IRGS_MAX_FEATURES) return bad("feature channels S must be in [0, 12]");
    normalize_outputs_backward_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_rays, S, threshold, color, normal, feature, depth, alpha, g_color, g_normal, g_feature, g_depth, g_alpha);
    count_launch();
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
