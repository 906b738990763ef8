// C ABI of libirgs_b200.so (see include/irgs_b200.h for the contract and the reference interface each entry replaces).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <new>

#include "internal.cuh"

namespace irgs {

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

void set_error(const std::string &msg) { g_err = msg; }
bool check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev) ok = check(cudaSetDevice(dev), "cudaSetDevice");
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static int fail(const char *msg) {
    set_error(msg);
    return 1;
}

// The slot of a stream: found or assigned on first use.  When more than MAX_SLOTS distinct streams have been seen the device
// is synchronised (nothing of this handle is in flight afterwards) and the table starts over.
int slot_for(irgs_tracer *h, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(h->slot_mutex);
    for (int i = 0; i < h->n_slots; ++i)
        if (h->slot_stream[i] == s) return i;
    if (h->n_slots == irgs_tracer::MAX_SLOTS) {
        if (!check(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return -1;
        h->n_slots = 0;
    }
    h->slot_stream[h->n_slots] = s;
    return h->n_slots++;
}

static int validate_trace(const irgs_tracer *h, int64_t n_rays, int S, int K, int deg, int hit_cap) {
    if (!h) return fail("null tracer handle");
    if (!h->built) return fail("trace called before build_bvh");
    if (n_rays < 0) return fail("n_rays < 0");
    if (S < 0 || S > IRGS_MAX_FEATURES) return fail("feature channels S must be in [0, 12] (MAX_FEATURE_SIZE)");
    if (deg < 0 || deg > 3) return fail("SH degree must be in [0, 3]");
    if (K < (deg + 1) * (deg + 1)) return fail("shs.size(1) must be >= (deg+1)^2");
    if (hit_cap < 0 || (hit_cap & 3)) return fail("hit_cap must be a non-negative multiple of 4");
    return 0;
}

}  // namespace irgs

using namespace irgs;

// Saved hit-list entries per ray on the host-buffer paths (and the Python default): the lists are reserved, not touched,
// beyond a ray's count.  96 and not less: p99.9 of the C3 rays is 46 hits, but every ray beyond the cap goes through the re-trace
// backward, a ~0.6 ms serial chain even for one ray -- a cap of 64 added 0.6 ms to the backward of every call from 2^16 rays up.
#define IRGS_HOST_HIT_CAP 96

extern "C" {

const char *irgs_last_error(void) { return g_err.c_str(); }
int irgs_version(void) { return 100; }

int irgs_tracer_create(irgs_tracer_t **out, int device) {
    if (!out) return fail("null out pointer");
    *out = nullptr;
    int count = 0;
    IRGS_CHECK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail("invalid CUDA device index");
    DeviceGuard guard(device);
    if (!guard.ok) return 1;
    irgs_tracer *h = new (std::nothrow) irgs_tracer();
    if (!h) return fail("out of host memory");
    h->device = device;
    cudaDeviceProp prop;
    IRGS_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete h;
        return fail("libirgs_b200 is built for sm_100a (B200) only");
    }
    h->sm_count = prop.multiProcessorCount;
    if (!check(cudaMalloc(&h->scene, 24 * sizeof(float)), "cudaMalloc") ||
        !check(cudaMalloc(&h->counter, irgs_tracer::MAX_SLOTS * sizeof(unsigned long long)), "cudaMalloc") ||
        !check(cudaMalloc(&h->stats, 8 * sizeof(unsigned long long)), "cudaMalloc")) {
        irgs_tracer_destroy(h);
        return 1;
    }
    cudaMemset(h->stats, 0, 8 * sizeof(unsigned long long));
    *out = h;
    return 0;
}

int irgs_tracer_destroy(irgs_tracer_t *h) {
    if (!h) return 0;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->nodes); cudaFree(h->qnodes); cudaFree(h->qnodes4); cudaFree(h->even); cudaFree(h->wide_kids); cudaFree(h->wide_counts); cudaFree(h->boxes); cudaFree(h->codes); cudaFree(h->codes_alt); cudaFree(h->order);
    cudaFree(h->order_alt); cudaFree(h->leaf_parent); cudaFree(h->node_parent); cudaFree(h->flags);
    cudaFree(h->ploc_cid); cudaFree(h->ploc_box); cudaFree(h->ploc_nn); cudaFree(h->ploc_counts); cudaFree(h->ploc_offs); cudaFree(h->ploc_totals);
    cudaFree(h->radix_hist); cudaFree(h->scene); cudaFree(h->recs); cudaFree(h->inv_order); cudaFree(h->counter); cudaFree(h->stats);
    for (int i = 0; i < irgs_tracer::MAX_SLOTS; ++i) {
        for (int k = 0; k < 2; ++k) { cudaFree(h->rsort_keys[i][k]); cudaFree(h->rsort_vals[i][k]); }
        cudaFree(h->rsort_hist[i]);
        cudaFree(h->cand[i]);
        cudaFree(h->inc_pts[i]);
        cudaFree(h->ray_scratch[i]);
        cudaFree(h->hit_rgb[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (h->stage[i]) cudaFree(h->stage[i]);
        if (h->hs[i]) cudaStreamDestroy(h->hs[i]);
        if (h->hev[i]) cudaEventDestroy(h->hev[i]);
    }
    for (int i = 0; i < irgs_tracer::RING; ++i) {
        if (h->ring[i]) cudaFree(h->ring[i]);
        if (h->ring_full[i]) cudaEventDestroy(h->ring_full[i]);
        if (h->ring_free[i]) cudaEventDestroy(h->ring_free[i]);
    }
    if (h->hcopy) cudaStreamDestroy(h->hcopy);
    cudaFree(h->pt_buf);
    delete h;
    return 0;
}

static int build_common(irgs_tracer_t *h, int64_t n, bool refit) {
    if (!h) return fail("null tracer handle");
    if (n <= 0) return fail("n_surfels must be positive");
    if (n > (int64_t)1 << 30) return fail("n_surfels too large");
    if (refit) {
        if (!h->built) return fail("update_bvh called before build_bvh");
        if (n != h->n) return fail("update_bvh must keep the number of surfels unchanged");
    } else {
        if (lbvh_reserve(h, n)) return 1;
        h->n = n;
    }
    return 0;
}

int irgs_build_from_proxy(irgs_tracer_t *h, const float *vertices_b, int64_t n, int vps, void *stream) {
    if (build_common(h, n, false)) return 1;
    if (vps <= 0) return fail("verts_per_surfel must be positive");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (launch_bounds_from_proxy(h, vertices_b, vps, s)) return 1;
    return lbvh_build(h, false, s);
}

int irgs_refit_from_proxy(irgs_tracer_t *h, const float *vertices_b, int64_t n, int vps, void *stream) {
    if (build_common(h, n, true)) return 1;
    if (vps <= 0) return fail("verts_per_surfel must be positive");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (launch_bounds_from_proxy(h, vertices_b, vps, s)) return 1;
    return lbvh_build(h, true, s);
}

int irgs_build_from_surfels(irgs_tracer_t *h, const float *means, const float *opacity, const float *ru,
                            const float *rv, const float *normals, int64_t n, float alpha_min, void *stream) {
    if (build_common(h, n, false)) return 1;
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (launch_bounds_from_surfels(h, means, opacity, ru, rv, normals, alpha_min, s)) return 1;
    return lbvh_build(h, false, s);
}

int irgs_refit_from_surfels(irgs_tracer_t *h, const float *means, const float *opacity, const float *ru,
                            const float *rv, const float *normals, int64_t n, float alpha_min, void *stream) {
    if (build_common(h, n, true)) return 1;
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (launch_bounds_from_surfels(h, means, opacity, ru, rv, normals, alpha_min, s)) return 1;
    return lbvh_build(h, true, s);
}

int64_t irgs_num_surfels(const irgs_tracer_t *h) { return h && h->built ? h->n : 0; }

int irgs_get_bounds(irgs_tracer_t *h, float *surfel_bounds, float *root_bound, void *stream) {
    if (!h || !h->built) return fail("get_bounds called before build_bvh");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (surfel_bounds)
        IRGS_CHECK(cudaMemcpyAsync(surfel_bounds, h->boxes, sizeof(float) * 6 * (size_t)h->n, cudaMemcpyDeviceToDevice, s));
    if (root_bound) IRGS_CHECK(cudaMemcpyAsync(root_bound, h->scene + 6, sizeof(float) * 6, cudaMemcpyDeviceToDevice, s));
    return 0;
}

static TraceArgs make_args(int64_t n_rays, int n_surf, int S, int K, int deg, const float *rays_o, const float *rays_d,
                           const float *means, const float *opacity, const float *ru, const float *rv,
                           const float *normals, const float *features, const float *shs, float alpha_min, float T_min,
                           int back_culling) {
    TraceArgs a;
    memset(&a, 0, sizeof a);
    a.n_rays = n_rays; a.n_surf = n_surf; a.S = S; a.K = K; a.deg = deg; a.back_culling = back_culling ? 1 : 0;
    a.alpha_min = alpha_min; a.T_min = T_min;
    a.rays_o = rays_o; a.rays_d = rays_d;
    a.means = means; a.opacity = opacity; a.ru = ru; a.rv = rv; a.normals = normals; a.features = features; a.shs = shs;
    return a;
}

int irgs_intersection_test(irgs_tracer_t *h, int64_t n_rays, const float *rays_o, const float *rays_d,
                           const float *means, const float *opacity, const float *ru, const float *rv,
                           const float *normals, float alpha_min, uint8_t *out, void *stream) {
    if (validate_trace(h, n_rays, 0, 16, 3, 0)) return 1;
    if (n_rays == 0) return 0;
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    TraceArgs a = make_args(n_rays, (int)h->n, 0, 16, 3, rays_o, rays_d, means, opacity, ru, rv, normals, nullptr, nullptr,
                            alpha_min, 0.f, 0);
    if (launch_pack_records(h, a, s)) return 1;
    return launch_intersection_test(h, a, out, s);
}

static int validate_incident(const irgs_incident_t *gen, bool one_call = true) {
    if (!gen) return fail("null incident-ray descriptor");
    if (gen->n_points < 0) return fail("n_points < 0");
    if (gen->sample_num < 1) return fail("sample_num must be positive");
    if (gen->n_points > 0 && (!gen->position || !gen->normals)) return fail("position / normals must not be null");
    if (one_call && gen->n_points * (int64_t)gen->sample_num >= ((int64_t)1 << 31)) return fail("n_points * sample_num must be below 2^31 per call");
    return 0;
}
static void set_generator(TraceArgs &a, const irgs_incident_t *gen) {
    if (!gen) return;
    a.gen_pos = gen->position; a.gen_nrm = gen->normals; a.gen_azim = gen->azimuth;
    a.gen_S = gen->sample_num; a.gen_tmin = gen->t_min; a.gen_P = gen->n_points;
}

static int validate_camera(const irgs_camera_t *cam) {
    if (!cam) return fail("null camera descriptor");
    if (cam->width < 1 || cam->height < 1) return fail("camera: width and height must be positive");
    if (!(cam->fx > 0.f) || !(cam->fy > 0.f)) return fail("camera: focal lengths must be positive");
    if ((int64_t)cam->width * cam->height >= ((int64_t)1 << 31)) return fail("camera: width * height must be below 2^31");
    return 0;
}
static void set_camera(TraceArgs &a, const irgs_camera_t *cam) {
    if (!cam) return;
    a.cam_W = cam->width; a.cam_H = cam->height; a.cam_fx = cam->fx; a.cam_fy = cam->fy;
    for (int k = 0; k < 3; ++k) a.cam_o[k] = cam->origin[k];
    for (int k = 0; k < 9; ++k) a.cam_M[k] = cam->cam_to_world[k];
}

static int trace_forward_impl(irgs_tracer_t *h, const irgs_incident_t *gen, int64_t n_rays, int S, int K, int deg,
                              const float *rays_o, const float *rays_d, const float *means, const float *opacity,
                              const float *ru, const float *rv, const float *normals, const float *features,
                              const float *shs, float *out_color, float *out_normal, float *out_feature, float *out_depth,
                              float *out_alpha, int32_t *out_hit_count, int32_t *out_hits, int hit_cap, float alpha_min,
                              float T_min, int back_culling, void *stream, const irgs_camera_t *cam = nullptr) {
    if (validate_trace(h, n_rays, S, K, deg, hit_cap)) return 1;
    if (n_rays == 0) return 0;
    if (out_hits && hit_cap == 0) out_hits = nullptr;
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    TraceArgs a = make_args(n_rays, (int)h->n, S, K, deg, rays_o, rays_d, means, opacity, ru, rv, normals, features, shs,
                            alpha_min, T_min, back_culling);
    set_generator(a, gen);
    set_camera(a, cam);
    a.color = out_color; a.normal = out_normal; a.feature = out_feature; a.depth = out_depth; a.alpha = out_alpha;
    a.hit_count = out_hit_count; a.hits = out_hits; a.hit_cap = hit_cap;
    if (launch_pack_records(h, a, s) || launch_incident_prepare(h, a, s)) return 1;
    return launch_trace_forward(h, a, s);
}

int irgs_trace_forward(irgs_tracer_t *h, int64_t n_rays, int S, int K, int deg, const float *rays_o,
                       const float *rays_d, const float *means, const float *opacity, const float *ru,
                       const float *rv, const float *normals, const float *features, const float *shs,
                       float *out_color, float *out_normal, float *out_feature, float *out_depth, float *out_alpha,
                       int32_t *out_hit_count, int32_t *out_hits, int hit_cap, float alpha_min, float T_min,
                       int back_culling, void *stream) {
    if (n_rays > 0 && (!rays_o || !rays_d)) return fail("rays_o / rays_d must not be null");
    return trace_forward_impl(h, nullptr, n_rays, S, K, deg, rays_o, rays_d, means, opacity, ru, rv, normals, features, shs,
                              out_color, out_normal, out_feature, out_depth, out_alpha, out_hit_count, out_hits, hit_cap,
                              alpha_min, T_min, back_culling, stream);
}

int irgs_trace_forward_incident(irgs_tracer_t *h, const irgs_incident_t *gen, int S, int K, int deg, const float *means,
                                const float *opacity, const float *ru, const float *rv, const float *normals,
                                const float *features, const float *shs, float *out_color, float *out_normal,
                                float *out_feature, float *out_depth, float *out_alpha, int32_t *out_hit_count,
                                int32_t *out_hits, int hit_cap, float alpha_min, float T_min, int back_culling,
                                void *stream) {
    if (validate_incident(gen)) return 1;
    return trace_forward_impl(h, gen, gen->n_points * gen->sample_num, S, K, deg, nullptr, nullptr, means, opacity, ru, rv,
                              normals, features, shs, out_color, out_normal, out_feature, out_depth, out_alpha,
                              out_hit_count, out_hits, hit_cap, alpha_min, T_min, back_culling, stream);
}

int irgs_incident_rays(const irgs_incident_t *gen, float *rays_o, float *rays_d, void *stream) {
    if (validate_incident(gen)) return 1;
    return launch_incident_rays(gen->position, gen->normals, gen->azimuth, gen->n_points, gen->sample_num, gen->t_min, rays_o,
                                rays_d, (cudaStream_t)stream);
}

static int trace_backward_impl(irgs_tracer_t *h, const irgs_incident_t *gen, int64_t n_rays, int S, int K, int deg,
                               const float *rays_o, const float *rays_d, const float *means, const float *opacity,
                               const float *ru, const float *rv, const float *normals, const float *features,
                               const float *shs, const float *color, const float *normal, const float *feature,
                               const float *depth, const float *alpha, const int32_t *hit_count, const int32_t *hits,
                               int hit_cap, const float *gout_color, const float *gout_normal, const float *gout_feature,
                               const float *gout_depth, const float *gout_alpha, float *grad_rays_o, float *grad_rays_d,
                               float *grad_fused, float *grad_features, float alpha_min, float T_min, int back_culling,
                               void *stream, const irgs_camera_t *cam = nullptr) {
    if (validate_trace(h, n_rays, S, K, deg, hit_cap)) return 1;
    if (n_rays == 0) return 0;
    if (!grad_fused) return fail("grad_fused must not be null");
    if (!grad_rays_o || !grad_rays_d) return fail("grad_rays_o / grad_rays_d must not be null");
    if (S > 0 && !grad_features) return fail("grad_features must not be null when S > 0");
    DeviceGuard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    TraceArgs a = make_args(n_rays, (int)h->n, S, K, deg, rays_o, rays_d, means, opacity, ru, rv, normals, features, shs,
                            alpha_min, T_min, back_culling);
    set_generator(a, gen);
    set_camera(a, cam);
    a.color = const_cast<float *>(color); a.normal = const_cast<float *>(normal);
    a.feature = const_cast<float *>(feature); a.depth = const_cast<float *>(depth); a.alpha = const_cast<float *>(alpha);
    if (hits && hit_count && hit_cap > 0) {
        a.hit_count = const_cast<int32_t *>(hit_count); a.hits = const_cast<int32_t *>(hits); a.hit_cap = hit_cap;
    }
    a.gC = gout_color; a.gN = gout_normal; a.gF = gout_feature; a.gD = gout_depth; a.gO = gout_alpha;
    a.g_rays_o = grad_rays_o; a.g_rays_d = grad_rays_d; a.grad_fused = grad_fused; a.grad_features = grad_features;
    // the records must be consistent with the arrays handed to this call: packed again unless the caller vouches that nothing
    // has packed or rebuilt since the forward of these very arrays (irgs_set_option("skip_next_pack"), see "pack_epoch")
    const bool skip = h->skip_next_pack != 0;
    h->skip_next_pack = 0;
    if ((!skip && launch_pack_records(h, a, s)) || launch_incident_prepare(h, a, s)) return 1;
    return launch_trace_backward(h, a, s);
}

int irgs_trace_backward(irgs_tracer_t *h, int64_t n_rays, int S, int K, int deg, const float *rays_o,
                        const float *rays_d, const float *means, const float *opacity, const float *ru,
                        const float *rv, const float *normals, const float *features, const float *shs,
                        const float *color, const float *normal, const float *feature, const float *depth,
                        const float *alpha, const int32_t *hit_count, const int32_t *hits, int hit_cap,
                        const float *gout_color, const float *gout_normal, const float *gout_feature,
                        const float *gout_depth, const float *gout_alpha, float *grad_rays_o, float *grad_rays_d,
                        float *grad_fused, float *grad_features, float alpha_min, float T_min, int back_culling,
                        void *stream) {
    if (n_rays > 0 && (!rays_o || !rays_d)) return fail("rays_o / rays_d must not be null");
    return trace_backward_impl(h, nullptr, n_rays, S, K, deg, rays_o, rays_d, means, opacity, ru, rv, normals, features, shs,
                               color, normal, feature, depth, alpha, hit_count, hits, hit_cap, gout_color, gout_normal,
                               gout_feature, gout_depth, gout_alpha, grad_rays_o, grad_rays_d, grad_fused, grad_features,
                               alpha_min, T_min, back_culling, stream);
}

int irgs_trace_backward_incident(irgs_tracer_t *h, const irgs_incident_t *gen, int S, int K, int deg, const float *means,
                                 const float *opacity, const float *ru, const float *rv, const float *normals,
                                 const float *features, const float *shs, const float *color, const float *normal,
                                 const float *feature, const float *depth, const float *alpha, const int32_t *hit_count,
                                 const int32_t *hits, int hit_cap, const float *gout_color, const float *gout_normal,
                                 const float *gout_feature, const float *gout_depth, const float *gout_alpha,
                                 float *scratch_grad_rays_o, float *scratch_grad_rays_d, float *grad_position,
                                 float *grad_normal_pt, float *grad_fused, float *grad_features, float alpha_min, float T_min,
                                 int back_culling, void *stream) {
    if (validate_incident(gen)) return 1;
    if (gen->n_points > 0 && (!grad_position || !grad_normal_pt)) return fail("grad_position / grad_normal_pt must not be null");
    if (trace_backward_impl(h, gen, gen->n_points * gen->sample_num, S, K, deg, nullptr, nullptr, means, opacity, ru, rv,
                            normals, features, shs, color, normal, feature, depth, alpha, hit_count, hits, hit_cap,
                            gout_color, gout_normal, gout_feature, gout_depth, gout_alpha, scratch_grad_rays_o,
                            scratch_grad_rays_d, grad_fused, grad_features, alpha_min, T_min, back_culling, stream))
        return 1;
    DeviceGuard guard(h->device);
    return launch_incident_backward(gen->position, gen->normals, gen->azimuth, gen->n_points, gen->sample_num, gen->t_min,
                                    scratch_grad_rays_o, scratch_grad_rays_d, grad_position, grad_normal_pt,
                                    (cudaStream_t)stream);
}

int irgs_trace_forward_camera(irgs_tracer_t *h, const irgs_camera_t *cam, int S, int K, int deg, const float *means,
                              const float *opacity, const float *ru, const float *rv, const float *normals,
                              const float *features, const float *shs, float *out_color, float *out_normal, float *out_feature,
                              float *out_depth, float *out_alpha, int32_t *out_hit_count, int32_t *out_hits, int hit_cap,
                              float alpha_min, float T_min, int back_culling, void *stream) {
    if (validate_camera(cam)) return 1;
    return trace_forward_impl(h, nullptr, (int64_t)cam->width * cam->height, S, K, deg, nullptr, nullptr, means, opacity, ru, rv,
                              normals, features, shs, out_color, out_normal, out_feature, out_depth, out_alpha, out_hit_count,
                              out_hits, hit_cap, alpha_min, T_min, back_culling, stream, cam);
}

int irgs_trace_backward_camera(irgs_tracer_t *h, const irgs_camera_t *cam, int S, int K, int deg, const float *means,
                               const float *opacity, const float *ru, const float *rv, const float *normals,
                               const float *features, const float *shs, const float *color, const float *normal,
                               const float *feature, const float *depth, const float *alpha, const int32_t *hit_count,
                               const int32_t *hits, int hit_cap, const float *gout_color, const float *gout_normal,
                               const float *gout_feature, const float *gout_depth, const float *gout_alpha,
                               float *scratch_grad_rays_o, float *scratch_grad_rays_d, float *grad_fused, float *grad_features,
                               float alpha_min, float T_min, int back_culling, void *stream) {
    if (validate_camera(cam)) return 1;
    return trace_backward_impl(h, nullptr, (int64_t)cam->width * cam->height, S, K, deg, nullptr, nullptr, means, opacity, ru, rv,
                               normals, features, shs, color, normal, feature, depth, alpha, hit_count, hits, hit_cap, gout_color,
                               gout_normal, gout_feature, gout_depth, gout_alpha, scratch_grad_rays_o, scratch_grad_rays_d,
                               grad_fused, grad_features, alpha_min, T_min, back_culling, stream, cam);
}

int irgs_camera_rays(const irgs_camera_t *cam, float *rays_o, float *rays_d, void *stream) {
    if (validate_camera(cam)) return 1;
    TraceArgs a;
    memset(&a, 0, sizeof a);
    set_camera(a, cam);
    a.n_rays = (int64_t)cam->width * cam->height;
    return launch_generated_rays(a, rays_o, rays_d, (cudaStream_t)stream);
}

int irgs_unpack_grads(const float *grad_fused, int64_t n, int K, float *gm, float *go, float *gru, float *grv,
                      float *gn, float *gsh, void *stream) {
    if (n <= 0) return 0;
    if (K < 1) return fail("K must be positive");
    return launch_unpack_grads(grad_fused, n, K, gm, go, gru, grv, gn, gsh, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ host-buffer path
static int host_prepare(irgs_tracer *h, int64_t floats_per_stream, int64_t ray_floats) {
    for (int i = 0; i < 2; ++i) {
        if (!h->hs[i]) IRGS_CHECK(cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking));
        if (!h->hev[i]) IRGS_CHECK(cudaEventCreateWithFlags(&h->hev[i], cudaEventDisableTiming));
    }
    if (!h->hcopy) IRGS_CHECK(cudaStreamCreateWithFlags(&h->hcopy, cudaStreamNonBlocking));
    for (int i = 0; i < irgs_tracer::RING; ++i) {
        if (!h->ring_full[i]) IRGS_CHECK(cudaEventCreateWithFlags(&h->ring_full[i], cudaEventDisableTiming));
        if (!h->ring_free[i]) IRGS_CHECK(cudaEventCreateWithFlags(&h->ring_free[i], cudaEventDisableTiming));
    }
    if (floats_per_stream > h->stage_floats) {
        for (int i = 0; i < 2; ++i) {
            if (h->stage[i]) cudaFree(h->stage[i]);
            h->stage[i] = nullptr;
            IRGS_CHECK(cudaMalloc(&h->stage[i], sizeof(float) * (size_t)floats_per_stream));
        }
        h->stage_floats = floats_per_stream;
    }
    if (ray_floats > 0 && ray_floats > h->ring_floats) {
        for (int i = 0; i < irgs_tracer::RING; ++i) {
            if (h->ring[i]) cudaFree(h->ring[i]);
            h->ring[i] = nullptr;
            IRGS_CHECK(cudaMalloc(&h->ring[i], sizeof(float) * (size_t)ray_floats));
        }
        h->ring_floats = ray_floats;
    }
    return 0;
}

// The internal streams of the host-buffer entry points start after everything the caller has submitted to `caller` so far
// (nullptr: the legacy default stream) -- an event, not a device-wide synchronisation.
static int host_order_after(irgs_tracer *h, cudaStream_t caller) {
    IRGS_CHECK(cudaEventRecord(h->hev[0], caller));
    IRGS_CHECK(cudaStreamWaitEvent(h->hs[0], h->hev[0], 0));
    IRGS_CHECK(cudaStreamWaitEvent(h->hs[1], h->hev[0], 0));
    IRGS_CHECK(cudaStreamWaitEvent(h->hcopy, h->hev[0], 0));
    return 0;
}

static int trace_host_impl(irgs_tracer *h, bool with_backward, int64_t n_rays, int S, int K, int deg,
                           const float *rays_o_host, const float *rays_d_host, const float *means,
                           const float *opacity, const float *ru, const float *rv, const float *normals,
                           const float *features, const float *shs, const float *gC, const float *gN, const float *gF,
                           const float *gD, const float *gO, int64_t gout_period, float *out_color_host,
                           float *out_normal_host, float *out_feature_host, float *out_depth_host,
                           float *out_alpha_host, float *g_rays_o_host, float *g_rays_d_host, float *grad_fused,
                           float *grad_features, float alpha_min, float T_min, int back_culling, int64_t chunk) {
    const int hit_cap = with_backward ? IRGS_HOST_HIT_CAP : 0;
    if (validate_trace(h, n_rays, S, K, deg, hit_cap)) return 1;
    if (n_rays == 0) return 0;
    if (chunk <= 0) chunk = (int64_t)1 << 21;
    if (chunk > n_rays) chunk = n_rays;
    if (with_backward && !grad_fused) return fail("grad_fused must not be null");
    if (with_backward && gout_period <= 0) return fail("gout_period must be positive");
    DeviceGuard guard(h->device);
    // Rays travel host -> device on a dedicated copy stream into a ring of RING buffers (o[3c] d[3c] each), running ahead
    // of the two compute streams: with the copies issued on the compute streams themselves, both streams ended up waiting
    // for their next chunk's rays at the same time (the two persistent forward kernels serialise, the two backward
    // kernels then finish together) and the GPU idled for one copy per pair of chunks -- 39 ms of a 314 ms C3 step.
    // Per-stream staging layout (floats): color[3c] normal[3c] feature[S c] depth[c] alpha[c] hit_count[c] hits[cap c]
    //                                     g_o[3c] g_d[3c]
    const int64_t per_ray = 3 + 3 + S + 1 + 1 + (with_backward ? 1 + hit_cap + 6 : 0);
    if (host_prepare(h, per_ray * chunk, 6 * chunk)) return 1;
    if (host_order_after(h, nullptr)) return 1;   // ordered after the caller's earlier work on the legacy default stream, no host sync
    TraceArgs base = make_args(0, (int)h->n, S, K, deg, nullptr, nullptr, means, opacity, ru, rv, normals, features, shs,
                               alpha_min, T_min, back_culling);
    if (launch_pack_records(h, base, h->hs[0])) return 1;
    IRGS_CHECK(cudaEventRecord(h->hev[0], h->hs[0]));
    IRGS_CHECK(cudaStreamWaitEvent(h->hs[1], h->hev[0], 0));
    int64_t done = 0;
    for (int it = 0; done < n_rays; ++it) {
        const int si = it & 1;
        cudaStream_t s = h->hs[si];
        int64_t c = (n_rays - done < chunk) ? n_rays - done : chunk;
        if (it == 0 && n_rays > chunk && chunk >= 4096) c = chunk / 4;   // a short first chunk: its copy is the only one nothing overlaps
        float *st = h->stage[si];
        const int rb = it % irgs_tracer::RING;
        float *d_o = h->ring[rb], *d_d = d_o + 3 * chunk;
        float *d_col = st, *d_nrm = d_col + 3 * chunk, *d_feat = d_nrm + 3 * chunk, *d_dep = d_feat + S * chunk,
              *d_alp = d_dep + chunk;
        int32_t *d_cnt = reinterpret_cast<int32_t *>(d_alp + chunk);
        int32_t *d_hits = d_cnt + chunk;
        float *d_go = reinterpret_cast<float *>(d_hits + (int64_t)hit_cap * chunk), *d_gd = d_go + 3 * chunk;
        // copy stream: wait until the chunk that used this ring buffer RING chunks ago is done with it, then copy
        if (it >= irgs_tracer::RING) IRGS_CHECK(cudaStreamWaitEvent(h->hcopy, h->ring_free[rb], 0));
        IRGS_CHECK(cudaMemcpyAsync(d_o, rays_o_host + 3 * done, sizeof(float) * 3 * c, cudaMemcpyHostToDevice, h->hcopy));
        IRGS_CHECK(cudaMemcpyAsync(d_d, rays_d_host + 3 * done, sizeof(float) * 3 * c, cudaMemcpyHostToDevice, h->hcopy));
        IRGS_CHECK(cudaEventRecord(h->ring_full[rb], h->hcopy));
        IRGS_CHECK(cudaStreamWaitEvent(s, h->ring_full[rb], 0));
        TraceArgs a = base;
        a.n_rays = c; a.rays_o = d_o; a.rays_d = d_d;
        a.color = d_col; a.normal = d_nrm; a.feature = d_feat; a.depth = d_dep; a.alpha = d_alp;
        if (with_backward) { a.hit_count = d_cnt; a.hits = d_hits; a.hit_cap = hit_cap; }
        int rc = launch_trace_forward(h, a, s);   // each stream has its own work counter and scratch (slot_for)
        if (!rc && with_backward) {
            a.gC = gC; a.gN = gN; a.gF = gF; a.gD = gD; a.gO = gO; a.gout_period = gout_period; a.gout_offset = done;
            a.g_rays_o = d_go; a.g_rays_d = d_gd; a.grad_fused = grad_fused; a.grad_features = grad_features;
            rc = launch_trace_backward(h, a, s);
        }
        if (rc) return 1;
        IRGS_CHECK(cudaEventRecord(h->ring_free[rb], s));   // the rays of this chunk have been consumed
        if (out_color_host) IRGS_CHECK(cudaMemcpyAsync(out_color_host + 3 * done, d_col, sizeof(float) * 3 * c, cudaMemcpyDeviceToHost, s));
        if (out_normal_host) IRGS_CHECK(cudaMemcpyAsync(out_normal_host + 3 * done, d_nrm, sizeof(float) * 3 * c, cudaMemcpyDeviceToHost, s));
        if (out_feature_host && S > 0) IRGS_CHECK(cudaMemcpyAsync(out_feature_host + S * done, d_feat, sizeof(float) * S * c, cudaMemcpyDeviceToHost, s));
        if (out_depth_host) IRGS_CHECK(cudaMemcpyAsync(out_depth_host + done, d_dep, sizeof(float) * c, cudaMemcpyDeviceToHost, s));
        if (out_alpha_host) IRGS_CHECK(cudaMemcpyAsync(out_alpha_host + done, d_alp, sizeof(float) * c, cudaMemcpyDeviceToHost, s));
        if (with_backward && g_rays_o_host) IRGS_CHECK(cudaMemcpyAsync(g_rays_o_host + 3 * done, d_go, sizeof(float) * 3 * c, cudaMemcpyDeviceToHost, s));
        if (with_backward && g_rays_d_host) IRGS_CHECK(cudaMemcpyAsync(g_rays_d_host + 3 * done, d_gd, sizeof(float) * 3 * c, cudaMemcpyDeviceToHost, s));
        done += c;
    }
    IRGS_CHECK(cudaStreamSynchronize(h->hcopy));
    IRGS_CHECK(cudaStreamSynchronize(h->hs[0]));
    IRGS_CHECK(cudaStreamSynchronize(h->hs[1]));
    return 0;
}

int irgs_trace_forward_host(irgs_tracer_t *h, int64_t n_rays, int S, int K, int deg, const float *rays_o_host,
                            const float *rays_d_host, const float *means, const float *opacity, const float *ru,
                            const float *rv, const float *normals, const float *features, const float *shs,
                            float *out_color_host, float *out_normal_host, float *out_feature_host,
                            float *out_depth_host, float *out_alpha_host, float alpha_min, float T_min,
                            int back_culling, int64_t chunk_rays) {
    return trace_host_impl(h, false, n_rays, S, K, deg, rays_o_host, rays_d_host, means, opacity, ru, rv, normals,
                           features, shs, nullptr, nullptr, nullptr, nullptr, nullptr, 0, out_color_host,
                           out_normal_host, out_feature_host, out_depth_host, out_alpha_host, nullptr, nullptr, nullptr,
                           nullptr, alpha_min, T_min, back_culling, chunk_rays);
}

int irgs_trace_fwd_bwd_host(irgs_tracer_t *h, int64_t n_rays, int S, int K, int deg, const float *rays_o_host,
                            const float *rays_d_host, const float *means, const float *opacity, const float *ru,
                            const float *rv, const float *normals, const float *features, const float *shs,
                            const float *gout_color, const float *gout_normal, const float *gout_feature,
                            const float *gout_depth, const float *gout_alpha, int64_t gout_period,
                            float *out_alpha_host, float *grad_rays_o_host, float *grad_rays_d_host,
                            float *grad_fused, float *grad_features, float alpha_min, float T_min, int back_culling,
                            int64_t chunk_rays) {
    return trace_host_impl(h, true, n_rays, S, K, deg, rays_o_host, rays_d_host, means, opacity, ru, rv, normals,
                           features, shs, gout_color, gout_normal, gout_feature, gout_depth, gout_alpha, gout_period,
                           nullptr, nullptr, nullptr, nullptr, out_alpha_host, grad_rays_o_host, grad_rays_d_host,
                           grad_fused, grad_features, alpha_min, T_min, back_culling, chunk_rays);
}

// Forward + backward on generated incident rays with the per-point inputs in HOST memory: 28 bytes per shading point travel
// host -> device (instead of 24 bytes per ray), 24 bytes per point of gradients travel back.  Chunks of `chunk_points` points
// alternate between the two internal compute streams.
int irgs_trace_fwd_bwd_incident_host(irgs_tracer_t *h, const irgs_incident_t *gen_host, int S, int K, int deg,
                                     const float *means, const float *opacity, const float *ru, const float *rv,
                                     const float *normals, const float *features, const float *shs, const float *gC,
                                     const float *gN, const float *gF, const float *gD, const float *gO, int64_t gout_period,
                                     float *out_alpha_host, float *grad_position_host, float *grad_normal_host,
                                     float *grad_fused, float *grad_features, float alpha_min, float T_min, int back_culling,
                                     int64_t chunk_points, void *stream) {
    if (validate_incident(gen_host, false)) return 1;
    const int64_t P = gen_host->n_points;
    const int NS = gen_host->sample_num;
    const int hit_cap = IRGS_HOST_HIT_CAP;
    if (validate_trace(h, P * NS, S, K, deg, hit_cap)) return 1;
    if (P == 0) return 0;
    if (!grad_fused) return fail("grad_fused must not be null");
    if (S > 0 && !grad_features) return fail("grad_features must not be null when S > 0");
    if (gout_period <= 0) return fail("gout_period must be positive");
    if (chunk_points <= 0) chunk_points = std::max<int64_t>(1, ((int64_t)1 << 22) / NS);
    if (chunk_points > P) chunk_points = P;
    if (chunk_points * NS >= ((int64_t)1 << 31)) return fail("chunk_points * sample_num must be below 2^31");
    DeviceGuard guard(h->device);
    const int64_t chunk = chunk_points * NS;   // rays per chunk
    const int64_t per_ray = 3 + 3 + S + 1 + 1 + 1 + hit_cap + 6;
    if (host_prepare(h, per_ray * chunk, 0)) return 1;
    // per-point inputs (position, normals, azimuth) and outputs (dL/dposition, dL/dnormal) of the whole batch: 13 floats a point
    if (13 * P > h->pt_floats) {
        if (h->pt_buf) { IRGS_CHECK(cudaDeviceSynchronize()); cudaFree(h->pt_buf); }
        h->pt_buf = nullptr;
        IRGS_CHECK(cudaMalloc(&h->pt_buf, sizeof(float) * (size_t)(13 * P)));
        h->pt_floats = 13 * P;
    }
    float *d_pos = h->pt_buf, *d_nrm = d_pos + 3 * P, *d_az = d_nrm + 3 * P, *d_gpos = d_az + P, *d_gnrm = d_gpos + 3 * P;
    if (host_order_after(h, (cudaStream_t)stream)) return 1;
    IRGS_CHECK(cudaMemcpyAsync(d_pos, gen_host->position, sizeof(float) * 3 * P, cudaMemcpyHostToDevice, h->hcopy));
    IRGS_CHECK(cudaMemcpyAsync(d_nrm, gen_host->normals, sizeof(float) * 3 * P, cudaMemcpyHostToDevice, h->hcopy));
    if (gen_host->azimuth) IRGS_CHECK(cudaMemcpyAsync(d_az, gen_host->azimuth, sizeof(float) * P, cudaMemcpyHostToDevice, h->hcopy));
    IRGS_CHECK(cudaEventRecord(h->ring_full[0], h->hcopy));
    TraceArgs base = make_args(0, (int)h->n, S, K, deg, nullptr, nullptr, means, opacity, ru, rv, normals, features, shs,
                               alpha_min, T_min, back_culling);
    if (launch_pack_records(h, base, h->hs[0])) return 1;
    IRGS_CHECK(cudaEventRecord(h->hev[1], h->hs[0]));
    IRGS_CHECK(cudaStreamWaitEvent(h->hs[1], h->hev[1], 0));
    for (int i = 0; i < 2; ++i) IRGS_CHECK(cudaStreamWaitEvent(h->hs[i], h->ring_full[0], 0));
    int it = 0;
    for (int64_t p0 = 0; p0 < P; p0 += chunk_points, ++it) {
        cudaStream_t s = h->hs[it & 1];
        const int64_t np = std::min(chunk_points, P - p0), c = np * NS;
        float *st = h->stage[it & 1];
        float *d_col = st, *d_nrm_o = d_col + 3 * chunk, *d_feat = d_nrm_o + 3 * chunk, *d_dep = d_feat + S * chunk,
              *d_alp = d_dep + chunk;
        int32_t *d_cnt = reinterpret_cast<int32_t *>(d_alp + chunk);
        int32_t *d_hits = d_cnt + chunk;
        float *d_go = reinterpret_cast<float *>(d_hits + (int64_t)hit_cap * chunk), *d_gd = d_go + 3 * chunk;
        TraceArgs a = base;
        a.n_rays = c;
        a.gen_pos = d_pos + 3 * p0; a.gen_nrm = d_nrm + 3 * p0; a.gen_azim = gen_host->azimuth ? d_az + p0 : nullptr;
        a.gen_S = NS; a.gen_tmin = gen_host->t_min; a.gen_P = np;
        a.color = d_col; a.normal = d_nrm_o; a.feature = d_feat; a.depth = d_dep; a.alpha = d_alp;
        a.hit_count = d_cnt; a.hits = d_hits; a.hit_cap = hit_cap;
        if (launch_incident_prepare(h, a, s) || launch_trace_forward(h, a, s)) return 1;
        a.gC = gC; a.gN = gN; a.gF = gF; a.gD = gD; a.gO = gO; a.gout_period = gout_period; a.gout_offset = p0 * NS;
        a.g_rays_o = d_go; a.g_rays_d = d_gd; a.grad_fused = grad_fused; a.grad_features = grad_features;
        if (launch_trace_backward(h, a, s)) return 1;
        if (launch_incident_backward(a.gen_pos, a.gen_nrm, a.gen_azim, np, NS, a.gen_tmin, d_go, d_gd, d_gpos + 3 * p0,
                                     d_gnrm + 3 * p0, s))
            return 1;
        if (out_alpha_host) IRGS_CHECK(cudaMemcpyAsync(out_alpha_host + p0 * NS, d_alp, sizeof(float) * c, cudaMemcpyDeviceToHost, s));
        if (grad_position_host) IRGS_CHECK(cudaMemcpyAsync(grad_position_host + 3 * p0, d_gpos + 3 * p0, sizeof(float) * 3 * np, cudaMemcpyDeviceToHost, s));
        if (grad_normal_host) IRGS_CHECK(cudaMemcpyAsync(grad_normal_host + 3 * p0, d_gnrm + 3 * p0, sizeof(float) * 3 * np, cudaMemcpyDeviceToHost, s));
    }
    // the caller's stream continues after both compute streams (grad_fused is complete for whatever it launches next); the
    // host waits because the host outputs must be readable on return
    for (int i = 0; i < 2; ++i) {
        IRGS_CHECK(cudaEventRecord(h->ring_free[i], h->hs[i]));
        IRGS_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, h->ring_free[i], 0));
    }
    IRGS_CHECK(cudaStreamSynchronize(h->hs[0]));
    IRGS_CHECK(cudaStreamSynchronize(h->hs[1]));
    return 0;
}

int64_t irgs_stride_multiplier(int64_t n_rays) { return stride_multiplier(n_rays); }

int64_t irgs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void irgs_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

int irgs_set_option(irgs_tracer_t *h, const char *name, int64_t value) {
    if (!h || !name) return fail("null argument");
    if (strcmp(name, "sort_rays_min") == 0) {
        h->sort_rays_min = value < 0 ? 0 : (value > INT32_MAX ? INT32_MAX : (int)value);
        return 0;
    }
    if (strcmp(name, "bwd_carveout_pct") == 0) {
        h->bwd_carveout_pct = value < 0 ? -1 : (value > 100 ? 100 : (int)value);
        return 0;
    }
    if (strcmp(name, "fwd_blocks_per_sm") == 0) {
        h->fwd_blocks_per_sm = value < 0 ? 0 : (value > 32 ? 32 : (int)value);
        return 0;
    }
    if (strcmp(name, "smem_carveout_pct") == 0) {
        h->carveout_pct = value < 0 ? -1 : (value > 100 ? 100 : (int)value);
        return 0;
    }
    if (strcmp(name, "stride_rays_max") == 0) {
        h->stride_rays_max = value < 0 ? 0 : value;
        return 0;
    }
    if (strcmp(name, "slot") == 0) return 0;   // accepted for compatibility: slots follow the stream of each call now
    if (strcmp(name, "builder") == 0) {   // takes effect at the next build_bvh / build_from_surfels
        h->builder = value == 1 ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "wide_fold") == 0) {   // takes effect at the next build_bvh / build_from_surfels
        h->wide_fold = value == 1 ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "contiguous_outputs") == 0) {
        h->contiguous_outputs = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "skip_next_pack") == 0) {
        h->skip_next_pack = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "gen_in_kernel") == 0) {
        h->gen_in_kernel = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "color_cache") == 0) {
        h->color_cache = value < 0 ? 0 : (value > 96 ? 96 : (int)value);
        return 0;
    }
    if (strcmp(name, "bwd_mode") == 0) {
        h->bwd_mode = (value == 1 || value == 2) ? (int)value : 0;
        return 0;
    }
    return fail("unknown option");
}

int64_t irgs_get_info(irgs_tracer_t *h, const char *name) {
    if (!h || !name) return -1;
    if (strcmp(name, "tree_depth") == 0) return h->tree_depth;        // 0: Karras tree (depth <= 62 by construction)
    if (strcmp(name, "ploc_iterations") == 0) return h->ploc_iterations;
    if (strcmp(name, "n_slots") == 0) return h->n_slots;
    if (strcmp(name, "n_surfels") == 0) return h->built ? h->n : 0;
    if (strcmp(name, "pack_epoch") == 0) return h->pack_epoch;
    if (strcmp(name, "color_cache_bytes") == 0) {
        int64_t b = 0;
        for (int i = 0; i < irgs_tracer::MAX_SLOTS; ++i) b += 4 * h->hit_rgb_floats[i];
        return b;
    }
    if (strcmp(name, "comp_stats") == 0 || strcmp(name, "full_rows") == 0) {
        // statistics build: "comp_stats" = sum over the packed compositing rounds of (longest segment << 32 | candidates),
        // "full_rows" = rounds << 32 | full-row sorts
        DeviceGuard guard(h->device);
        unsigned long long v = 0;
        if (cudaMemcpy(&v, h->stats + (strcmp(name, "comp_stats") == 0 ? 6 : 7), sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return (int64_t)v;
    }
    if (strcmp(name, "grazing_pairs") == 0 || strcmp(name, "grazing_pairs_compositing") == 0) {
        // statistics of the last forward with irgs_set_stats(h, 1): ray / surfel pairs with |n.d| < 1e-3 that cross the surfel's
        // support geometrically (dropped by the hit test; the reference evaluates them with its clamped depth), and how many of
        // them that clamped evaluation would have composited
        DeviceGuard guard(h->device);
        unsigned long long v = 0;
        if (cudaMemcpy(&v, h->stats + (strcmp(name, "grazing_pairs") == 0 ? 4 : 5), sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return (int64_t)v;
    }
    return -1;
}

int irgs_set_stats(irgs_tracer_t *h, int enable) {
    if (!h) return fail("null tracer handle");
    h->stats_enabled = enable ? 1 : 0;
    return 0;
}

int irgs_get_stats(irgs_tracer_t *h, int64_t out[4]) {
    if (!h) return fail("null tracer handle");
    DeviceGuard guard(h->device);
    unsigned long long tmp[4];
    IRGS_CHECK(cudaMemcpy(tmp, h->stats, sizeof tmp, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) out[i] = (int64_t)tmp[i];
    return 0;
}

}  // extern "C"
