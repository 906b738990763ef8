/*
 * irgs_b200.h -- C ABI of the B200-native differentiable surfel ray tracer.
 *
 * This is the drop-in boundary for IRGS's `surfel_tracer` extension: every entry point below replaces one method
 * of the reference's pybind11 class `surfel_tracer::GaussianTracer`
 * (/root/reference/submodules/surfel_tracer/src/bindings.cu:24-116) or of the `TriangleBvh` it forwards to
 * (src/bvh.cu:163-252).  Plain pointers and sizes only; no torch types.  All `const float*` / `float*` arguments
 * are DEVICE pointers to contiguous row-major float32 arrays unless the name ends in `_host`; `stream` is a
 * `cudaStream_t` passed as `void*` (0 = legacy default stream).  Trace, refit and unpack calls are asynchronous with respect to
 * the host and ordered on `stream`, exactly like the reference, which launches on at::cuda::getCurrentCUDAStream()
 * (bindings.cu:32,38,49,66,82).  Exceptions, all outside the per-iteration path: a BUILD (irgs_build_*) synchronises `stream`
 * once per clustering iteration (the host sizes the next grid from the cluster count, ~55 times at 300k surfels), the `_host`
 * entry points return when their host outputs are complete, and the first call that needs a larger scratch block on a stream
 * synchronises that stream to reallocate it.
 *
 * Error convention: 0 = success; non-zero = failure, irgs_last_error() returns a thread-local message.  The
 * reference throws std::runtime_error through pybind (gpu_memory.h:50-55); the Python layer maps a non-zero status
 * to RuntimeError.  Unlike the reference (which validates nothing, bindings.cu:42-59) shape limits are checked.
 *
 * Array layouts (SURVEY.md 8a'):  rays_o, rays_d [R,3];  means3D, ru, rv, normals [N,3];  opacity [N] (the
 * reference's [N,1]);  features [N,S], S <= 12;  shs [N,K,3], K >= (deg+1)^2;  outputs color, normal [R,3],
 * feature [R,S], depth, alpha [R].
 */
#ifndef IRGS_B200_H
#define IRGS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IRGS_MAX_FEATURES 12      /* MAX_FEATURE_SIZE, src/optix/auxiliary.h:12 */
#define IRGS_T_SCENE_MAX 100.0f   /* T_SCENE_MAX,      src/optix/auxiliary.h:11 */
#define IRGS_GRAD_STRIDE 64       /* floats per surfel in the fused gradient buffer, see irgs_trace_backward */

typedef struct irgs_tracer irgs_tracer_t;

const char *irgs_last_error(void);
int irgs_version(void);

/* create_gaussiantracer() (bindings.cu:101-103) / TriangleBvh ctor+dtor (bvh.cu:165-178).  One handle per device
 * (per rank); unlike the reference's process-global OptiX context the handle owns all of its state. */
int irgs_tracer_create(irgs_tracer_t **out, int device);
int irgs_tracer_destroy(irgs_tracer_t *h);

/* GaussianTracer::build_bvh (bindings.cu:30-34 -> bvh.cu:231-234 -> optix::Gas ctor :71-111, optixAccelBuild BUILD).
 * Takes the reference caller's proxy mesh vertices_b [verts_per_surfel*N, 3] (scene/gaussian_model.py:712-723:
 * 12 icosahedron vertices per surfel) instead of the gathered 20N-triangle soup: the bound of surfel g is the AABB
 * of its vertices (NaN vertices -- opacity < alpha_min -- give an empty bound that no ray can hit).
 * Morton code (30 bit) of the AABB centroid -> radix sort -> Karras hierarchy -> bottom-up bounds. */
int irgs_build_from_proxy(irgs_tracer_t *h, const float *vertices_b, int64_t n_surfels, int verts_per_surfel, void *stream);
/* GaussianTracer::update_bvh (bindings.cu:36-40 -> bvh.cu:236-239 -> Gas::update :113-147, optixAccelBuild UPDATE):
 * topology frozen, bounds recomputed bottom-up.  n_surfels must equal the built count. */
int irgs_refit_from_proxy(irgs_tracer_t *h, const float *vertices_b, int64_t n_surfels, int verts_per_surfel, void *stream);

/* Native variants of the two calls above that take the surfel parameters the tracer is given anyway and derive
 * the exact elliptical bound { alpha >= alpha_min } analytically (no 12N-vertex proxy mesh needed). */
int irgs_build_from_surfels(irgs_tracer_t *h, const float *means3D, const float *opacity, const float *ru,
                            const float *rv, const float *normals, int64_t n_surfels, float alpha_min, void *stream);
int irgs_refit_from_surfels(irgs_tracer_t *h, const float *means3D, const float *opacity, const float *ru,
                            const float *rv, const float *normals, int64_t n_surfels, float alpha_min, void *stream);

/* Introspection for tests: number of surfels in the structure; copy of the per-surfel bounds [N,6] (lo, hi) in
 * surfel order and of the root bound [6] to DEVICE buffers. */
int64_t irgs_num_surfels(const irgs_tracer_t *h);
int irgs_get_bounds(irgs_tracer_t *h, float *surfel_bounds /* [N,6] or NULL */, float *root_bound /* [6] or NULL */, void *stream);

/* GaussianTracer::intersection_test (bindings.cu:61-71 -> gaussiantrace_intersection_test.cu:12-35): out[r] = 1 if
 * ray r crosses any surfel bound support within (FLT_EPSILON, 100), else 0.  Kept for API completeness; the native
 * trace does not need the mask/compaction pre-pass of raytracer.py:103-112. */
int irgs_intersection_test(irgs_tracer_t *h, int64_t n_rays, const float *rays_o, const float *rays_d,
                           const float *means3D, const float *opacity, const float *ru, const float *rv,
                           const float *normals, float alpha_min, uint8_t *out, void *stream);

/* GaussianTracer::trace_forward (bindings.cu:42-59 -> gaussiantrace_forward.cu:12-141).
 * Outputs are OVERWRITTEN for all R rays (rays that hit nothing get exact zeros).
 *   out_hit_count [R] int32 or NULL : number of composited hits per ray (the north-star's "hit counts").
 *   out_hits [R, hit_cap] int32 or NULL : surfel ids in compositing order (first min(count, hit_cap) entries of each
 *     row are valid); saved for irgs_trace_backward's replay.  hit_cap must be a multiple of 4. */
int irgs_trace_forward(irgs_tracer_t *h, int64_t n_rays, int S, int K, int deg, const float *rays_o,
                       const float *rays_d, const float *means3D, const float *opacity, const float *ru,
                       const float *rv, const float *normals, const float *features, const float *shs,
                       float *out_color, float *out_normal, float *out_feature, float *out_depth, float *out_alpha,
                       int32_t *out_hit_count, int32_t *out_hits, int hit_cap, float alpha_min,
                       float transmittance_min, int back_culling, void *stream);

/* GaussianTracer::trace_backward (bindings.cu:73-94 -> gaussiantrace_backward.cu:11-171).
 * color..alpha are the forward outputs, gout_* the incoming gradients.  grad_rays_o/grad_rays_d [R,3] are
 * OVERWRITTEN; per-surfel gradients are ACCUMULATED (atomically) into
 *   grad_fused [N, IRGS_GRAD_STRIDE]: floats 0-2 d/dmeans3D, 3 d/dopacity, 4-6 d/dru, 7-9 d/drv, 10-12 d/dnormals,
 *       13-15 unused, 16.. d/dshs[k][c] at 16 + 3k + c for k < 16 (coefficients k >= 16 never receive gradient);
 *   grad_features [N,S] (may be NULL when S == 0).
 * One buffer so that the only collective of the multi-GPU path is a single all-reduce over it.
 * If hit_count/hits are given, rays with hit_count <= hit_cap are replayed from the saved list without touching
 * the acceleration structure; the others (and all rays when hits == NULL) re-trace like the reference does. */
int irgs_trace_backward(irgs_tracer_t *h, int64_t n_rays, int S, int K, int deg, const float *rays_o,
                        const float *rays_d, const float *means3D, const float *opacity, const float *ru,
                        const float *rv, const float *normals, const float *features, const float *shs,
                        const float *color, const float *normal, const float *feature, const float *depth,
                        const float *alpha, const int32_t *hit_count, const int32_t *hits, int hit_cap,
                        const float *gout_color, const float *gout_normal, const float *gout_feature,
                        const float *gout_depth, const float *gout_alpha, float *grad_rays_o, float *grad_rays_d,
                        float *grad_fused, float *grad_features, float alpha_min, float transmittance_min,
                        int back_culling, void *stream);

/* De-interleave the fused gradient buffer into the nine-tensor form raytracer.py:27-66 returns.
 * grad_shs [N,K,3] is fully overwritten (zeros for k >= 16). */
int irgs_unpack_grads(const float *grad_fused, int64_t n_surfels, int K, float *grad_means3D, float *grad_opacity,
                      float *grad_ru, float *grad_rv, float *grad_normals, float *grad_shs, void *stream);

/* End-to-end entry points on HOST buffers (pinned or pageable): rays are copied host->device in chunks on internal
 * streams, traced, and results copied device->host, overlapping copies with the kernels.  Surfel arrays and the
 * incoming-gradient arrays stay DEVICE pointers (they live on the GPU in IRGS); the internal streams start after everything the
 * caller has already submitted to the legacy default stream (an event, no device-wide synchronisation).  gout_* are periodic device arrays
 * with gout_period rows (ray r of the whole batch uses row r % gout_period).  Host outputs may be NULL to skip the copy back.
 * irgs_trace_fwd_bwd_host runs forward + backward per chunk and accumulates into grad_fused (device). */
int irgs_trace_forward_host(irgs_tracer_t *h, int64_t n_rays, int S, int K, int deg, const float *rays_o_host,
                            const float *rays_d_host, const float *means3D, const float *opacity, const float *ru,
                            const float *rv, const float *normals, const float *features, const float *shs,
                            float *out_color_host, float *out_normal_host, float *out_feature_host,
                            float *out_depth_host, float *out_alpha_host, float alpha_min, float transmittance_min,
                            int back_culling, int64_t chunk_rays);
int irgs_trace_fwd_bwd_host(irgs_tracer_t *h, int64_t n_rays, int S, int K, int deg, const float *rays_o_host,
                            const float *rays_d_host, const float *means3D, const float *opacity, const float *ru,
                            const float *rv, const float *normals, const float *features, const float *shs,
                            const float *gout_color, const float *gout_normal, const float *gout_feature,
                            const float *gout_depth, const float *gout_alpha, int64_t gout_period,
                            float *out_alpha_host, float *grad_rays_o_host, float *grad_rays_d_host,
                            float *grad_fused, float *grad_features, float alpha_min, float transmittance_min,
                            int back_culling, int64_t chunk_rays);

/* Launch counter: number of kernels this library has launched since the last irgs_reset_launch_count()
 * (bench.py's "gpu_launches"). */
int64_t irgs_launch_count(void);
void irgs_reset_launch_count(void);

/* ---- Fused incident-ray generation -------------------------------------------------------------
 * Replaces the caller-side ray construction of /root/reference/gaussian_renderer/__init__.py:324-332,376
 * (sample_incident_rays; `position.unsqueeze(1) + incident_dirs * light_t_min`) and utils/graphics_utils.py:19-47,133-165
 * (fibonacci_sphere_sampling, rotation_between_z): the n_points * sample_num rays are generated inside the tracing
 * kernels from one (position, normal, azimuth) triple per shading point instead of being read from a [P, S, 3] pair of
 * arrays.  Ray index = point * sample_num + sample, outputs are laid out [P, S, ...].  azimuth[p] is the per-point
 * `rand * 2 pi` of the reference's training mode; NULL = evaluation mode (no random rotation).  All device pointers. */
typedef struct {
    const float *position;   /* [P,3] shading points */
    const float *normals;    /* [P,3] unit shading normals */
    const float *azimuth;    /* [P] or NULL */
    int64_t n_points;
    int32_t sample_num;      /* S, pipe.diffuse_sample_num */
    float t_min;             /* pipe.light_t_min */
} irgs_incident_t;

/* The generated rays themselves, [P,S,3] each (either output may be NULL): for shading code that needs the directions. */
int irgs_incident_rays(const irgs_incident_t *gen, float *rays_o, float *rays_d, void *stream);

/* irgs_trace_forward / irgs_trace_backward on generated rays.  The backward additionally returns dL/dposition [P,3] and
 * dL/dnormal [P,3] of the shading points (through rotation_between_z; zero on its constant -identity branch), reduced over the
 * S samples of each point from the per-ray gradients, which are left in the two [P*S,3] scratch arrays. */
int irgs_trace_forward_incident(irgs_tracer_t *h, const irgs_incident_t *gen, int S, int K, int deg, const float *means,
                                const float *opacity, const float *ru, const float *rv, const float *normals,
                                const float *features, const float *shs, float *out_color, float *out_normal,
                                float *out_feature, float *out_depth, float *out_alpha, int32_t *out_hit_count,
                                int32_t *out_hits, int hit_cap, float alpha_min, float T_min, int back_culling,
                                void *stream);
int irgs_trace_backward_incident(irgs_tracer_t *h, const irgs_incident_t *gen, int S, int K, int deg, const float *means,
                                 const float *opacity, const float *ru, const float *rv, const float *normals,
                                 const float *features, const float *shs, const float *color, const float *normal,
                                 const float *feature, const float *depth, const float *alpha, const int32_t *hit_count,
                                 const int32_t *hits, int hit_cap, const float *gout_color, const float *gout_normal,
                                 const float *gout_feature, const float *gout_depth, const float *gout_alpha,
                                 float *scratch_grad_rays_o, float *scratch_grad_rays_d, float *grad_position,
                                 float *grad_normal_pt, float *grad_fused, float *grad_features, float alpha_min, float T_min,
                                 int back_culling, void *stream);

/* irgs_trace_fwd_bwd_host for generated incident rays with the per-point inputs on the HOST: gen_host->position / normals /
 * azimuth are host pointers (pinned or pageable); 28 bytes per shading point go host -> device instead of 24 bytes per ray.
 * Forward + backward run chunk by chunk (chunk_points shading points each, 0 = 2^22 rays' worth) on two internal streams that
 * start after everything already submitted to `stream`; `stream` continues after them.  Host outputs (any may be NULL):
 * out_alpha_host [P*S], grad_position_host [P,3], grad_normal_host [P,3]; per-surfel gradients accumulate into the DEVICE
 * buffers grad_fused / grad_features; gout_* as in irgs_trace_fwd_bwd_host.  Returns when the host outputs are complete. */
int irgs_trace_fwd_bwd_incident_host(irgs_tracer_t *h, const irgs_incident_t *gen_host, int S, int K, int deg,
                                     const float *means3D, const float *opacity, const float *ru, const float *rv,
                                     const float *normals, const float *features, const float *shs, const float *gout_color,
                                     const float *gout_normal, const float *gout_feature, const float *gout_depth,
                                     const float *gout_alpha, int64_t gout_period, float *out_alpha_host,
                                     float *grad_position_host, float *grad_normal_host, float *grad_fused,
                                     float *grad_features, float alpha_min, float transmittance_min, int back_culling,
                                     int64_t chunk_points, void *stream);

/* ---- Shading epilogue around the incident-ray trace (SURVEY.md 8f rank 1 + rank 3) ----------------------------------
 * Replaces the element-wise torch code of /root/reference/gaussian_renderer/__init__.py:334-415 (rendering_equation, the
 * diffuse_sample_num > 0 / light_sample_num == 0 / non-relight path) together with :417-457 (GGX_specular) and the
 * environment lookup scene/light.py:287-297,315 (EnvLight.__call__(mode='pure_env'): lat-long uv, nvdiffrast
 * texture(filter 'linear', boundary 'wrap'), activation, clamp_min(0)).  One kernel per direction: a warp per shading
 * point regenerates the S incident directions of `gen` (the same device function the tracing kernels use), reads the
 * tracer's RAW colour / alpha of those rays and reduces over S. */
typedef struct {
    const float *base;       /* [H,W,3] pre-activation texels (EnvLight.base), device */
    int32_t height, width;
    int32_t activation;      /* 0 none, 1 exp, 2 sigmoid (EnvLight.activation_name) */
    int32_t has_transform;   /* EnvLight.transform set? */
    float transform[9];      /* row-major 3x3: lookup direction = d @ transform^T (light.py:299-300) */
} irgs_envmap_t;

/* The light_sample_num > 0 branch (gaussian_renderer/__init__.py:340-357): Fibonacci samples mixed with directions drawn
 * from the environment map (EnvLight.sample_light_directions, scene/light.py:181-205).  One shade call per kind of sample;
 * every sample is weighted by 1 / clamp_min(p_diffuse / (2 pi) + p_light * light_pdf(dir), 1e-6) with light_pdf =
 * EnvLight.light_pdf (light.py:207-223), and the means run over total_samples = diffuse + light samples, so that the
 * [P,16] rows of the two calls ADD UP to rendering_equation's result.  Pass NULL for pure Fibonacci sampling. */
typedef struct {
    const float *dirs;       /* [P*S,3] explicit unit directions of this call's samples (the light samples), or NULL: generated */
    const float *pdf;        /* [H,W] texel probabilities (EnvLight._pdf after update_pdf), or NULL: weight 2 pi */
    float p_diffuse, p_light;   /* diffuse_sample_num / (diffuse + light), light_sample_num / (diffuse + light) */
    int32_t total_samples;   /* diffuse + light samples: the divisor of the means */
} irgs_shade_sampling_t;

/* trace_color [P*S,3], trace_alpha [P*S]: outputs of irgs_trace_forward_incident for the same `gen`.  saturate_alpha =
 * 1 - transmittance_min applies GaussianModel.trace's normalisation of saturated rays (scene/gaussian_model.py:748-752:
 * colour / alpha and alpha = 1 where alpha >= 1 - transmittance_min); a negative value skips it.  base_color [P,3],
 * roughness [P], viewdirs [P,3] (any length).  gen->position is not used.
 * out [P,16], means over the S samples: 0-2 diffuse, 3-5 specular, 6-8 light_direct, 9 visibility, 10-12 light,
 * 13-15 light_indirect (the keys of rendering_equation's result dict, __init__.py:399-414). */
int irgs_shade_forward(const irgs_incident_t *gen, const irgs_envmap_t *env, const irgs_shade_sampling_t *sampling,
                       const float *base_color, const float *roughness, const float *viewdirs, const float *trace_color,
                       const float *trace_alpha, float saturate_alpha, float *out, void *stream);
/* g_out [P,16] in the layout of `out`.  OVERWRITTEN: g_trace_color [P*S,3], g_trace_alpha [P*S] (the gradients to hand to
 * irgs_trace_backward_incident as gout_color / gout_alpha) and g_point [P,16]: 0-2 dL/dbase_color, 3 dL/droughness,
 * 4-6 dL/dnormal (n_d_i, the GGX normal AND the dependence of the sampled directions on the normal through
 * rotation_between_z; explicit directions are constants), 7-9 dL/dviewdirs.  Texel gradients are ADDED (atomically) into
 * grad_env [H,W,3] (may be NULL). */
int irgs_shade_backward(const irgs_incident_t *gen, const irgs_envmap_t *env, const irgs_shade_sampling_t *sampling,
                        const float *base_color, const float *roughness, const float *viewdirs, const float *trace_color,
                        const float *trace_alpha, float saturate_alpha, const float *g_out, float *g_trace_color,
                        float *g_trace_alpha, float *g_point, float *grad_env, void *stream);
/* EnvLight.__call__(dirs, mode='pure_env') on its own (e.g. gaussian_renderer/__init__.py:244: the background radiance of
 * the primary rays): dirs [n,3] -> out [n,3]; backward: g_dirs [n,3] overwritten (may be NULL), grad_env added to. */
int irgs_env_lookup_forward(const irgs_envmap_t *env, const float *dirs, int64_t n_dirs, float *out, void *stream);
int irgs_env_lookup_backward(const irgs_envmap_t *env, const float *dirs, const float *g_out, int64_t n_dirs, float *g_dirs,
                             float *grad_env, void *stream);

/* ---- Primary pass through the tracer -------------------------------------------------------------
 * The G-buffer the reference takes from its tile rasteriser (gaussian_renderer/__init__.py:121-131: alpha, normal, depth, base
 * colour + roughness as features, SH colour) traced instead: the H x W pinhole rays of scene/cameras.py:87-100 are generated in
 * the kernels from the camera (ray = v * width + u; d = normalize(cam_to_world ((u - W/2 + 0.5) / fx, (v - H/2 + 0.5) / fy, 1)),
 * origin = the camera centre; cam_to_world = world_view_transform[:3,:3] of the reference, row-major), so no ray array exists.
 * Outputs / gradients exactly as irgs_trace_forward / irgs_trace_backward (outputs [H*W, ...]); ray gradients go to the two
 * scratch arrays [H*W,3] (the camera is not a parameter).  irgs_camera_rays writes the generated rays out (tests, callers that
 * need rays_d_hw).  Not reproduced (stated parity risk): the rasteriser's per-tile ray-splat intersection in screen space, its
 * low-pass filter, its median depth and distortion maps. */
typedef struct {
    float origin[3];          /* camera centre, world space */
    float cam_to_world[9];    /* row-major 3x3: d_world = cam_to_world d_camera */
    float fx, fy;             /* focal lengths in pixels */
    int32_t width, height;
} irgs_camera_t;
int irgs_camera_rays(const irgs_camera_t *cam, float *rays_o, float *rays_d, void *stream);
int irgs_trace_forward_camera(irgs_tracer_t *h, const irgs_camera_t *cam, int S, int K, int deg, const float *means3D,
                              const float *opacity, const float *ru, const float *rv, const float *normals,
                              const float *features, const float *shs, float *out_color, float *out_normal, float *out_feature,
                              float *out_depth, float *out_alpha, int32_t *out_hit_count, int32_t *out_hits, int hit_cap,
                              float alpha_min, float transmittance_min, int back_culling, void *stream);
int irgs_trace_backward_camera(irgs_tracer_t *h, const irgs_camera_t *cam, int S, int K, int deg, const float *means3D,
                               const float *opacity, const float *ru, const float *rv, const float *normals,
                               const float *features, const float *shs, const float *color, const float *normal,
                               const float *feature, const float *depth, const float *alpha, const int32_t *hit_count,
                               const int32_t *hits, int hit_cap, const float *gout_color, const float *gout_normal,
                               const float *gout_feature, const float *gout_depth, const float *gout_alpha,
                               float *scratch_grad_rays_o, float *scratch_grad_rays_d, float *grad_fused, float *grad_features,
                               float alpha_min, float transmittance_min, int back_culling, void *stream);

/* ---- Relight branch of rendering_equation (gaussian_renderer/__init__.py:362-381) ---------------------------------------------
 * The hit point of every secondary ray is shaded under the novel environment: irgs_relight_hit turns the tracer's raw normal
 * [R,3] / feature [R,4] = (base colour, roughness) / alpha [R] of rays with directions dirs [R,3] into the arguments of the
 * caller's two environment lookups -- hit_normal [R,3] for mode 'diffuse', reflected [R,3] + roughness [R] for mode 'specular'
 * (cube-map prefilters of scene/light.py:264-328, evaluated by the caller) -- and a packed row pack [R,8];
 * irgs_relight_combine takes the two lookups' results [R,3] and the FG table fg_lut [H,W,2] (nvdiffrast 'linear' / 'clamp'
 * lookup at (N.V, roughness)) and writes local_light [R,3] = (base * diffuse + specular * (f0 * FG.x + FG.y)) * alpha (zeros when
 * wo_indirect) and alpha_out [R], which irgs_shade_forward takes as (trace_color, trace_alpha) with saturate_alpha < 0.
 * saturate_alpha = 1 - transmittance_min (GaussianModel.trace's normalisation), negative to skip.  Forward only (the reference
 * runs this branch under torch.no_grad()). */
int irgs_relight_hit(int64_t n_rays, const float *dirs, const float *trace_normal, const float *trace_feature,
                     const float *trace_alpha, float saturate_alpha, float *hit_normal, float *reflected, float *roughness,
                     float *pack, void *stream);
int irgs_relight_combine(int64_t n_rays, const float *pack, const float *env_diffuse, const float *env_specular,
                         const float *fg_lut, int lut_height, int lut_width, float f0, int wo_indirect, float *local_light,
                         float *alpha_out, void *stream);

/* ---- Parameter-level entry -----------------------------------------------------------------------
 * The caller glue of /root/reference/scene/gaussian_model.py:733-756 as kernels: surfel PARAMETERS (means [N,3], activated
 * scales [N,2], quaternions (w,x,y,z) [N,4], not necessarily normalised) -> the tracer's ru / rv / normals (utils/general_utils.py
 * :78-99 build_rotation, :135-146 safe_normalize + flip_align_view towards camera_center_host, a HOST float[3] or NULL = no flip),
 * the gradients of the parameters from the fused [N,64] buffer (chain rule through ru = R[:,0]/s_u, rv = R[:,1]/s_v,
 * normals = +-R[:,2]/|R[:,2]| and the quaternion normalisation; means / opacity / SH rows are copied out), and the
 * normalisation of saturated rays (alpha >= threshold = 1 - T_min: accumulations divided by alpha, alpha set to 1) with its
 * backward (g_* arrive for the normalised outputs and are rewritten in place as gradients of the raw ones). */
int irgs_surfel_frames(const float *means3D, const float *scales, const float *rotations, const float *camera_center_host,
                       int64_t n_surfels, float *ru, float *rv, float *normals, void *stream);
int irgs_unpack_grads_params(const float *grad_fused, int64_t n_surfels, int K, const float *means3D, const float *scales,
                             const float *rotations, const float *camera_center_host, float *grad_means3D, float *grad_opacity,
                             float *grad_scales, float *grad_rotations, float *grad_shs, void *stream);
int irgs_normalize_outputs(int64_t n_rays, int S, float threshold, const float *color, const float *normal, const float *feature,
                           const float *depth, const float *alpha, float *out_color, float *out_normal, float *out_feature,
                           float *out_depth, float *out_alpha, void *stream);
int irgs_normalize_outputs_backward(int64_t n_rays, int S, float threshold, const float *color, const float *normal,
                                    const float *feature, const float *depth, const float *alpha, float *g_color,
                                    float *g_normal, float *g_feature, float *g_depth, float *g_alpha, void *stream);

/* Tuning knobs (never change results).  "sort_rays_min": forward calls with at least this many rays process them in
 * a coherence-sorted order (origin cell, direction bin); 0 disables the sort.  "bwd_mode": 0 (default) replays the saved
 * hit lists one hit per lane (segmented warp scans, 256-byte row reductions), 1 one ray per thread.
 * "stride_rays_max": forward calls with at most this many rays (default and upper limit 2^19) start their rays in a stride
 * order instead of the caller's order (latency of small calls; 0 disables).
 * Calls issued on different streams at the same time are safe: every stream that launches on a handle owns a slot (work counter,
 * candidate scratch) looked up from the call's own `stream`, so a backward uses the slot of the stream it runs on ("slot" is
 * still accepted and ignored).  A handle is not thread-safe beyond that: do not build / refit while traces are in flight.
 * "builder": 0 (default) PLOC clustering over the Morton order, 1 Karras LBVH; takes effect at the next build.
 * "wide_fold": 0 (default) the 4-wide traversal nodes are a greedy collapse of the binary tree (the internal child with the largest
 * surface area is expanded while a node has a free slot), 1 the fixed fold of every other level; takes effect at the next build.
 * "fwd_blocks_per_sm": resident blocks per SM of the forward kernel's persistent grid (0 = what the occupancy allows; experiments).
 * "contiguous_outputs": 1 = the caller promises that whenever a forward call's output arrays are back to back in memory (color,
 * normal, feature, depth, alpha, hit_count) they are views of ONE allocation; they are then zero-filled with a single memset
 * instead of six (default 0: arrays that merely happen to be adjacent may belong to different allocations).
 * "skip_next_pack": 1 makes the next irgs_trace_backward* call on this handle reuse the packed surfel records instead of packing
 * them again -- valid when irgs_get_info("pack_epoch") still has the value it had right after the forward of the same arrays
 * (every pack, build and refit bumps it); the Python layer does exactly that.
 * "gen_in_kernel": 0 (default) incident / camera rays of a forward call are written to an internal scratch block by a small
 * kernel and read back by the forward kernel (measured faster: DRAM is idle, the persistent walk is not), 1 generates them
 * inside the forward kernel (24 B per ray of a call less memory).  The backward always regenerates them in its kernels.
 * "color_cache": entries per ray (default 32, 0 = off, at most hit_cap) of the replay's colour cache: a forward call that saves
 * hit lists also leaves the SH colour of each ray's first entries (12 B each) in a block owned by the handle, one per stream
 * (n_rays x entries x 12 B, kept until a larger call needs more); the irgs_trace_backward* call on the same stream whose `hits`
 * pointer is the one that forward wrote reads them instead of gathering a 192-byte SH row per hit.  Any other backward (another
 * stream, a forward in between, a copied list) gathers as before: same results either way.
 * Returns non-zero for unknown names. */
int irgs_set_option(irgs_tracer_t *h, const char *name, int64_t value);

/* The multiplier m of that stride order for a call of n_rays rays (the i-th ray started is (i * m) mod n_rays), 0 when the
 * caller's order is kept (fewer than 64 or more than 2^19 rays).  Pure host function, exposed for tests. */
int64_t irgs_stride_multiplier(int64_t n_rays);

/* Introspection (tests, diagnostics): "tree_depth" (levels of the PLOC tree of the last build; 0 = Karras tree), "ploc_iterations",
 * "n_slots" (streams seen so far), "n_surfels", "pack_epoch", "color_cache_bytes" (size of the colour-cache blocks), "grazing_pairs" / "grazing_pairs_compositing" (statistics build).
 * -1 for unknown names. */
int64_t irgs_get_info(irgs_tracer_t *h, const char *name);

/* Traversal statistics of the last irgs_trace_forward on this handle when statistics were enabled with
 * irgs_set_stats(h, 1): sums over rays of node visits, surfel tests, composited hits, traversal passes. */
int irgs_set_stats(irgs_tracer_t *h, int enable);
int irgs_get_stats(irgs_tracer_t *h, int64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* IRGS_B200_H */
