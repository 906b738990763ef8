// Internal declarations shared by the translation units of libirgs_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <mutex>
#include <string>

#include "../../include/irgs_b200.h"

namespace irgs {

// ---------------------------------------------------------------------------------------------------------------
// Acceleration structure: binary LBVH, one surfel per leaf, 64-byte nodes that hold BOTH children's bounds so one
// node visit is four 16-byte loads (replaces the closed-source OptiX GAS of src/bvh.cu:69-160).
//   a = (L.lo.x, L.lo.y, L.lo.z, L.hi.x)  b = (L.hi.y, L.hi.z, R.lo.x, R.lo.y)  c = (R.lo.z, R.hi.x, R.hi.y, R.hi.z)
//   d = (left, right, parent*2+side, unused); child >= 0: internal node index; child < 0: leaf, position ~child
//   in Morton order.  An empty child bound is stored as the degenerate far-away box EMPTY_FAR so that the slab
//   test can never pass for it.
// ---------------------------------------------------------------------------------------------------------------
struct __align__(16) Node {
    float4 a, b, c;
    int4 d;
};
static_assert(sizeof(Node) == 64, "node must be 64 bytes");

// Per-leaf surfel record in Morton order, packed from the caller's arrays at every trace call (64 bytes):
//   r0 = (mu.x, mu.y, mu.z, support radius^2)   r1 = (n.x, n.y, n.z, bits(surfel id))
//   r2 = (ru.x, ru.y, ru.z, rv.x)               r3 = (rv.y, rv.z, opacity, 0)
struct __align__(16) SurfelRec {
    float4 r0, r1, r2, r3;
};
static_assert(sizeof(SurfelRec) == 64, "record must be 64 bytes");

#define IRGS_EMPTY_FAR 1.0e30f

// Traversal copy of a node: 32 bytes, i.e. TWO 16-byte loads per visit instead of four (ncu: the walk is bound by the
// L1TEX tag stage -- one lookup per 16-byte request -- not by DRAM or L2 bandwidth).  Child bounds are quantised
// conservatively (lo rounded down, hi up) to 16 bits per coordinate on a grid spanning the padded root bound:
//   per child: w0 = lo.x | hi.x << 16,  w1 = lo.y | hi.y << 16,  w2 = lo.z | hi.z << 16,  w3 = child reference
//   (both planes of an axis in one word: the ray walk picks the NEAR and the FAR plane of the axis by the sign of the ray
//   direction with a per-ray byte-permute selector instead of evaluating both and taking min / max)
// child reference: >= 0 internal node, < 0 leaf ~pos, IRGS_CHILD_NONE = no child (empty bound).
// Coordinate q decodes to frame_lo + q * cell; the ray walk folds the decode into one byte-permute (assembling the
// float 2^23 + q) and one fma per plane.
struct __align__(16) QNode {
    uint4 l, r;
};
static_assert(sizeof(QNode) == 32, "quantised node must be 32 bytes");
#define IRGS_CHILD_NONE INT_MIN

// 4-wide traversal node of the forward kernel: up to four subtrees of the binary tree, each in the same 16-byte form as a QNode
// child (six 16-bit planes + reference).  One visit = two 256-bit loads from one 64-byte line and four slab tests; the chain of
// dependent node fetches of a ray is about half as long as on the binary tree.  Which subtrees a wide node holds is decided at
// build time (lbvh.cu): by default a GREEDY COLLAPSE -- start from the two children of a binary node and, while a slot is free,
// replace the internal child with the largest surface area by its own two children -- or, irgs_set_option("wide_fold", 1), the
// fixed fold of every other level (a node at even depth + its two children: a slot stays empty wherever a child is a leaf).
// A wide node is stored at the index of the binary node it starts from (the other entries are unused), so references stay valid.
struct __align__(32) QNode4 {
    uint4 c[4];
};
static_assert(sizeof(QNode4) == 64, "wide node must be 64 bytes");

}  // namespace irgs

struct irgs_tracer {
    int device = 0;
    int sm_count = 148;
    int64_t n = 0;         // surfels in the structure
    int64_t cap = 0;       // allocated capacity (surfels)
    irgs::Node *nodes = nullptr;        // [max(n-1,1)] float bounds (refit works on these)
    irgs::QNode *qnodes = nullptr;      // [max(n-1,1)] quantised binary nodes (re-trace backward, intersection test)
    irgs::QNode4 *qnodes4 = nullptr;    // [max(n-1,1)] 4-wide nodes the forward walk reads (valid where `even` is set)
    int *even = nullptr;                // [n] 1 where a wide node lives at the binary node's index (greedy collapse: its wide roots; fixed fold: even depth)
    int4 *wide_kids = nullptr;          // [n] greedy collapse: per wide root, the four slots as binary parent * 2 + side (-1: empty); frozen by the build
    int *wide_counts = nullptr;         // queue lengths of the collapse waves
    int wide_fold = 0;                  // irgs_set_option("wide_fold"): 0 greedy collapse (default), 1 fixed fold of every other level; next build
    int wide_fold_built = 0;            // what the current structure was built with
    float *boxes = nullptr;             // [n,6] unpadded per-surfel bounds, surfel order
    uint32_t *codes = nullptr, *codes_alt = nullptr;  // [n]
    int *order = nullptr, *order_alt = nullptr;       // [n] leaf position -> surfel id
    int *leaf_parent = nullptr;         // [n]   parent*2+side
    int *node_parent = nullptr;         // [n]   parent*2+side, -1 for the root
    int *flags = nullptr;               // [n]   bottom-up arrival counters
    int *ploc_cid = nullptr;            // [2][cap] PLOC cluster -> node reference (ping-pong)
    float *ploc_box = nullptr;          // [2][cap] PLOC cluster bounds, 8 floats each
    int *ploc_nn = nullptr, *ploc_counts = nullptr, *ploc_offs = nullptr, *ploc_totals = nullptr;
    int builder = 0;                    // 0: PLOC over the Morton order (default), 1: Karras LBVH
    int ploc_iterations = 0;            // clustering iterations of the last PLOC build
    int tree_depth = 0;                 // depth of the PLOC tree of the last build (0: Karras tree, depth <= 62 by construction)
    int *radix_hist = nullptr;          // [256 * n_tiles]
    int64_t radix_tiles_cap = 0;
    float *scene = nullptr;             // [24]: 0-5 centroid bounds as ordered ints, 6-11 root bound (floats),
                                        //       12-14 quantisation frame lo, 15-17 cell size, 18-23 bounds of the surfel boxes (ordered ints)
    irgs::SurfelRec *recs = nullptr;    // [n] leaf order
    int *inv_order = nullptr;           // [n] surfel id -> leaf position (written with the records)
    // Stream slots: every CUDA stream that launches on this handle owns one slot -- a persistent-kernel work counter, a
    // candidate scratch region and a sort scratch -- looked up from the stream of the call itself (slot_for), so that a
    // backward always uses the slot of the stream it runs on and calls on different streams never share one.
    static constexpr int MAX_SLOTS = 8;
    cudaStream_t slot_stream[MAX_SLOTS] = {};
    int n_slots = 0;
    std::mutex slot_mutex;                  // autograd runs backward on its own host thread
    unsigned long long *counter = nullptr;  // persistent-kernel work counters [MAX_SLOTS]
    uint4 *cand[MAX_SLOTS] = {};            // forward kernel candidate scratch of a slot: [threads][32] (t, id, alpha, leaf position)
    int64_t cand_threads[MAX_SLOTS] = {};   // threads the slot's scratch has room for
    // ray-coherence sort scratch, per stream slot
    uint32_t *rsort_keys[MAX_SLOTS][2] = {};
    int *rsort_vals[MAX_SLOTS][2] = {};
    int *rsort_hist[MAX_SLOTS] = {};
    int64_t rsort_cap[MAX_SLOTS] = {};
    float *ray_scratch[MAX_SLOTS] = {};     // generated rays of a forward call, materialised per stream slot (o[3n] d[3n])
    int64_t ray_scratch_cap[MAX_SLOTS] = {};
    // Colour cache of the backward replay, per stream slot: the forward kernel leaves the SH colour of a ray's first `hit_rgb_cc`
    // composited hits here (12 B each) when it saves hit lists; a backward on the same stream whose hit-list pointer is the one
    // the slot's last saving forward wrote reads them instead of gathering a 192-byte SH row per hit (launch_trace_backward).
    float *hit_rgb[MAX_SLOTS] = {};
    int64_t hit_rgb_floats[MAX_SLOTS] = {};     // allocated size
    const void *hit_rgb_key[MAX_SLOTS] = {};    // hit-list pointer of the forward whose colours the block holds (nullptr: none)
    int64_t hit_rgb_rays[MAX_SLOTS] = {};       // rays of that forward
    int hit_rgb_cc[MAX_SLOTS] = {};             // entries per ray of that forward
    int color_cache = 32;                       // entries per ray for the next forward calls (irgs_set_option("color_cache"); 0: off)
    long long pack_epoch = 0;               // bumped by every pack of the records and by every build / refit
    int contiguous_outputs = 0;             // the caller allocates a call's outputs as ONE block whenever they are back to back
    int skip_next_pack = 0;                 // the next irgs_trace_backward* reuses the records as they are (irgs_set_option)
    int gen_in_kernel = 0;                  // 1: generate incident / camera rays inside the forward kernel instead
    void *inc_pts[MAX_SLOTS] = {};          // per-point records of generated incident rays (IncPoint[inc_cap]), per stream slot
    int64_t inc_cap[MAX_SLOTS] = {};
    int bwd_carveout_pct = -1;              // backward replay kernel: carve-out hint in percent (-1: the driver's default)
    int fwd_blocks_per_sm = 0;              // forward kernel: resident blocks per SM of the persistent grid (0: what the occupancy allows)
    int carveout_pct = -1;                  // forward kernel: shared-memory carve-out hint in percent (-1: what the resident blocks need)
    int64_t stride_rays_max = 1 << 19;      // forward calls with at most this many rays start them in a stride order (0: never)
    int sort_rays_min = 0;                  // forward calls with at least this many rays are coherence-sorted (0: never;
                                            // measured on B200: the sort costs more than it saves, profiles/r01_notes.md)
    unsigned long long *stats = nullptr;    // [4]
    int stats_enabled = 0;
    int bwd_mode = 0;                       // 0: hit-parallel replay (default), 1: thread-per-ray replay
    bool built = false;
    // host-streaming resources
    cudaStream_t hs[2] = {nullptr, nullptr};
    cudaEvent_t hev[2] = {nullptr, nullptr};
    float *stage[2] = {nullptr, nullptr};
    int64_t stage_floats = 0;
    // ring of ray buffers filled by a dedicated copy stream, so that host->device copies run ahead of the two compute streams
    static constexpr int RING = 4;
    cudaStream_t hcopy = nullptr;
    float *ring[RING] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ring_full[RING] = {nullptr, nullptr, nullptr, nullptr}, ring_free[RING] = {nullptr, nullptr, nullptr, nullptr};
    int64_t ring_floats = 0;
    float *pt_buf = nullptr;                // incident host path: per-point inputs and gradients of a batch, 13 floats a point
    int64_t pt_floats = 0;
};

namespace irgs {

void set_error(const std::string &msg);
bool check(cudaError_t e, const char *what);
void count_launch(int n = 1);

#define IRGS_CHECK(call)                                   \
    do {                                                   \
        if (!irgs::check((call), #call)) return 1;         \
    } while (0)

// lbvh.cu
int lbvh_build(irgs_tracer *h, bool refit_only, cudaStream_t s);  // boxes[] already filled
int lbvh_reserve(irgs_tracer *h, int64_t n);
int launch_ray_order(irgs_tracer *h, int slot, const float *rays_o, const float *rays_d, int64_t n_rays, int **order_out,
                     cudaStream_t s);
int slot_for(irgs_tracer *h, cudaStream_t s);   // capi.cu; -1 on error
int launch_bounds_from_proxy(irgs_tracer *h, const float *verts, int vps, cudaStream_t s);
int launch_bounds_from_surfels(irgs_tracer *h, const float *means, const float *opacity, const float *ru,
                               const float *rv, const float *normals, float alpha_min, cudaStream_t s);

// trace.cu
struct TraceArgs {
    int64_t n_rays;
    int n_surf, S, K, deg, back_culling;
    float alpha_min, T_min;
    const float *rays_o, *rays_d;
    const float *means, *opacity, *ru, *rv, *normals, *features, *shs;
    // forward outputs / backward saved outputs
    float *color, *normal, *feature, *depth, *alpha;
    int32_t *hit_count, *hits;
    int hit_cap;
    float *hit_rgb;   // colour cache of the stream slot ([n_rays, rgb_cap, 3]) or nullptr; set by the launchers, not by the ABI
    int rgb_cap;
    // fused incident-ray generation: when gen_pos != nullptr the rays are not read from rays_o / rays_d
    // but generated from (shading point, normal, azimuth) and the sample index: ray = point * gen_S + sample
    const float *gen_pos, *gen_nrm, *gen_azim;   // [P,3], [P,3], [P] or nullptr (no random rotation)
    int gen_S;
    float gen_tmin;
    int64_t gen_P;
    const struct IncPoint *gen_pts;   // [P] per-point records (rotation, azimuth sin / cos), filled by launch_incident_prepare
    const struct IncTab *gen_tab;     // [gen_S] per-sample table (incident_table)
    // fused PRIMARY-ray generation: when cam_W > 0 ray = v * cam_W + u is the pinhole ray of pixel (u, v),
    // scene/cameras.py:87-100: d = normalize(M ((u - W/2 + 0.5) / fx, (v - H/2 + 0.5) / fy, 1)), origin = the camera centre
    int cam_W, cam_H;
    float cam_fx, cam_fy, cam_o[3], cam_M[9];
    const int *ray_order;  // forward: optional processing order (coherence sort); results are still written per ray id
    int64_t ray_mul;       // forward: 0, or an odd multiplier coprime to n_rays (n_rays <= 2^19): the i-th ray started is (i * ray_mul) % n_rays
                           // (small launches: spreads the heavy rays of one pixel bundle over the warps, launch_trace_forward)
    // backward
    const float *gC, *gN, *gF, *gD, *gO;
    int64_t gout_period;  // 0: gout arrays have n_rays rows; >0: ray r reads row (gout_offset + r) % period
    int64_t gout_offset;
    float *g_rays_o, *g_rays_d, *grad_fused, *grad_features;
};
int launch_pack_records(irgs_tracer *h, const TraceArgs &a, cudaStream_t s);
int launch_trace_forward(irgs_tracer *h, const TraceArgs &a, cudaStream_t s);
int64_t stride_multiplier(int64_t n_rays);
int launch_trace_backward(irgs_tracer *h, const TraceArgs &a, cudaStream_t s);
int launch_intersection_test(irgs_tracer *h, const TraceArgs &a, uint8_t *out, cudaStream_t s);
// incident-ray generation: the per-sample table of a sample count (cached per device, filled once) and the per-point records
// of a call (scratch owned by the handle, one block per stream slot); launch_incident_prepare fills a.gen_pts / a.gen_tab
const struct IncTab *incident_table(int sample_num, cudaStream_t s);
int launch_incident_prepare(irgs_tracer *h, TraceArgs &a, cudaStream_t s);
int launch_incident_backward(const float *position, const float *normals, const float *azimuth, int64_t n_points,
                             int sample_num, float t_min, const float *g_rays_o, const float *g_rays_d, float *grad_position,
                             float *grad_normal_pt, cudaStream_t s);
int launch_incident_rays(const float *position, const float *normals, const float *azimuth, int64_t n_points, int sample_num,
                         float t_min, float *rays_o, float *rays_d, cudaStream_t s);
int launch_generated_rays(const TraceArgs &a, float *rays_o, float *rays_d, cudaStream_t s);   // the rays load_ray() produces, written out
int launch_unpack_grads(const float *fused, int64_t n, int K, float *gm, float *go, float *gru, float *grv,
                        float *gn, float *gsh, cudaStream_t s);

}  // namespace irgs
