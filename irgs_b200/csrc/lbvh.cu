// LBVH build / refit over per-surfel bounds (sm_100a).
//
// Replaces optix::Gas (BUILD and UPDATE) of /root/reference/submodules/surfel_tracer/src/bvh.cu:69-160, which hands a
// 20-triangles-per-surfel soup to the closed-source optixAccelBuild.  Here the primitive is the surfel itself:
//   1. per-surfel AABB (from the caller's proxy vertices, or analytically from the surfel parameters) + scene bounds
//   2. 30-bit Morton code of the AABB centroid
//   3. hand-written LSD radix sort (4 passes x 8 bits, stable; histogram -> scan -> ranked scatter)
//   4. Karras-2012 hierarchy, one thread per internal node
//   5. bottom-up bound propagation with arrival counters; refit (update_bvh) re-runs only steps 1 and 5.
// All work is HBM-bound integer/float streaming over N surfels; no tensor-core shaped work exists here.
#include <cfloat>
#include <climits>

#include "internal.cuh"

namespace irgs {

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void scene_init_kernel(int *scene_i) {
    int t = threadIdx.x;
    if (t < 3) scene_i[t] = INT_MAX;
    else if (t < 6) scene_i[t] = INT_MIN;
    else if (t < 18) scene_i[t] = 0;
    else if (t < 21) scene_i[t] = INT_MAX;   // 18-23: bounds of the surfel boxes themselves (ordered ints)
    else if (t < 24) scene_i[t] = INT_MIN;
}

__device__ __forceinline__ void scene_reduce(int *scene_i, bool valid, const float lo[3], const float hi[3]) {
    // centroid bounds, warp-reduced before the six atomics
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float c = 0.5f * (lo[k] + hi[k]);
        int vmin = valid ? f2ord(c) : INT_MAX, vmax = valid ? f2ord(c) : INT_MIN;
        vmin = __reduce_min_sync(0xffffffffu, vmin);
        vmax = __reduce_max_sync(0xffffffffu, vmax);
        int bmin = valid ? f2ord(lo[k]) : INT_MAX, bmax = valid ? f2ord(hi[k]) : INT_MIN;
        bmin = __reduce_min_sync(0xffffffffu, bmin);
        bmax = __reduce_max_sync(0xffffffffu, bmax);
        if ((threadIdx.x & 31) == 0) {
            if (vmin != INT_MAX) atomicMin(&scene_i[k], vmin);
            if (vmax != INT_MIN) atomicMax(&scene_i[3 + k], vmax);
            if (bmin != INT_MAX) atomicMin(&scene_i[18 + k], bmin);
            if (bmax != INT_MIN) atomicMax(&scene_i[21 + k], bmax);
        }
    }
}

// Step 1a: bounds from the reference caller's proxy vertices (scene/gaussian_model.py:712-723): AABB of the
// `vps` vertices of surfel g.  fminf/fmaxf drop NaNs, so an all-NaN proxy (opacity < alpha_min) stays empty.
__global__ void bounds_from_proxy_kernel(const float *__restrict__ verts, int vps, int n, float *__restrict__ boxes,
                                         int *scene_i) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool valid = false;
    if (g < n) {
        const float *v = verts + (size_t)g * vps * 3;
        for (int i = 0; i < vps; ++i) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float x = v[3 * i + k];
                lo[k] = fminf(lo[k], x);
                hi[k] = fmaxf(hi[k], x);
            }
        }
        valid = lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2] && hi[0] < FLT_MAX && lo[0] > -FLT_MAX &&
                hi[1] < FLT_MAX && lo[1] > -FLT_MAX && hi[2] < FLT_MAX && lo[2] > -FLT_MAX;
        if (!valid) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { boxes[6 * (size_t)g + k] = lo[k]; boxes[6 * (size_t)g + 3 + k] = hi[k]; }
    }
    scene_reduce(scene_i, valid, lo, hi);
}

// Step 1b: analytic bound of { x : (ru.(x-mu))^2 + (rv.(x-mu))^2 <= 2 ln(opacity/alpha_min), n.(x-mu) = 0 }, the
// support on which alpha >= alpha_min (the set the reference's proxy icosahedron is scaled to cover).
__global__ void bounds_from_surfels_kernel(const float *__restrict__ means, const float *__restrict__ opacity,
                                           const float *__restrict__ ru, const float *__restrict__ rv,
                                           const float *__restrict__ normals, float alpha_min, int n,
                                           float *__restrict__ boxes, int *scene_i) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool valid = false;
    if (g < n) {
        float op = opacity[g];
        float nx = normals[3 * (size_t)g], ny = normals[3 * (size_t)g + 1], nz = normals[3 * (size_t)g + 2];
        float nn = sqrtf(nx * nx + ny * ny + nz * nz);
        if (op > alpha_min && nn > 0.0f) {
            float r = sqrtf(2.0f * logf(op / alpha_min));
            nx /= nn; ny /= nn; nz /= nn;
            float e1[3], e2[3];
            if (fabsf(nx) < 0.6f) { e1[0] = 0.f; e1[1] = -nz; e1[2] = ny; } else { e1[0] = -nz; e1[1] = 0.f; e1[2] = nx; }
            float l1 = rsqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
            e1[0] *= l1; e1[1] *= l1; e1[2] *= l1;
            e2[0] = ny * e1[2] - nz * e1[1]; e2[1] = nz * e1[0] - nx * e1[2]; e2[2] = nx * e1[1] - ny * e1[0];
            const float *a = ru + 3 * (size_t)g, *b = rv + 3 * (size_t)g;
            float m00 = a[0] * e1[0] + a[1] * e1[1] + a[2] * e1[2], m01 = a[0] * e2[0] + a[1] * e2[1] + a[2] * e2[2];
            float m10 = b[0] * e1[0] + b[1] * e1[1] + b[2] * e1[2], m11 = b[0] * e2[0] + b[1] * e2[1] + b[2] * e2[2];
            float det = m00 * m11 - m01 * m10;
            if (fabsf(det) > 1e-30f) {
                float i00 = m11 / det, i01 = -m01 / det, i10 = -m10 / det, i11 = m00 / det;
                valid = true;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float c0 = e1[k] * i00 + e2[k] * i10, c1 = e1[k] * i01 + e2[k] * i11;
                    float hk = r * sqrtf(c0 * c0 + c1 * c1);
                    float mu = means[3 * (size_t)g + k];
                    lo[k] = mu - hk; hi[k] = mu + hk;
                    valid = valid && (hk < FLT_MAX) && (fabsf(mu) < FLT_MAX);
                }
            }
        }
        if (!valid) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { boxes[6 * (size_t)g + k] = lo[k]; boxes[6 * (size_t)g + 3 + k] = hi[k]; }
    }
    scene_reduce(scene_i, valid, lo, hi);
}

// Step 2: 30-bit Morton codes of the AABB centroids; invalid (empty) surfels sort to the end.
__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void morton_kernel(const float *__restrict__ boxes, const int *__restrict__ scene_i, int n,
                              uint32_t *__restrict__ codes, int *__restrict__ order) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    uint32_t code = 0x3fffffffu;
    const float *bx = boxes + 6 * (size_t)g;
    if (bx[0] <= bx[3]) {
        uint32_t q[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float clo = ord2f(scene_i[k]), chi = ord2f(scene_i[3 + k]);
            float c = 0.5f * (bx[k] + bx[3 + k]);
            float ext = chi - clo;
            float u = ext > 0.f ? (c - clo) / ext : 0.f;
            int v = (int)(u * 1024.0f);
            q[k] = (uint32_t)min(max(v, 0), 1023);
        }
        code = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    }
    codes[g] = code;
    order[g] = g;
}

// Step 3: LSD radix sort, 8 bits per pass, stable.  Tile = 256 threads x 16 keys; each warp owns a contiguous
// 512-key segment of the tile so that ranking by (warp, round, lane) preserves input order.
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint32_t *__restrict__ keys, int n, int shift,
                                                                int n_tiles, int *__restrict__ hist) {
    __shared__ int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        int idx = base + i;
        if (idx < n) atomicAdd(&sh[(keys[idx] >> shift) & 0xff], 1);
    }
    __syncthreads();
    hist[threadIdx.x * n_tiles + blockIdx.x] = sh[threadIdx.x];
}

// exclusive scan over hist[256 * n_tiles] (digit-major), single block
__global__ void __launch_bounds__(1024) radix_scan_kernel(int *hist, int total) {
    __shared__ int part[1024];
    int per = (total + 1023) / 1024;
    int b = threadIdx.x * per, e = min(b + per, total);
    int s = 0;
    for (int i = b; i < e; ++i) s += hist[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
        int v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = b; i < e; ++i) { int v = hist[i]; hist[i] = run; run += v; }
}

__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                   const int *__restrict__ vals_in, int n, int shift,
                                                                   int n_tiles, const int *__restrict__ hist,
                                                                   uint32_t *__restrict__ keys_out,
                                                                   int *__restrict__ vals_out) {
    __shared__ int warp_hist[RS_THREADS / 32][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const int seg = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS];
    int rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int idx = seg + r * 32 + lane;
        bool ok = idx < n;
        key[r] = ok ? keys_in[idx] : 0u;
        uint32_t digit = ok ? ((key[r] >> shift) & 0xff) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, digit);
        int before = __popc(peers & ((1u << lane) - 1u));
        int basecnt = ok ? warp_hist[warp][digit] : 0;
        __syncwarp();
        if (ok && before == 0) warp_hist[warp][digit] = basecnt + __popc(peers);
        __syncwarp();
        rank[r] = basecnt + before;
    }
    __syncthreads();
    {   // exclusive scan over warps per digit, plus the tile's global base
        int d = threadIdx.x;  // 256 threads == 256 digits
        int run = hist[d * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) { int v = warp_hist[w][d]; warp_hist[w][d] = run; run += v; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int idx = seg + r * 32 + lane;
        if (idx < n) {
            int dst = warp_hist[warp][(key[r] >> shift) & 0xff] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = vals_in[idx];
        }
    }
}

// Stable LSD sort of (key, value) pairs on the low 8*n_passes key bits; the sorted data ends up in keys/vals
// (pointers are swapped as needed).  hist must hold 256 * ceil(n / RS_TILE) ints.
int radix_sort_pairs(uint32_t *&keys, uint32_t *&keys_alt, int *&vals, int *&vals_alt, int n, int n_passes, int *hist,
                     cudaStream_t s) {
    const int n_tiles = (n + RS_TILE - 1) / RS_TILE;
    for (int pass = 0; pass < n_passes; ++pass) {
        const int shift = 8 * pass;
        radix_hist_kernel<<<n_tiles, RS_THREADS, 0, s>>>(keys, n, shift, n_tiles, hist);
        radix_scan_kernel<<<1, 1024, 0, s>>>(hist, 256 * n_tiles);
        radix_scatter_kernel<<<n_tiles, RS_THREADS, 0, s>>>(keys, vals, n, shift, n_tiles, hist, keys_alt, vals_alt);
        count_launch(3);
        uint32_t *tk = keys; keys = keys_alt; keys_alt = tk;
        int *tv = vals; vals = vals_alt; vals_alt = tv;
    }
    IRGS_CHECK(cudaGetLastError());
    return 0;
}
int radix_hist_ints(int n) { return 256 * ((n + RS_TILE - 1) / RS_TILE); }

// Coherence key of a ray: 15-bit Morton code of the origin cell (32^3 grid over the root bound) followed by an
// 8-bit octahedral direction bin.  Rays that start close together and point the same way become neighbours in a
// warp of the forward kernel, so they walk the same nodes and reach their leaves together.
__global__ void ray_key_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d, int n,
                               const float *__restrict__ root, uint32_t *__restrict__ keys, int *__restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float lo = root[k], hi = root[3 + k];
        const float ext = hi - lo;
        float u = ext > 0.f ? (rays_o[3 * (size_t)i + k] - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);
        q[k] = (uint32_t)min((int)(u * 32.0f), 31);
    }
    const uint32_t cell = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    float dx = rays_d[3 * (size_t)i], dy = rays_d[3 * (size_t)i + 1], dz = rays_d[3 * (size_t)i + 2];
    const float inv = 1.0f / fmaxf(fabsf(dx) + fabsf(dy) + fabsf(dz), 1e-30f);
    float px = dx * inv, py = dy * inv;
    if (dz < 0.f) {
        const float ax = (1.0f - fabsf(py)) * (px >= 0.f ? 1.f : -1.f);
        const float ay = (1.0f - fabsf(px)) * (py >= 0.f ? 1.f : -1.f);
        px = ax; py = ay;
    }
    const uint32_t du = (uint32_t)min(max((int)((px * 0.5f + 0.5f) * 16.0f), 0), 15);
    const uint32_t dv = (uint32_t)min(max((int)((py * 0.5f + 0.5f) * 16.0f), 0), 15);
    keys[i] = (cell << 8) | (du << 4) | dv;
    vals[i] = i;
}

int launch_ray_order(irgs_tracer *h, int slot, const float *rays_o, const float *rays_d, int64_t n_rays, int **order_out,
                     cudaStream_t s) {
    const int n = (int)n_rays;
    if (n_rays > h->rsort_cap[slot]) {
        IRGS_CHECK(cudaDeviceSynchronize());
        for (int k = 0; k < 2; ++k) {
            if (h->rsort_keys[slot][k]) cudaFree(h->rsort_keys[slot][k]);
            if (h->rsort_vals[slot][k]) cudaFree(h->rsort_vals[slot][k]);
            h->rsort_keys[slot][k] = nullptr; h->rsort_vals[slot][k] = nullptr;
            IRGS_CHECK(cudaMalloc(&h->rsort_keys[slot][k], sizeof(uint32_t) * (size_t)n_rays));
            IRGS_CHECK(cudaMalloc(&h->rsort_vals[slot][k], sizeof(int) * (size_t)n_rays));
        }
        if (h->rsort_hist[slot]) cudaFree(h->rsort_hist[slot]);
        h->rsort_hist[slot] = nullptr;
        IRGS_CHECK(cudaMalloc(&h->rsort_hist[slot], sizeof(int) * (size_t)radix_hist_ints(n)));
        h->rsort_cap[slot] = n_rays;
    }
    uint32_t *k0 = h->rsort_keys[slot][0], *k1 = h->rsort_keys[slot][1];
    int *v0 = h->rsort_vals[slot][0], *v1 = h->rsort_vals[slot][1];
    ray_key_kernel<<<(n + 255) / 256, 256, 0, s>>>(rays_o, rays_d, n, h->scene + 6, k0, v0);
    count_launch();
    if (radix_sort_pairs(k0, k1, v0, v1, n, 3, h->rsort_hist[slot], s)) return 1;
    *order_out = v0;
    return 0;
}

// Step 4: Karras 2012.  Keys are made unique by appending the sorted position.
__device__ __forceinline__ int lcp(const uint32_t *__restrict__ codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = codes[i], b = codes[j];
    return a == b ? 32 + __clz(i ^ j) : __clz(a ^ b);
}
__global__ void hierarchy_kernel(const uint32_t *__restrict__ codes, int n, Node *__restrict__ nodes,
                                 int *__restrict__ leaf_parent, int *__restrict__ node_parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 1) {
        if (i == 0) {
            nodes[0].d = make_int4(~0, ~0, -1, 1 /* right slot unused */);
            leaf_parent[0] = 0;
            node_parent[0] = -1;
        }
        return;
    }
    if (i >= n - 1) return;
    int d = (lcp(codes, n, i, i + 1) - lcp(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lcp(codes, n, i, i - d);
    int lmax = 2;
    while (lcp(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lcp(codes, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lcp(codes, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lcp(codes, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left = (lo == gamma) ? ~gamma : gamma;
    int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    if (left < 0) leaf_parent[gamma] = 2 * i; else node_parent[gamma] = 2 * i;
    if (right < 0) leaf_parent[gamma + 1] = 2 * i + 1; else node_parent[gamma + 1] = 2 * i + 1;
    if (i == 0) node_parent[0] = -1;
    nodes[i].d = make_int4(left, right, 0, 0);
}

// Step 4b: PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) instead of the Karras hierarchy.
// The Morton-sorted leaves are the initial clusters; every iteration each cluster looks PLOC_R places to either side
// for the neighbour whose union with it has the smallest surface area, mutual nearest neighbours are merged into a
// new node, and the cluster array is compacted in order.  On the C3 rays the resulting tree needs 40 % fewer node
// visits than the Karras LBVH over the same Morton order (scripts/exp/exp_bvh.c: 87 -> 52.5 per ray), for ~50
// iterations of four small kernels at build time; refits (update_bvh) are unaffected.
// Nodes are numbered from the top down as they are created (the last merge, the root, gets index 0), children of
// neighbouring clusters next to each other.
#ifndef IRGS_PLOC_R
#define IRGS_PLOC_R 8
#endif
constexpr int PLOC_R = IRGS_PLOC_R;
constexpr int PLOC_TB = 256;     // nearest-neighbour search block
constexpr int WIDE_WAVES = 96;   // waves of the greedy collapse into 4-wide nodes (>= the depth of any tree the walk's stack accepts)
constexpr int PLOC_SB = 1024;    // compaction block

struct __align__(16) PlocBox { float4 lo, hi; };

__global__ void ploc_init_kernel(const float *__restrict__ boxes, const int *__restrict__ order, int n,
                                 int *__restrict__ cid, PlocBox *__restrict__ cbox) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    const float *b = boxes + 6 * (size_t)order[pos];
    PlocBox x;
    if (b[0] <= b[3]) { x.lo = make_float4(b[0], b[1], b[2], 0.f); x.hi = make_float4(b[3], b[4], b[5], 0.f); }
    else { x.lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f); x.hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f); }
    cid[pos] = ~pos;
    cbox[pos] = x;
}

// Merge cost of two clusters: the surface area of their union.  EMPTY clusters (surfels with opacity < alpha_min: inverted
// bounds, all sorted to the end of the Morton order) cost nothing among themselves and are infinitely expensive for a
// non-empty cluster, so that they pair up with each other -- a balanced subtree of depth log2(M) -- instead of being
// absorbed one per iteration by the clusters at the end of the occupied range (M iterations and a chain of depth M).
__device__ __forceinline__ float union_area(const float4 alo, const float4 ahi, const float4 blo, const float4 bhi) {
    const bool ea = !(alo.x <= ahi.x), eb = !(blo.x <= bhi.x);
    if (ea || eb) return (ea && eb) ? 0.f : INFINITY;
    const float ex = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x);
    const float ey = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y);
    const float ez = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return __fmaf_rn(ex, ey, __fmaf_rn(ey, ez, __fmul_rn(ez, ex)));   // symmetric in (a, b): mutual choices are consistent
}

__global__ void __launch_bounds__(PLOC_TB) ploc_nn_kernel(const PlocBox *__restrict__ cbox, int m, int *__restrict__ nn) {
    __shared__ float4 s_lo[PLOC_TB + 2 * PLOC_R], s_hi[PLOC_TB + 2 * PLOC_R];
    const int base = blockIdx.x * PLOC_TB - PLOC_R;
    for (int t = threadIdx.x; t < PLOC_TB + 2 * PLOC_R; t += PLOC_TB) {
        const int j = base + t;
        if (j >= 0 && j < m) { s_lo[t] = cbox[j].lo; s_hi[t] = cbox[j].hi; }
    }
    __syncthreads();
    const int i = blockIdx.x * PLOC_TB + threadIdx.x;
    if (i >= m) return;
    const float4 lo = s_lo[threadIdx.x + PLOC_R], hi = s_hi[threadIdx.x + PLOC_R];
    // Ties go to the pair partner i ^ 1 first (then to the lowest index): where many costs are equal -- runs of empty or of
    // identical clusters -- everybody picks its partner and all of them merge in one iteration.
    float best = INFINITY;
    int bj = -1;
    {
        const int j = i ^ 1;
        if (j < m) {
            const int dlt = j - i;
            best = union_area(lo, hi, s_lo[threadIdx.x + PLOC_R + dlt], s_hi[threadIdx.x + PLOC_R + dlt]);
            bj = j;
        }
    }
#pragma unroll
    for (int dlt = -PLOC_R; dlt <= PLOC_R; ++dlt) {
        const int j = i + dlt;
        if (dlt == 0 || j < 0 || j >= m) continue;
        const float ar = union_area(lo, hi, s_lo[threadIdx.x + PLOC_R + dlt], s_hi[threadIdx.x + PLOC_R + dlt]);
        if (ar < best || bj < 0) { best = ar; bj = j; }
    }
    nn[i] = bj;
}

// flags of cluster i: bit 0 = survives the compaction (not the upper half of a merged pair), bit 16 = creates a node
__device__ __forceinline__ int ploc_flags(const int *__restrict__ nn, int m, int i) {
    if (i >= m) return 0;
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    return (mutual && i > j ? 0 : 1) | (mutual && i < j ? (1 << 16) : 0);
}

__device__ __forceinline__ int block_excl_scan(int v, int *warp_sums, int &block_total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - w;
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    block_total = warp_sums[32];
    return warp_sums[wid] + inc - v;
}

__global__ void __launch_bounds__(PLOC_SB) ploc_count_kernel(const int *__restrict__ nn, int m, int *__restrict__ counts) {
    __shared__ int ws[33];
    int total;
    block_excl_scan(ploc_flags(nn, m, blockIdx.x * PLOC_SB + threadIdx.x), ws, total);
    if (threadIdx.x == 0) counts[blockIdx.x] = total;   // survivors | merges << 16 (both < 2^16 per block)
}

// exclusive scan of the per-block (survivors, merges) counts, single block; totals[0] = clusters after the iteration,
// totals[1] = nodes created
__global__ void __launch_bounds__(1024) ploc_scan_kernel(const int *__restrict__ counts, int n_blocks, int2 *__restrict__ offs,
                                                         int *__restrict__ totals) {
    __shared__ int ws_a[33], ws_b[33];
    int run_a = 0, run_b = 0;
    for (int b0 = 0; b0 < n_blocks; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const int c = b < n_blocks ? counts[b] : 0;
        int ta, tb;
        const int ea = block_excl_scan(c & 0xffff, ws_a, ta);
        const int eb = block_excl_scan(c >> 16, ws_b, tb);
        if (b < n_blocks) offs[b] = make_int2(run_a + ea, run_b + eb);
        run_a += ta; run_b += tb;
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = run_a; totals[1] = run_b; }
}

__global__ void __launch_bounds__(PLOC_SB) ploc_scatter_kernel(const int *__restrict__ nn, int m, const int *__restrict__ cid_in,
                                                               const PlocBox *__restrict__ cbox_in,
                                                               const int2 *__restrict__ offs, const int *__restrict__ totals,
                                                               int free_hi, int *__restrict__ cid_out,
                                                               PlocBox *__restrict__ cbox_out, Node *__restrict__ nodes,
                                                               int *__restrict__ leaf_parent, int *__restrict__ node_parent) {
    __shared__ int ws[33];
    const int i = blockIdx.x * PLOC_SB + threadIdx.x;
    const int f = ploc_flags(nn, m, i);
    int total;
    const int ex = block_excl_scan(f, ws, total);
    if (i >= m || !(f & 1)) return;
    const int2 off = offs[blockIdx.x];
    const int dst = off.x + (ex & 0xffff);
    if (f >> 16) {
        const int j = nn[i];
        const int id = free_hi - totals[1] + off.y + (ex >> 16);   // new node
        const int cl = cid_in[i], cr = cid_in[j];
        const PlocBox a = cbox_in[i], b = cbox_in[j];
        PlocBox u;
        u.lo = make_float4(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z), 0.f);
        u.hi = make_float4(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z), 0.f);
        nodes[id].d = make_int4(cl, cr, 0, 0);
        if (cl < 0) leaf_parent[~cl] = 2 * id; else node_parent[cl] = 2 * id;
        if (cr < 0) leaf_parent[~cr] = 2 * id + 1; else node_parent[cr] = 2 * id + 1;
        if (id == 0) node_parent[0] = -1;
        cid_out[dst] = id;
        cbox_out[dst] = u;
    } else {
        cid_out[dst] = cid_in[i];
        cbox_out[dst] = cbox_in[i];
    }
}

// Depth of the deepest leaf (the ray walk keeps one stack entry per level at most: STACK = 64 entries).
__global__ void tree_depth_kernel(const int *__restrict__ leaf_parent, const int *__restrict__ node_parent, int n,
                                  int *__restrict__ max_depth) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    int depth = 0;
    if (leaf < n) {
        int p = leaf_parent[leaf];
        depth = 1;
        while ((p = node_parent[p >> 1]) >= 0) ++depth;
    }
    depth = __reduce_max_sync(0xffffffffu, depth);
    if ((threadIdx.x & 31) == 0) atomicMax(max_depth, depth);
}

int ploc_build(irgs_tracer *h, cudaStream_t s) {
    const int n = (int)h->n;
    int *cid[2] = {h->ploc_cid, h->ploc_cid + h->cap};
    PlocBox *cbox[2] = {reinterpret_cast<PlocBox *>(h->ploc_box), reinterpret_cast<PlocBox *>(h->ploc_box) + h->cap};
    ploc_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(h->boxes, h->order, n, cid[0], cbox[0]);
    count_launch();
    int m = n, free_hi = n - 1, cur = 0;
    h->ploc_iterations = 0;
    while (m > 1) {
        ++h->ploc_iterations;
        const int nb = (m + PLOC_SB - 1) / PLOC_SB;
        ploc_nn_kernel<<<(m + PLOC_TB - 1) / PLOC_TB, PLOC_TB, 0, s>>>(cbox[cur], m, h->ploc_nn);
        ploc_count_kernel<<<nb, PLOC_SB, 0, s>>>(h->ploc_nn, m, h->ploc_counts);
        ploc_scan_kernel<<<1, 1024, 0, s>>>(h->ploc_counts, nb, reinterpret_cast<int2 *>(h->ploc_offs), h->ploc_totals);
        ploc_scatter_kernel<<<nb, PLOC_SB, 0, s>>>(h->ploc_nn, m, cid[cur], cbox[cur], reinterpret_cast<const int2 *>(h->ploc_offs),
                                                   h->ploc_totals, free_hi, cid[cur ^ 1], cbox[cur ^ 1], h->nodes,
                                                   h->leaf_parent, h->node_parent);
        count_launch(4);
        int tot[2] = {0, 0};
        IRGS_CHECK(cudaMemcpyAsync(tot, h->ploc_totals, sizeof tot, cudaMemcpyDeviceToHost, s));
        IRGS_CHECK(cudaStreamSynchronize(s));
        if (tot[1] < 1 || tot[0] != m - tot[1]) { set_error("PLOC iteration made no progress"); return 1; }
        m = tot[0];
        free_hi -= tot[1];
        cur ^= 1;
    }
    if (free_hi != 0) { set_error("PLOC node count mismatch"); return 1; }
    // clustering gives no depth guarantee (the Karras tree over 30 + 32 key bits does): measure it
    IRGS_CHECK(cudaMemsetAsync(h->ploc_totals, 0, sizeof(int), s));
    tree_depth_kernel<<<(n + 255) / 256, 256, 0, s>>>(h->leaf_parent, h->node_parent, n, h->ploc_totals);
    count_launch();
    int depth = 0;
    IRGS_CHECK(cudaMemcpyAsync(&depth, h->ploc_totals, sizeof(int), cudaMemcpyDeviceToHost, s));
    IRGS_CHECK(cudaStreamSynchronize(s));
    h->tree_depth = depth;
    return 0;
}

// Step 5: bottom-up bounds.  Each leaf thread writes its (padded) bound into its parent's slot and climbs; the
// second arrival at a node reads the sibling slot, forms the union and continues.
__device__ __forceinline__ void store_slot(Node *nd, int side, const float lo[3], const float hi[3]) {
    float l[3], h[3];
    bool empty = !(lo[0] <= hi[0]);
#pragma unroll
    for (int k = 0; k < 3; ++k) { l[k] = empty ? IRGS_EMPTY_FAR : lo[k]; h[k] = empty ? IRGS_EMPTY_FAR : hi[k]; }
    float *f = reinterpret_cast<float *>(nd);
    if (side == 0) { f[0] = l[0]; f[1] = l[1]; f[2] = l[2]; f[3] = h[0]; f[4] = h[1]; f[5] = h[2]; }
    else { f[6] = l[0]; f[7] = l[1]; f[8] = l[2]; f[9] = h[0]; f[10] = h[1]; f[11] = h[2]; }
}
__device__ __forceinline__ void load_slot(const Node *nd, int side, float lo[3], float hi[3]) {
    const float *f = reinterpret_cast<const float *>(nd) + (side ? 6 : 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = __ldcg(f + k); hi[k] = __ldcg(f + 3 + k); }
    if (lo[0] >= IRGS_EMPTY_FAR) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
    }
}

__global__ void refit_kernel(const float *__restrict__ boxes, const int *__restrict__ order,
                             const int *__restrict__ leaf_parent, const int *__restrict__ node_parent, int n,
                             const int *__restrict__ scene_i, Node *nodes, int *flags, float *root_bound) {
    int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    // absolute pad: a few float ulps of the scene scale, so that neither the rounding of the fma slab test nor that
    // of the plane-hit arithmetic (trace_common.cuh leaf_test) can make the walk reject a surfel the hit test accepts;
    // relative pad covers the proxy's 0.999993 in-radius.  The scale is ISOTROPIC -- the largest extent / coordinate of
    // the surfel boxes over all three axes: a flat scene (all surfels in one axis-aligned plane) is looked at from
    // distances given by its other two axes, and the rounding of a ray's plane crossings scales with those.
    float scale = 1e-3f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float blo = ord2f(scene_i[18 + k]), bhi = ord2f(scene_i[21 + k]);
        if (blo <= bhi) scale = fmaxf(scale, fmaxf(bhi - blo, fmaxf(fabsf(blo), fabsf(bhi))));
    }
    const float pad_abs = 8e-6f * scale;
    const float *bx = boxes + 6 * (size_t)order[leaf];
    float lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float a = bx[k], b = bx[3 + k];
        float pad = 1e-4f * (b - a) + pad_abs;
        lo[k] = a - pad; hi[k] = b + pad;
    }
    if (!(bx[0] <= bx[3])) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = INFINITY; hi[k] = -INFINITY; }
    }
    int p = leaf_parent[leaf];
    if (n == 1) {
        store_slot(&nodes[0], 0, lo, hi);
        float e_lo[3] = {INFINITY, INFINITY, INFINITY}, e_hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        store_slot(&nodes[0], 1, e_lo, e_hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) { root_bound[k] = lo[k]; root_bound[3 + k] = hi[k]; }
        return;
    }
    while (true) {
        int parent = p >> 1, side = p & 1;
        store_slot(&nodes[parent], side, lo, hi);
        __threadfence();
        int old = atomicAdd(&flags[parent], 1);
        if (old == 0) return;  // first arrival: the sibling subtree is still in flight
        float slo[3], shi[3];
        load_slot(&nodes[parent], side ^ 1, slo, shi);
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], slo[k]); hi[k] = fmaxf(hi[k], shi[k]); }
        p = node_parent[parent];
        if (p < 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { root_bound[k] = lo[k]; root_bound[3 + k] = hi[k]; }
            return;
        }
    }
}

// Step 6: quantised traversal copy of the nodes (see QNode in internal.cuh).
// Frame: code q decodes to frame_lo + q * cell with cell = root extent / 65520 and frame_lo = root lo - 4 cells, so
// every bound quantises into [4, 65524]; lo is rounded down and hi up, then each is moved outward by two more cells:
// the ray walk evaluates the plane as fma(2^23 + q, cell/d, (frame_lo - 2^23 cell - o)/d), whose two rounded
// constants can be off by up to ~one cell in space (see trace_common.cuh).
__global__ void quant_frame_kernel(float *scene) {
    const int k = threadIdx.x;
    if (k < 3) {
        // one cell size for all three axes (the largest root extent / 65520): the two-cell outward shift of every plane
        // is then a pad in SPACE that does not vanish along an axis in which the scene happens to be thin
        float ext = 0.f;
        for (int j = 0; j < 3; ++j) ext = fmaxf(ext, (scene[9 + j] >= scene[6 + j]) ? (scene[9 + j] - scene[6 + j]) : 0.f);
        const float lo = scene[6 + k];
        const float cell = fmaxf(ext / 65520.0f, 1e-30f);
        scene[12 + k] = lo - 4.0f * cell;
        scene[15 + k] = cell;
    }
}
__device__ __forceinline__ unsigned quant_lo(float x, float lo, float cell) {
    int q = (int)floorf((x - lo) / cell);
    q = min(max(q, 0), 65535);
    while (q > 0 && fmaf((float)q, cell, lo) > x) --q;
    return (unsigned)max(q - 2, 0);
}
__device__ __forceinline__ unsigned quant_hi(float x, float lo, float cell) {
    int q = (int)ceilf((x - lo) / cell);
    q = min(max(q, 0), 65535);
    while (q < 65535 && fmaf((float)q, cell, lo) < x) ++q;
    return (unsigned)min(q + 2, 65535);
}
__global__ void quantize_nodes_kernel(const Node *__restrict__ nodes, int n_int, const float *__restrict__ scene,
                                      QNode *__restrict__ qnodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_int) return;
    const float *f = reinterpret_cast<const float *>(nodes + i);
    const int4 d = nodes[i].d;
    uint4 out[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const float *b = f + 6 * c;
        unsigned q[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        int ref = c == 0 ? d.x : d.y;
        if (b[0] >= IRGS_EMPTY_FAR) ref = IRGS_CHILD_NONE;
        else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                q[k] = quant_lo(b[k], scene[12 + k], scene[15 + k]);
                q[3 + k] = quant_hi(b[3 + k], scene[12 + k], scene[15 + k]);
            }
        }
        out[c] = make_uint4(q[0] | (q[3] << 16), q[1] | (q[4] << 16), q[2] | (q[5] << 16), (unsigned)ref);
    }
    qnodes[i].l = out[0];
    qnodes[i].r = out[1];
}

// Step 7: 4-wide traversal nodes (QNode4 in internal.cuh).  Topology part (build only): depth parity of every binary node.
__global__ void depth_parity_kernel(const int *__restrict__ node_parent, int n_int, int *__restrict__ even) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_int) return;
    int depth = 0, p = node_parent[i];
    while (p >= 0) { ++depth; p = node_parent[p >> 1]; }
    even[i] = (depth & 1) ? 0 : 1;
}

__device__ __forceinline__ uint4 quant_child(const float *b, int ref, const float *scene) {
    if (b[0] >= IRGS_EMPTY_FAR) return make_uint4(0u, 0u, 0u, (unsigned)IRGS_CHILD_NONE);
    unsigned q[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        q[k] = quant_lo(b[k], scene[12 + k], scene[15 + k]);
        q[3 + k] = quant_hi(b[3 + k], scene[12 + k], scene[15 + k]);
    }
    return make_uint4(q[0] | (q[3] << 16), q[1] | (q[4] << 16), q[2] | (q[5] << 16), (unsigned)ref);
}

// Topology part, default (build only): GREEDY COLLAPSE of the binary tree into 4-wide nodes.  A wide node starts as the two children
// of its binary node; while it has a free slot, the internal child with the largest surface area is replaced by its own two
// children.  The internal children that remain become wide nodes themselves.  Unlike the fixed fold of every other level, which
// leaves a slot empty wherever a child is a leaf (and two where both are), this fills the nodes wherever the subtree allows: fewer
// wide nodes, fewer visits.  Top-down in waves: wave w turns the wide roots at wide depth w into their child lists (`kids`: for each
// slot the binary parent * 2 + side that holds the slot's bounds and reference, -1 = empty) and queues the next wave.
__device__ __forceinline__ int slot_ref(const Node *__restrict__ nodes, int desc) {
    const int4 d = nodes[desc >> 1].d;
    return (desc & 1) ? d.y : d.x;
}
__global__ void wide_init_kernel(int *__restrict__ frontier, int *__restrict__ counts, int n_counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_counts) counts[i] = i == 0 ? 1 : 0;
    if (i == 0) frontier[0] = 0;   // the root
}
__global__ void wide_collapse_kernel(const Node *__restrict__ nodes, const int *__restrict__ fin, const int *__restrict__ cnt_in,
                                     int *__restrict__ fout, int *__restrict__ cnt_out, int4 *__restrict__ kids,
                                     int *__restrict__ is_root) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *cnt_in) return;
    const int n = fin[i];
    int desc[4] = {2 * n, 2 * n + 1, -1, -1};
    int m = 2;
    while (m < 4) {
        int best = -1;
        float best_a = -1.f;
        for (int k = 0; k < m; ++k) {
            const float *b = reinterpret_cast<const float *>(nodes + (desc[k] >> 1)) + 6 * (desc[k] & 1);
            if (slot_ref(nodes, desc[k]) >= 0 && b[0] < IRGS_EMPTY_FAR) {
                const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
                const float ar = dx * dy + dy * dz + dz * dx;
                if (ar > best_a) { best_a = ar; best = k; }
            }
        }
        if (best < 0) break;
        const int ref = slot_ref(nodes, desc[best]);
        desc[best] = 2 * ref;
        desc[m++] = 2 * ref + 1;
    }
    kids[n] = make_int4(desc[0], desc[1], desc[2], desc[3]);
    is_root[n] = 1;
    for (int k = 0; k < m; ++k) {
        const int ref = slot_ref(nodes, desc[k]);
        if (ref >= 0) fout[atomicAdd(cnt_out, 1)] = ref;   // (a subtree of invisible surfels is never entered, but gets valid nodes too)
    }
}

// Bounds part of the greedy collapse (every build and refit): one thread per wide root quantises the slots of its child list.
__global__ void quantize_wide_kids_kernel(const Node *__restrict__ nodes, const int *__restrict__ is_root,
                                          const int4 *__restrict__ kids, int n_int, const float *__restrict__ scene,
                                          QNode4 *__restrict__ wide) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_int || !is_root[i]) return;
    const int4 kd = kids[i];
    const int desc[4] = {kd.x, kd.y, kd.z, kd.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint4 out = make_uint4(0u, 0u, 0u, (unsigned)IRGS_CHILD_NONE);
        if (desc[k] >= 0) {
            const float *b = reinterpret_cast<const float *>(nodes + (desc[k] >> 1)) + 6 * (desc[k] & 1);
            out = quant_child(b, slot_ref(nodes, desc[k]), scene);
        }
        wide[i].c[k] = out;
    }
}

// Fixed fold (irgs_set_option("wide_fold", 1)), bounds part (every build and refit): one thread per even-depth binary node gathers its (up to four) grandchildren.
__global__ void quantize_wide_kernel(const Node *__restrict__ nodes, const int *__restrict__ even, int n_int,
                                     const float *__restrict__ scene, QNode4 *__restrict__ wide) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_int || !even[i]) return;
    const float *f = reinterpret_cast<const float *>(nodes + i);
    const int4 d = nodes[i].d;
    uint4 out[4];
    int m = 0;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int c = side == 0 ? d.x : d.y;
        if (c < 0 || f[6 * side] >= IRGS_EMPTY_FAR) {
            out[m++] = quant_child(f + 6 * side, c, scene);       // a leaf (or an empty slot) stays one slot
        } else {
            const float *g = reinterpret_cast<const float *>(nodes + c);
            const int4 dc = nodes[c].d;
            out[m++] = quant_child(g, dc.x, scene);
            out[m++] = quant_child(g + 6, dc.y, scene);
        }
    }
    for (; m < 4; ++m) out[m] = make_uint4(0u, 0u, 0u, (unsigned)IRGS_CHILD_NONE);
#pragma unroll
    for (int k = 0; k < 4; ++k) wide[i].c[k] = out[k];
}

// ------------------------------------------------------------------------------------------------ host side
template <typename T>
static bool realloc_dev(T *&p, size_t count) {
    if (p) cudaFree(p);
    p = nullptr;
    return check(cudaMalloc(&p, sizeof(T) * (count ? count : 1)), "cudaMalloc");
}

int lbvh_reserve(irgs_tracer *h, int64_t n) {
    if (n <= h->cap) return 0;
    int64_t c = n;
    if (!realloc_dev(h->nodes, (size_t)c) || !realloc_dev(h->qnodes, (size_t)c) || !realloc_dev(h->qnodes4, (size_t)c) ||
        !realloc_dev(h->even, (size_t)c) || !realloc_dev(h->boxes, (size_t)c * 6) || !realloc_dev(h->codes, (size_t)c) ||
        !realloc_dev(h->codes_alt, (size_t)c) || !realloc_dev(h->order, (size_t)c) || !realloc_dev(h->order_alt, (size_t)c) ||
        !realloc_dev(h->leaf_parent, (size_t)c) || !realloc_dev(h->node_parent, (size_t)c) ||
        !realloc_dev(h->flags, (size_t)c) || !realloc_dev(h->recs, (size_t)c) || !realloc_dev(h->inv_order, (size_t)c) || !realloc_dev(h->ploc_cid, (size_t)c * 2) ||
        !realloc_dev(h->ploc_box, (size_t)c * 2 * 8) || !realloc_dev(h->ploc_nn, (size_t)c) ||
        !realloc_dev(h->ploc_counts, (size_t)(c / PLOC_SB + 2)) || !realloc_dev(h->ploc_offs, (size_t)(c / PLOC_SB + 2) * 2) ||
        !realloc_dev(h->ploc_totals, 2) || !realloc_dev(h->wide_kids, (size_t)c) || !realloc_dev(h->wide_counts, (size_t)WIDE_WAVES + 2))
        return 1;
    int64_t tiles = (c + RS_TILE - 1) / RS_TILE;
    if (!realloc_dev(h->radix_hist, (size_t)tiles * 256)) return 1;
    h->radix_tiles_cap = tiles;
    h->cap = c;
    return 0;
}

int launch_bounds_from_proxy(irgs_tracer *h, const float *verts, int vps, cudaStream_t s) {
    int n = (int)h->n;
    scene_init_kernel<<<1, 32, 0, s>>>(reinterpret_cast<int *>(h->scene));
    bounds_from_proxy_kernel<<<(n + 255) / 256, 256, 0, s>>>(verts, vps, n, h->boxes, reinterpret_cast<int *>(h->scene));
    count_launch(2);
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int launch_bounds_from_surfels(irgs_tracer *h, const float *means, const float *opacity, const float *ru,
                               const float *rv, const float *normals, float alpha_min, cudaStream_t s) {
    int n = (int)h->n;
    scene_init_kernel<<<1, 32, 0, s>>>(reinterpret_cast<int *>(h->scene));
    bounds_from_surfels_kernel<<<(n + 255) / 256, 256, 0, s>>>(means, opacity, ru, rv, normals, alpha_min, n, h->boxes,
                                                               reinterpret_cast<int *>(h->scene));
    count_launch(2);
    IRGS_CHECK(cudaGetLastError());
    return 0;
}

int lbvh_build(irgs_tracer *h, bool refit_only, cudaStream_t s) {
    ++h->pack_epoch;   // a build changes the leaf order the records are packed in
    const int n = (int)h->n;
    int *scene_i = reinterpret_cast<int *>(h->scene);
    if (!refit_only) {
        morton_kernel<<<(n + 255) / 256, 256, 0, s>>>(h->boxes, scene_i, n, h->codes, h->order);
        count_launch();
        if (radix_sort_pairs(h->codes, h->codes_alt, h->order, h->order_alt, n, 4, h->radix_hist, s)) return 1;
        int n_int = n > 1 ? n - 1 : 1;
        h->tree_depth = 0;
        if (h->builder == 0 && n > 2) {
            if (ploc_build(h, s)) return 1;
        }
        if (h->builder != 0 || n <= 2 || h->tree_depth > 60) {   // a degenerate clustering falls back to the bounded-depth tree
            hierarchy_kernel<<<(n_int + 255) / 256, 256, 0, s>>>(h->codes, n, h->nodes, h->leaf_parent, h->node_parent);
            count_launch();
        }
        h->wide_fold_built = h->wide_fold;
        if (h->wide_fold_built) {
            depth_parity_kernel<<<(n_int + 255) / 256, 256, 0, s>>>(h->node_parent, n_int, h->even);
            count_launch();
        }
    }
    IRGS_CHECK(cudaMemsetAsync(h->flags, 0, sizeof(int) * (size_t)n, s));
    refit_kernel<<<(n + 255) / 256, 256, 0, s>>>(h->boxes, h->order, h->leaf_parent, h->node_parent, n, scene_i, h->nodes,
                                                 h->flags, h->scene + 6);
    const int n_internal = n > 1 ? n - 1 : 1;
    quant_frame_kernel<<<1, 32, 0, s>>>(h->scene);
    quantize_nodes_kernel<<<(n_internal + 255) / 256, 256, 0, s>>>(h->nodes, n_internal, h->scene, h->qnodes);
    if (h->wide_fold_built) {
        quantize_wide_kernel<<<(n_internal + 255) / 256, 256, 0, s>>>(h->nodes, h->even, n_internal, h->scene, h->qnodes4);
    } else {
        if (!refit_only) {
            // the collapse needs the bounds the refit just wrote; one wave per level of the wide tree, whose depth is at most the
            // binary tree's (PLOC: measured at build time; Karras: <= 62).  Waves past the last level find an empty queue.
            IRGS_CHECK(cudaMemsetAsync(h->even, 0, sizeof(int) * (size_t)n_internal, s));
            wide_init_kernel<<<1, 128, 0, s>>>(h->ploc_cid, h->wide_counts, WIDE_WAVES + 2);
            // (a PLOC tree that was kept has the depth measured at build time; everything else is the Karras tree: <= 62 levels)
            const int waves = ((h->builder == 0 && h->tree_depth > 0 && h->tree_depth <= 60) ? h->tree_depth : 62) + 2;
            for (int w = 0; w < waves && w < WIDE_WAVES; ++w) {
                int *fin = h->ploc_cid + (size_t)(w & 1) * (size_t)h->cap, *fout = h->ploc_cid + (size_t)((w + 1) & 1) * (size_t)h->cap;
                wide_collapse_kernel<<<(n_internal + 255) / 256, 256, 0, s>>>(h->nodes, fin, h->wide_counts + w, fout,
                                                                                h->wide_counts + w + 1, h->wide_kids, h->even);
            }
            count_launch(1 + waves);
        }
        quantize_wide_kids_kernel<<<(n_internal + 255) / 256, 256, 0, s>>>(h->nodes, h->even, h->wide_kids, n_internal, h->scene, h->qnodes4);
    }
    count_launch(4);
    IRGS_CHECK(cudaGetLastError());
    h->built = true;
    return 0;
}

}  // namespace irgs
