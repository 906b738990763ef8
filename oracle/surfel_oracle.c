/*
 * surfel_oracle.c -- CPU restatement of the IRGS surfel tracer math.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under irgs_b200/ may import, link or execute this file; it is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker and the
 * CPU baseline, never as the product path.
 *
 * PARITY PIN: the reference ships no tests / golden vectors for this path, and its
 * traversal lives in the closed-source OptiX runtime.  This oracle is pinned two ways:
 *   (1) tests/golden/ref_optix_*.npz -- outputs of the UNMODIFIED reference extension (oracle/build_ref.sh ->
 *       baseline/_ref) run on a B200 by oracle/gen_golden_ref.py, when libnvoptix is available on the box;
 *   (2) a float64 torch-autograd twin of the same math (tests/torch_twin.py) for the gradients.
 * If (1) is absent the header of DESIGN.md says "parity unpinned vs the OptiX runtime".
 *
 * What is restated (reference file:line, all under /root/reference/submodules/surfel_tracer/):
 *   forward  loop / plane hit / compositing / termination   src/optix/gaussiantrace_forward.cu:12-112
 *   k-nearest semantics of the any-hit buffer                src/optix/gaussiantrace_forward.cu:120-141
 *   backward gradient formulas                               src/optix/gaussiantrace_backward.cu:11-171
 *   SH evaluation and its backward                           src/optix/auxiliary.h:16-33, 52-89, 91-143
 *   constants MAX_BUFFER_SIZE / T_SCENE_MAX                  src/optix/auxiliary.h:10-12
 *
 * Deliberate, documented deviations:
 *   - candidate set = "alpha >= alpha_min at the ray/plane intersection and eps < t < 100" instead of the
 *     hexagonal proxy-triangle hit (identical up to a 7e-6-wide sliver of hits with alpha ~ alpha_min);
 *   - sort key = plane depth t with ties broken by surfel id (the reference sorts by proxy-triangle t, which
 *     differs from the plane depth by ~1e-6 and leaves exact ties to traversal order);
 *   - ray/surfel pairs with |n.d| < 1e-3 (where max(1e-6, (n.d)^2) clamps and the depth formula stops being the
 *     geometric plane hit) are never candidates;
 *   - one monotone pass over the ordered hit list (the reference restarts every 16 hits from o + t_last*d with
 *     tmin FLT_EPSILON fwd / 0 bwd, which can re-hit or skip the 16th surfel by rounding).
 *
 * The depth t is computed with an explicit, fixed sequence of IEEE-754 single-precision operations
 * (dot3 below: one multiply and two fused multiply-adds) so that the CUDA kernels, which use the same sequence
 * through __fmul_rn/__fmaf_rn/__fdiv_rn, obtain BIT-IDENTICAL depths and therefore an identical hit order even
 * for near ties.  Compile with -ffp-contract=off so gcc adds no fusions of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define T_SCENE_MAX 100.0f          /* auxiliary.h:11 */
#define T_EPS 1.1920929e-07f        /* FLT_EPSILON, gaussiantrace_forward.cu:38 */
#define ORACLE_K_MAX 64
/* k-buffer depth and termination-aware culling of the LBVH walk.  The CANONICAL structure that defines the
 * per-ray counters V, P, H is K = 16 (MAX_BUFFER_SIZE, auxiliary.h:10) without the opacity bound;
 * oracle_set_variant() exists to evaluate traversal designs on the CPU and never changes results. */
static int ORACLE_K = 16;
static int ORACLE_OPACITY_CULL = 0;
static float ORACLE_T_MIN_CULL = 0.0f;
void oracle_set_variant(int k, int opacity_cull, float t_min) {
    ORACLE_K = k < 1 ? 1 : (k > ORACLE_K_MAX ? ORACLE_K_MAX : k);
    ORACLE_OPACITY_CULL = opacity_cull;
    ORACLE_T_MIN_CULL = t_min;
}
#define MAX_FEATURE_SIZE 12         /* auxiliary.h:12 */

/* auxiliary.h:16-33 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

typedef struct { float x, y, z; } v3;

static inline v3 ld3(const float *p, int64_t i) { v3 r = {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; return r; }
/* fixed-order dot product shared bit-for-bit with the CUDA kernels */
static inline float dot3(v3 a, v3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

/* SH basis values Y_k(d), k < (deg+1)^2, such that colour = 0.5 + sum_k Y_k * sh[k]   (auxiliary.h:52-89) */
static void sh_basis(int deg, v3 d, float *Y) {
    Y[0] = SH_C0;
    if (deg > 0) {
        float x = d.x, y = d.y, z = d.z;
        Y[1] = -SH_C1 * y; Y[2] = SH_C1 * z; Y[3] = -SH_C1 * x;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            Y[4] = SH_C2[0] * xy; Y[5] = SH_C2[1] * yz; Y[6] = SH_C2[2] * (2.0f * zz - xx - yy);
            Y[7] = SH_C2[3] * xz; Y[8] = SH_C2[4] * (xx - yy);
            if (deg > 2) {
                Y[9]  = SH_C3[0] * y * (3.0f * xx - yy);
                Y[10] = SH_C3[1] * xy * z;
                Y[11] = SH_C3[2] * y * (4.0f * zz - xx - yy);
                Y[12] = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
                Y[13] = SH_C3[4] * x * (4.0f * zz - xx - yy);
                Y[14] = SH_C3[5] * z * (xx - yy);
                Y[15] = SH_C3[6] * x * (xx - 3.0f * yy);
            }
        }
    }
}

typedef struct {
    float t;      /* plane depth along the ray, measured from rays_o */
    int   g;      /* surfel id */
    float alpha;  /* min(0.99, opacity * G) */
    float G;      /* exp(-0.5 |p|^2) */
    float og, dg; /* n.(o-mu), n.d */
    v3    rel;    /* o - mu */
    v3    pos;    /* o + t d - mu */
    float pu, pv; /* (ru.pos, rv.pos) */
    float m;      /* +1 if the surfel faces the ray, else -1 */
} Hit;

/* Evaluate surfel g against ray (o,d).  Returns 1 if it is a compositing candidate
 * (gaussiantrace_forward.cu:61-81); *raw_alpha receives opacity*G whenever t is in range (for margins). */
static inline int eval_surfel(v3 o, v3 d, int g, const float *means, const float *opacity, const float *ru,
                              const float *rv, const float *normals, float alpha_min, int back_culling,
                              Hit *h, float *raw_alpha) {
    v3 mu = ld3(means, g), n = ld3(normals, g);
    v3 rel = {o.x - mu.x, o.y - mu.y, o.z - mu.z};
    float og = dot3(n, rel);
    float dg = dot3(n, d);
    float dg2 = dg * dg;
    float den = fmaxf(1e-6f, dg2);
    float t = (-og * dg) / den;
    *raw_alpha = -1.0f;
    /* grazing pairs (|n.d| < 1e-3): the clamped formula above no longer yields the geometric plane hit, so the
     * reference's candidate (a proxy-triangle hit at the TRUE depth) and its depth formula disagree there; such
     * pairs are excluded (a ~1e-6 fraction of hits, each with a 1000x foreshortened footprint). */
    if (!(dg2 >= 1e-6f)) return 0;
    if (!(t > T_EPS && t < T_SCENE_MAX)) return 0;
    float m = (-dg > 0.0f) ? 1.0f : -1.0f;
    v3 pos = {fmaf(t, d.x, rel.x), fmaf(t, d.y, rel.y), fmaf(t, d.z, rel.z)};
    float pu = dot3(ld3(ru, g), pos), pv = dot3(ld3(rv, g), pos);
    float G = expf(-0.5f * (pu * pu + pv * pv));
    float a_raw = opacity[g] * G;
    *raw_alpha = a_raw;
    if (m < 0.0f && back_culling) return 0;
    float alpha = fminf(0.99f, a_raw);
    if (alpha < alpha_min) return 0;
    h->t = t; h->g = g; h->alpha = alpha; h->G = G; h->og = og; h->dg = dg; h->rel = rel; h->pos = pos;
    h->pu = pu; h->pv = pv; h->m = m;
    return 1;
}

static int hit_less(const Hit *a, const Hit *b) { return a->t < b->t || (a->t == b->t && a->g < b->g); }
static int hit_cmp(const void *a, const void *b) {
    const Hit *x = (const Hit *)a, *y = (const Hit *)b;
    return hit_less(x, y) ? -1 : (hit_less(y, x) ? 1 : 0);
}

/* ------------------------------------------------------------------------------------------------------------
 * Canonical LBVH (the structure SURVEY.md 8d defines V, P, H on): binary, 30-bit Morton code of the AABB
 * centroid, one surfel per leaf, analytic ellipse AABBs, near-first stack traversal that culls a node once its
 * entry distance exceeds the current 16th-best depth, restarted every 16 hits like the reference's chunks.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    float lo[3], hi[3];
    int left, right; /* >=0: internal node index; <0: leaf, surfel id = ~child */
} BNode;

typedef struct {
    int n;        /* surfels */
    int n_nodes;  /* internal nodes */
    BNode *nodes;
    float (*box)[6]; /* per surfel */
    int root;     /* internal node index, or ~g if n==1 */
} Lbvh;

typedef struct { uint32_t code; int g; } MKey;
static int mkey_cmp(const void *a, const void *b) {
    const MKey *x = (const MKey *)a, *y = (const MKey *)b;
    if (x->code != y->code) return x->code < y->code ? -1 : 1;
    return x->g < y->g ? -1 : (x->g > y->g ? 1 : 0);
}
static uint32_t expand10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

/* Analytic bound of { x : (ru.(x-mu))^2 + (rv.(x-mu))^2 <= r^2, n.(x-mu) = 0 }, r^2 = 2 ln(opacity/alpha_min):
 * the locus where alpha >= alpha_min; the same support IRGS's proxy icosahedron is scaled to
 * (scene/gaussian_model.py:712-723).  General (non-orthogonal ru/rv) form via the in-plane 2x2 inverse. */
static int surfel_box(int g, const float *means, const float *opacity, const float *ru, const float *rv,
                      const float *normals, float alpha_min, float *box) {
    float op = opacity[g];
    if (!(op > alpha_min)) return 0;
    double r = sqrt(2.0 * log((double)op / alpha_min));
    v3 n = ld3(normals, g), a = ld3(ru, g), b = ld3(rv, g), mu = ld3(means, g);
    /* orthonormal basis of the plane */
    double nn = sqrt((double)n.x * n.x + (double)n.y * n.y + (double)n.z * n.z);
    if (!(nn > 0)) return 0;
    double nx = n.x / nn, ny = n.y / nn, nz = n.z / nn;
    double e1[3], e2[3];
    if (fabs(nx) < 0.6) { e1[0] = 0; e1[1] = -nz; e1[2] = ny; } else { e1[0] = -nz; e1[1] = 0; e1[2] = nx; }
    double l1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    e1[0] /= l1; e1[1] /= l1; e1[2] /= l1;
    e2[0] = ny * e1[2] - nz * e1[1]; e2[1] = nz * e1[0] - nx * e1[2]; e2[2] = nx * e1[1] - ny * e1[0];
    double m00 = a.x * e1[0] + a.y * e1[1] + a.z * e1[2], m01 = a.x * e2[0] + a.y * e2[1] + a.z * e2[2];
    double m10 = b.x * e1[0] + b.y * e1[1] + b.z * e1[2], m11 = b.x * e2[0] + b.y * e2[1] + b.z * e2[2];
    double det = m00 * m11 - m01 * m10;
    if (!(fabs(det) > 1e-30)) return 0;
    double i00 = m11 / det, i01 = -m01 / det, i10 = -m10 / det, i11 = m00 / det;
    const float muv[3] = {mu.x, mu.y, mu.z};
    for (int k = 0; k < 3; ++k) {
        double c0 = e1[k] * i00 + e2[k] * i10, c1 = e1[k] * i01 + e2[k] * i11;
        double h = r * sqrt(c0 * c0 + c1 * c1) * (1.0 + 1e-4) + 2e-6;
        box[k] = (float)(muv[k] - h); box[3 + k] = (float)(muv[k] + h);
    }
    return 1;
}

static int build_rec(Lbvh *b, const MKey *keys, int lo, int hi, int bit) {
    /* range [lo,hi) of sorted keys -> node; returns child reference */
    if (hi - lo == 1) return ~keys[lo].g;
    int split = -1;
    while (bit >= 0) {
        uint32_t mask = 1u << bit;
        if ((keys[lo].code & mask) != (keys[hi - 1].code & mask)) {
            int a = lo, c = hi - 1; /* first index with the bit set */
            while (a + 1 < c) { int mid = (a + c) / 2; if (keys[mid].code & mask) c = mid; else a = mid; }
            split = c; break;
        }
        --bit;
    }
    if (split < 0) split = (lo + hi) / 2; /* identical codes: median split */
    int id = b->n_nodes++;
    int l = build_rec(b, keys, lo, split, bit - 1), r = build_rec(b, keys, split, hi, bit - 1);
    BNode *nd = &b->nodes[id];
    nd->left = l; nd->right = r;
    for (int k = 0; k < 3; ++k) { nd->lo[k] = INFINITY; nd->hi[k] = -INFINITY; }
    int ch[2] = {l, r};
    for (int c = 0; c < 2; ++c) {
        const float *lo3, *hi3;
        if (ch[c] < 0) { lo3 = b->box[~ch[c]]; hi3 = lo3 + 3; } else { lo3 = b->nodes[ch[c]].lo; hi3 = b->nodes[ch[c]].hi; }
        for (int k = 0; k < 3; ++k) { nd->lo[k] = fminf(nd->lo[k], lo3[k]); nd->hi[k] = fmaxf(nd->hi[k], hi3[k]); }
    }
    return id;
}

void *oracle_lbvh_build(int n, const float *means, const float *opacity, const float *ru, const float *rv,
                        const float *normals, float alpha_min) {
    Lbvh *b = (Lbvh *)calloc(1, sizeof(Lbvh));
    b->n = n;
    b->box = malloc(sizeof(float[6]) * (size_t)(n > 0 ? n : 1));
    b->nodes = malloc(sizeof(BNode) * (size_t)(n > 1 ? n - 1 : 1));
    MKey *keys = malloc(sizeof(MKey) * (size_t)(n > 0 ? n : 1));
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int g = 0; g < n; ++g) {
        if (!surfel_box(g, means, opacity, ru, rv, normals, alpha_min, b->box[g])) {
            for (int k = 0; k < 3; ++k) { b->box[g][k] = INFINITY; b->box[g][3 + k] = -INFINITY; }
            continue;
        }
        for (int k = 0; k < 3; ++k) {
            float c = 0.5f * (b->box[g][k] + b->box[g][3 + k]);
            clo[k] = fminf(clo[k], c); chi[k] = fmaxf(chi[k], c);
        }
    }
    for (int g = 0; g < n; ++g) {
        keys[g].g = g;
        if (!(b->box[g][0] <= b->box[g][3])) { keys[g].code = 0x3fffffffu; continue; }
        uint32_t q[3];
        for (int k = 0; k < 3; ++k) {
            float c = 0.5f * (b->box[g][k] + b->box[g][3 + k]);
            float ext = chi[k] - clo[k];
            float u = ext > 0 ? (c - clo[k]) / ext : 0.0f;
            int v = (int)(u * 1024.0f); if (v < 0) v = 0; if (v > 1023) v = 1023;
            q[k] = (uint32_t)v;
        }
        keys[g].code = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    }
    qsort(keys, (size_t)n, sizeof(MKey), mkey_cmp);
    b->n_nodes = 0;
    b->root = n > 0 ? build_rec(b, keys, 0, n, 29) : 0;
    free(keys);
    return b;
}
void oracle_lbvh_free(void *p) {
    Lbvh *b = (Lbvh *)p; if (!b) return;
    free(b->box); free(b->nodes); free(b);
}

static inline int box_hit(const float *lo, const float *hi, v3 o, v3 inv, float tmin, float tmax, float *tenter) {
    float t0 = (lo[0] - o.x) * inv.x, t1 = (hi[0] - o.x) * inv.x;
    float a = fminf(t0, t1), b = fmaxf(t0, t1);
    t0 = (lo[1] - o.y) * inv.y; t1 = (hi[1] - o.y) * inv.y;
    a = fmaxf(a, fminf(t0, t1)); b = fminf(b, fmaxf(t0, t1));
    t0 = (lo[2] - o.z) * inv.z; t1 = (hi[2] - o.z) * inv.z;
    a = fmaxf(a, fminf(t0, t1)); b = fminf(b, fmaxf(t0, t1));
    a = fmaxf(a, tmin); b = fminf(b, tmax);
    *tenter = a;
    return a <= b;
}

/* Collect the <=16 nearest candidates strictly after (t_last, g_last), sorted ascending. */
static int collect_bvh(const Lbvh *b, v3 o, v3 d, float t_last, int g_last, const float *means,
                       const float *opacity, const float *ru, const float *rv, const float *normals,
                       float alpha_min, int back_culling, Hit *buf, int64_t *cnt /* V,P */, float T_start) {
    int nbuf = 0, term = 0;
    if (b->n == 0) return 0;
    v3 inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    int stack[128]; float stack_t[128]; int sp = 0;
    int cur = b->root; float cur_t = 0.0f;
    /* the slab test is padded so that it can never reject a surfel the exact arithmetic of eval_surfel accepts */
    const float pad = 1e-4f;
    for (;;) {
        float tmax = (nbuf == ORACLE_K || term) ? buf[nbuf - 1].t : T_SCENE_MAX;
        if (cur < 0) {
            int g = ~cur; Hit h; float raw;
            cnt[1]++;
            if (eval_surfel(o, d, g, means, opacity, ru, rv, normals, alpha_min, back_culling, &h, &raw)) {
                int after = h.t > t_last || (h.t == t_last && h.g > g_last);
                int full = (nbuf == ORACLE_K || term);
                if (after && (!full || hit_less(&h, &buf[nbuf - 1]))) {
                    int i = (nbuf < ORACLE_K) ? nbuf++ : ORACLE_K - 1;
                    term = 0;
                    while (i > 0 && hit_less(&h, &buf[i - 1])) { buf[i] = buf[i - 1]; --i; }
                    buf[i] = h;
                    if (ORACLE_OPACITY_CULL) {
                        /* the buffered hits alone already drive T below T_min at entry j: nothing beyond j can ever be
                         * composited, so the buffer (and the traversal range) ends there */
                        float T = T_start;
                        for (int j = 0; j < nbuf; ++j) {
                            T *= (1.0f - buf[j].alpha);
                            if (T < ORACLE_T_MIN_CULL) { nbuf = j + 1; term = 1; break; }
                        }
                    }
                }
            }
        } else if (cur_t <= tmax + pad) {
            const BNode *nd = &b->nodes[cur];
            int ch[2] = {nd->left, nd->right}; float te[2]; int ok[2];
            for (int c = 0; c < 2; ++c) {
                const float *lo3, *hi3;
                if (ch[c] < 0) { lo3 = b->box[~ch[c]]; hi3 = lo3 + 3; } else { lo3 = b->nodes[ch[c]].lo; hi3 = b->nodes[ch[c]].hi; }
                cnt[0]++;
                ok[c] = box_hit(lo3, hi3, o, inv, fmaxf(t_last, 0.0f) - pad, tmax + pad, &te[c]);
            }
            if (ok[0] && ok[1]) {
                int nearc = te[1] < te[0];
                stack[sp] = ch[!nearc]; stack_t[sp] = te[!nearc]; ++sp;
                cur = ch[nearc]; cur_t = te[nearc];
                continue;
            } else if (ok[0]) { cur = ch[0]; cur_t = te[0]; continue; }
            else if (ok[1]) { cur = ch[1]; cur_t = te[1]; continue; }
        }
        if (sp == 0) break;
        --sp; cur = stack[sp]; cur_t = stack_t[sp];
    }
    return nbuf;
}

static void composite_hit(const Hit *h, v3 d, int deg, int K, int S, const float *normals, const float *features,
                          const float *shs, float *T, float *C, float *N, float *D, float *O, float *F, float *c_out, float *w_out) {
    float Y[16];
    sh_basis(deg, d, Y);
    int nb = (deg + 1) * (deg + 1);
    float c[3] = {0, 0, 0};
    for (int k = 0; k < nb; ++k)
        for (int j = 0; j < 3; ++j) c[j] += Y[k] * shs[((int64_t)h->g * K + k) * 3 + j];
    for (int j = 0; j < 3; ++j) c[j] = fmaxf(c[j] + 0.5f, 0.0f);
    v3 n = ld3(normals, h->g);
    float w = *T * h->alpha;
    C[0] += w * c[0]; C[1] += w * c[1]; C[2] += w * c[2];
    N[0] += w * h->m * n.x; N[1] += w * h->m * n.y; N[2] += w * h->m * n.z;
    *D += w * h->t; *O += w;
    for (int j = 0; j < S; ++j) F[j] += w * features[(int64_t)h->g * S + j];
    *T *= (1.0f - h->alpha);
    if (c_out) { c_out[0] = c[0]; c_out[1] = c[1]; c_out[2] = c[2]; }
    if (w_out) *w_out = w;
}

/* Ordered candidate list of one ray, either brute force over all surfels (bvh == NULL) or through the canonical
 * LBVH in 16-hit passes.  Calls visit(hit) front to back until it returns 0.  */
typedef int (*visit_fn)(const Hit *h, void *ctx);

static void walk_ray(const Lbvh *bvh, int n_surf, v3 o, v3 d, const float *means, const float *opacity,
                     const float *ru, const float *rv, const float *normals, float alpha_min, int back_culling,
                     visit_fn visit, void *ctx, int64_t *cnt, float *alpha_margin, const float *T_cur) {
    if (!bvh) {
        Hit *all = NULL; int n_all = 0, cap = 0;
        for (int g = 0; g < n_surf; ++g) {
            Hit h; float raw;
            int ok = eval_surfel(o, d, g, means, opacity, ru, rv, normals, alpha_min, back_culling, &h, &raw);
            if (raw >= 0.0f && alpha_margin) {
                float mg = fabsf(raw - alpha_min) / alpha_min;
                if (mg < *alpha_margin) *alpha_margin = mg;
            }
            cnt[1]++;
            if (!ok) continue;
            if (n_all == cap) { cap = cap ? 2 * cap : 64; all = realloc(all, sizeof(Hit) * (size_t)cap); }
            all[n_all++] = h;
        }
        qsort(all, (size_t)n_all, sizeof(Hit), hit_cmp);
        for (int i = 0; i < n_all; ++i) if (!visit(&all[i], ctx)) break;
        free(all);
        return;
    }
    float t_last = -INFINITY; int g_last = -1;
    for (;;) {
        Hit buf[ORACLE_K_MAX];
        int nb = collect_bvh(bvh, o, d, t_last, g_last, means, opacity, ru, rv, normals, alpha_min, back_culling, buf, cnt, *T_cur);
        int stop = 0;
        for (int i = 0; i < nb; ++i) if (!visit(&buf[i], ctx)) { stop = 1; break; }
        if (stop || nb < ORACLE_K) break;
        t_last = buf[nb - 1].t; g_last = buf[nb - 1].g;
    }
}

typedef struct {
    v3 d; int deg, K, S; const float *normals, *features, *shs; float T_min;
    float T, C[3], N[3], D, O, F[MAX_FEATURE_SIZE];
    int n_hits; int *hits; int hit_cap; float T_margin; float min_dt; float last_t;
} FwdCtx;

static int fwd_visit(const Hit *h, void *p) {
    FwdCtx *c = (FwdCtx *)p;
    composite_hit(h, c->d, c->deg, c->K, c->S, c->normals, c->features, c->shs, &c->T, c->C, c->N, &c->D, &c->O, c->F, NULL, NULL);
    if (c->hits && c->n_hits < c->hit_cap) c->hits[c->n_hits] = h->g;
    if (c->n_hits > 0) { float dt = h->t - c->last_t; if (dt < c->min_dt) c->min_dt = dt; }
    c->last_t = h->t;
    c->n_hits++;
    float mg = fabsf(c->T - c->T_min) / c->T_min;
    if (mg < c->T_margin) c->T_margin = mg;
    return !(c->T < c->T_min);
}

/* Forward trace of n_rays rays.  Layouts as SURVEY.md 8a': all float32 row-major.
 *   out_hits     [n_rays, hit_cap] int32 surfel ids in compositing order (may be NULL)
 *   out_hit_count[n_rays]          int32 number of composited hits (may exceed hit_cap)
 *   out_margin   [n_rays, 3]       (relative distance of any in-range opacity*G to alpha_min [brute force only],
 *                                   relative distance of any running T to T_min, smallest gap between
 *                                   consecutive composited depths); may be NULL
 *   counters     [3] int64         sums over rays of: boxes tested (V), surfel tests (P), composited hits (H)
 *   bvh          handle from oracle_lbvh_build, or NULL for brute force over all surfels            */
int oracle_trace_forward(int64_t n_rays, int n_surf, int S, int K, int deg, int back_culling, float alpha_min,
                         float T_min, const float *rays_o, const float *rays_d, const float *means,
                         const float *opacity, const float *ru, const float *rv, const float *normals,
                         const float *features, const float *shs, float *out_color, float *out_normal,
                         float *out_feature, float *out_depth, float *out_alpha, int *out_hit_count,
                         int *out_hits, int hit_cap, float *out_margin, int64_t *counters, const void *bvh) {
    if (S > MAX_FEATURE_SIZE || K < (deg + 1) * (deg + 1) || deg < 0 || deg > 3) return 1;
    int64_t V = 0, P = 0, H = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : V, P, H)
    for (int64_t r = 0; r < n_rays; ++r) {
        FwdCtx c; memset(&c, 0, sizeof c);
        v3 o = ld3(rays_o, r); c.d = ld3(rays_d, r);
        c.deg = deg; c.K = K; c.S = S; c.normals = normals; c.features = features; c.shs = shs; c.T_min = T_min;
        c.T = 1.0f; c.hits = out_hits ? out_hits + r * hit_cap : NULL; c.hit_cap = hit_cap;
        c.T_margin = INFINITY; c.min_dt = INFINITY;
        int64_t cnt[2] = {0, 0}; float am = INFINITY;
        walk_ray((const Lbvh *)bvh, n_surf, o, c.d, means, opacity, ru, rv, normals, alpha_min, back_culling,
                 fwd_visit, &c, cnt, &am, &c.T);
        for (int j = 0; j < 3; ++j) { out_color[3 * r + j] = c.C[j]; out_normal[3 * r + j] = c.N[j]; }
        for (int j = 0; j < S; ++j) out_feature[r * S + j] = c.F[j];
        out_depth[r] = c.D; out_alpha[r] = c.O;
        if (out_hit_count) out_hit_count[r] = c.n_hits;
        if (out_margin) { out_margin[3 * r] = am; out_margin[3 * r + 1] = c.T_margin; out_margin[3 * r + 2] = c.min_dt; }
        V += cnt[0]; P += cnt[1]; H += c.n_hits;
    }
    if (counters) { counters[0] = V; counters[1] = P; counters[2] = H; }
    return 0;
}

typedef struct {
    v3 o, d; int deg, K, S; const float *ru, *rv, *normals, *opacity, *features, *shs; float T_min;
    float T, C[3], N[3], D, O, F[MAX_FEATURE_SIZE];
    float Cf[3], Nf[3], Df, Of, Ff[MAX_FEATURE_SIZE];
    float gC[3], gN[3], gD, gO, gF[MAX_FEATURE_SIZE];
    double g_o[3], g_d[3];
    double *G_means, *G_opacity, *G_ru, *G_rv, *G_normals, *G_features, *G_shs; /* shared across rays */
} BwdCtx;

static inline void add_d(double *p, double v) {
#pragma omp atomic
    *p += v;
}

/* gaussiantrace_backward.cu:61-166, formulas verbatim (including: no derivative masks for min(0.99,.) and
 * max(.,0); no SH-direction term; division by the raw d_g).  Per-hit terms are formed in float like the
 * reference; only the cross-ray/cross-hit accumulation is done in double so the oracle's sums carry no
 * ordering noise (the reference's float atomics make its own sums nondeterministic at ~1e-6 relative). */
static int bwd_visit(const Hit *h, void *p) {
    BwdCtx *c = (BwdCtx *)p;
    float col[3], w;
    composite_hit(h, c->d, c->deg, c->K, c->S, c->normals, c->features, c->shs, &c->T, c->C, c->N, &c->D, &c->O,
                  c->F, col, &w);
    const int g = h->g;
    const float alpha = h->alpha, T = c->T; /* T already multiplied by (1-alpha), backward.cu:112 */
    v3 n = ld3(c->normals, g), a = ld3(c->ru, g), b = ld3(c->rv, g);
    const float m = h->m, t = h->t;
    float nf[3] = {m * n.x, m * n.y, m * n.z};
    float dL_dalpha = c->gC[0] * (T * col[0] - (c->Cf[0] - c->C[0])) + c->gC[1] * (T * col[1] - (c->Cf[1] - c->C[1])) +
                      c->gC[2] * (T * col[2] - (c->Cf[2] - c->C[2])) +
                      c->gN[0] * (T * nf[0] - (c->Nf[0] - c->N[0])) + c->gN[1] * (T * nf[1] - (c->Nf[1] - c->N[1])) +
                      c->gN[2] * (T * nf[2] - (c->Nf[2] - c->N[2])) +
                      c->gD * (T * t - (c->Df - c->D)) + c->gO * (1.0f - c->Of);
    for (int j = 0; j < c->S; ++j)
        dL_dalpha += c->gF[j] * (T * c->features[(int64_t)g * c->S + j] - (c->Ff[j] - c->F[j]));
    dL_dalpha /= (1.0f - alpha);
    /* SH coefficients: basis * dL/dc, no clamp mask (auxiliary.h:91-143) */
    float Y[16]; sh_basis(c->deg, c->d, Y);
    int nb = (c->deg + 1) * (c->deg + 1);
    for (int k = 0; k < nb; ++k)
        for (int j = 0; j < 3; ++j) add_d(&c->G_shs[((int64_t)g * c->K + k) * 3 + j], (double)(Y[k] * (c->gC[j] * w)));
    float op = c->opacity[g];
    float dL_do = dL_dalpha * h->G;
    float dL_dG = dL_dalpha * op;
    float dpu = -dL_dG * h->G * h->pu, dpv = -dL_dG * h->G * h->pv;
    v3 pos = h->pos;
    float dru[3] = {dpu * pos.x, dpu * pos.y, dpu * pos.z};
    float drv[3] = {dpv * pos.x, dpv * pos.y, dpv * pos.z};
    float dpos[3] = {dpu * a.x + dpv * b.x, dpu * a.y + dpv * b.y, dpu * a.z + dpv * b.z};
    float dL_dd = c->gD * w + (dpos[0] * c->d.x + dpos[1] * c->d.y + dpos[2] * c->d.z);
    float dL_dog = -dL_dd / h->dg;
    float dL_ddg = dL_dd * h->og / fmaxf(1e-6f, h->dg * h->dg);
    float rel[3] = {h->rel.x, h->rel.y, h->rel.z}; /* ray_o_mean3D, backward.cu:84 */
    float dn[3] = {m * c->gN[0] * w + dL_ddg * c->d.x + dL_dog * rel[0],
                   m * c->gN[1] * w + dL_ddg * c->d.y + dL_dog * rel[1],
                   m * c->gN[2] * w + dL_ddg * c->d.z + dL_dog * rel[2]};
    float nn[3] = {n.x, n.y, n.z};
    for (int j = 0; j < 3; ++j) {
        add_d(&c->G_means[3 * (int64_t)g + j], (double)(-dpos[j] - dL_dog * nn[j]));
        add_d(&c->G_ru[3 * (int64_t)g + j], (double)dru[j]);
        add_d(&c->G_rv[3 * (int64_t)g + j], (double)drv[j]);
        add_d(&c->G_normals[3 * (int64_t)g + j], (double)dn[j]);
        c->g_o[j] += (double)(dpos[j] + dL_dog * nn[j]);
        c->g_d[j] += (double)(t * dpos[j] + dL_ddg * nn[j]);
    }
    add_d(&c->G_opacity[g], (double)dL_do);
    for (int j = 0; j < c->S; ++j) add_d(&c->G_features[(int64_t)g * c->S + j], (double)(c->gF[j] * w));
    return !(c->T < c->T_min);
}

/* Backward.  color..alpha are the forward outputs (saved tensors); gout_* the incoming gradients; the nine grad
 * outputs are OVERWRITTEN (not accumulated).  Rays with alpha == 0 are skipped (gaussiantrace_backward.cu:13-14). */
int oracle_trace_backward(int64_t n_rays, int n_surf, int S, int K, int deg, int back_culling, float alpha_min,
                          float T_min, const float *rays_o, const float *rays_d, const float *means,
                          const float *opacity, const float *ru, const float *rv, const float *normals,
                          const float *features, const float *shs, const float *color, const float *normal,
                          const float *feature, const float *depth, const float *alpha, const float *gout_color,
                          const float *gout_normal, const float *gout_feature, const float *gout_depth,
                          const float *gout_alpha, float *grad_rays_o, float *grad_rays_d, float *grad_means,
                          float *grad_opacity, float *grad_ru, float *grad_rv, float *grad_normals,
                          float *grad_features, float *grad_shs, const void *bvh) {
    if (S > MAX_FEATURE_SIZE || K < (deg + 1) * (deg + 1) || deg < 0 || deg > 3) return 1;
    size_t n = (size_t)n_surf;
    double *G_means = calloc(3 * n + 1, 8), *G_op = calloc(n + 1, 8), *G_ru = calloc(3 * n + 1, 8),
           *G_rv = calloc(3 * n + 1, 8), *G_n = calloc(3 * n + 1, 8), *G_f = calloc(n * (size_t)S + 1, 8),
           *G_sh = calloc(n * (size_t)K * 3 + 1, 8);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n_rays; ++r) {
        for (int j = 0; j < 3; ++j) { grad_rays_o[3 * r + j] = 0.0f; grad_rays_d[3 * r + j] = 0.0f; }
        if (alpha[r] == 0.0f) continue;
        BwdCtx c; memset(&c, 0, sizeof c);
        c.o = ld3(rays_o, r); c.d = ld3(rays_d, r);
        c.deg = deg; c.K = K; c.S = S; c.ru = ru; c.rv = rv; c.normals = normals; c.opacity = opacity;
        c.features = features; c.shs = shs; c.T_min = T_min; c.T = 1.0f;
        for (int j = 0; j < 3; ++j) {
            c.Cf[j] = color[3 * r + j]; c.Nf[j] = normal[3 * r + j];
            c.gC[j] = gout_color[3 * r + j]; c.gN[j] = gout_normal[3 * r + j];
        }
        for (int j = 0; j < S; ++j) { c.Ff[j] = feature[r * S + j]; c.gF[j] = gout_feature[r * S + j]; }
        c.Df = depth[r]; c.Of = alpha[r]; c.gD = gout_depth[r]; c.gO = gout_alpha[r];
        c.G_means = G_means; c.G_opacity = G_op; c.G_ru = G_ru; c.G_rv = G_rv; c.G_normals = G_n;
        c.G_features = G_f; c.G_shs = G_sh;
        int64_t cnt[2] = {0, 0};
        walk_ray((const Lbvh *)bvh, n_surf, c.o, c.d, means, opacity, ru, rv, normals, alpha_min, back_culling,
                 bwd_visit, &c, cnt, NULL, &c.T);
        for (int j = 0; j < 3; ++j) { grad_rays_o[3 * r + j] = (float)c.g_o[j]; grad_rays_d[3 * r + j] = (float)c.g_d[j]; }
    }
    for (size_t i = 0; i < 3 * n; ++i) {
        grad_means[i] = (float)G_means[i]; grad_ru[i] = (float)G_ru[i]; grad_rv[i] = (float)G_rv[i];
        grad_normals[i] = (float)G_n[i];
    }
    for (size_t i = 0; i < n; ++i) grad_opacity[i] = (float)G_op[i];
    for (size_t i = 0; i < n * (size_t)S; ++i) grad_features[i] = (float)G_f[i];
    for (size_t i = 0; i < n * (size_t)K * 3; ++i) grad_shs[i] = (float)G_sh[i];
    free(G_means); free(G_op); free(G_ru); free(G_rv); free(G_n); free(G_f); free(G_sh);
    return 0;
}

/* Per-surfel analytic AABBs (for tests of the CUDA bounds kernel). */
int oracle_surfel_boxes(int n_surf, const float *means, const float *opacity, const float *ru, const float *rv,
                        const float *normals, float alpha_min, float *boxes /* [n,6] */) {
    for (int g = 0; g < n_surf; ++g)
        if (!surfel_box(g, means, opacity, ru, rv, normals, alpha_min, boxes + 6 * g))
            for (int k = 0; k < 3; ++k) { boxes[6 * g + k] = INFINITY; boxes[6 * g + 3 + k] = -INFINITY; }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
