// BVH quality experiment (CPU, design tool only -- not part of the product or the oracle).
// Compares builders on the same surfel boxes and rays: node visits / leaf tests per ray for a clipped any-hit walk.
//   exp_bvh boxes.bin N rays.bin R mode radius
// boxes: N x 6 float (lo, hi); rays: R x 7 float (o, d, tclip).  mode 0 = LBVH (30-bit Morton), 1 = PLOC(radius)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

typedef struct { float lo[3], hi[3]; int left, right; } BNode;  // child >= 0 internal, < 0 leaf ~g
static float (*box)[6];
static BNode *nodes; static int n_nodes;

static uint32_t expand10(uint32_t v) { v &= 0x3ff; v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249; return v; }
typedef struct { uint64_t code; int g; } MKey;
static int mcmp(const void *a, const void *b) { const MKey *x = a, *y = b; if (x->code != y->code) return x->code < y->code ? -1 : 1; return x->g - y->g; }
static void child_box(int c, const float **lo, const float **hi) { if (c < 0) { *lo = box[~c]; *hi = box[~c] + 3; } else { *lo = nodes[c].lo; *hi = nodes[c].hi; } }
static void set_bounds(int id) { BNode *nd = &nodes[id]; const float *l0, *h0, *l1, *h1; child_box(nd->left, &l0, &h0); child_box(nd->right, &l1, &h1);
    for (int k = 0; k < 3; ++k) { nd->lo[k] = fminf(l0[k], l1[k]); nd->hi[k] = fmaxf(h0[k], h1[k]); } }
static int build_rec(const MKey *keys, int lo, int hi, int bit) {
    if (hi - lo == 1) return ~keys[lo].g;
    int split = -1;
    while (bit >= 0) { uint64_t mask = 1ull << bit;
        if ((keys[lo].code & mask) != (keys[hi - 1].code & mask)) { int a = lo, c = hi - 1; while (a + 1 < c) { int mid = (a + c) / 2; if (keys[mid].code & mask) c = mid; else a = mid; } split = c; break; }
        --bit; }
    if (split < 0) split = (lo + hi) / 2;
    int id = n_nodes++;
    int l = build_rec(keys, lo, split, bit - 1), r = build_rec(keys, split, hi, bit - 1);
    nodes[id].left = l; nodes[id].right = r; set_bounds(id);
    return id;
}
static float area2(const float *alo, const float *ahi, const float *blo, const float *bhi) {
    float e[3]; for (int k = 0; k < 3; ++k) e[k] = fmaxf(ahi[k], bhi[k]) - fminf(alo[k], blo[k]);
    return e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
}
static float area1(const float *lo, const float *hi) { float e[3]; for (int k = 0; k < 3; ++k) e[k] = hi[k] - lo[k]; return e[0] * e[1] + e[1] * e[2] + e[2] * e[0]; }

static int ploc(const MKey *keys, int n, int radius) {
    int *C = malloc(sizeof(int) * n), *C2 = malloc(sizeof(int) * n), *nn = malloc(sizeof(int) * n);
    for (int i = 0; i < n; ++i) C[i] = ~keys[i].g;
    int m = n, iters = 0;
    while (m > 1) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < m; ++i) {
            const float *lo, *hi; child_box(C[i], &lo, &hi);
            float best = INFINITY; int bj = -1;
            int a = i - radius < 0 ? 0 : i - radius, b = i + radius >= m ? m - 1 : i + radius;
            for (int j = a; j <= b; ++j) { if (j == i) continue; const float *l2, *h2; child_box(C[j], &l2, &h2);
                float ar = area2(lo, hi, l2, h2); if (ar < best) { best = ar; bj = j; } }
            nn[i] = bj;
        }
        int out = 0;
        for (int i = 0; i < m; ++i) {
            int j = nn[i];
            if (nn[j] == i) { if (i < j) { int id = n_nodes++; nodes[id].left = C[i]; nodes[id].right = C[j]; set_bounds(id); C2[out++] = id; } }
            else C2[out++] = C[i];
        }
        int *t = C; C = C2; C2 = t; m = out; ++iters;
    }
    fprintf(stderr, "ploc iters %d\n", iters);
    int root = C[0]; free(C); free(C2); free(nn); return root;
}

static inline int box_hit(const float *lo, const float *hi, const float *o, const float *inv, float tmax, float *tn) {
    float a = 0.f, b = tmax;
    for (int k = 0; k < 3; ++k) { float t0 = (lo[k] - o[k]) * inv[k], t1 = (hi[k] - o[k]) * inv[k]; a = fmaxf(a, fminf(t0, t1)); b = fminf(b, fmaxf(t0, t1)); }
    *tn = a; return a <= b;
}

// wide collapse: W-wide nodes made by repeatedly opening the child with the largest area
typedef struct { int nch; int ch[8]; } WNode;
static WNode *wn; static int n_w;
static int collapse(int root, int W) {
    int id = n_w++; int ch[8]; int nch = 2; ch[0] = nodes[root].left; ch[1] = nodes[root].right;
    while (nch < W) { int best = -1; float ba = -1; for (int i = 0; i < nch; ++i) if (ch[i] >= 0) { float a = area1(nodes[ch[i]].lo, nodes[ch[i]].hi); if (a > ba) { ba = a; best = i; } }
        if (best < 0) break; int c = ch[best]; ch[best] = nodes[c].left; ch[nch++] = nodes[c].right; }
    wn[id].nch = nch;
    for (int i = 0; i < nch; ++i) wn[id].ch[i] = ch[i] >= 0 ? collapse(ch[i], W) | 0 : ch[i];
    // remember original binary id for bounds: store separately
    return id;
}
static float (*wbox)[8][6];

int main(int argc, char **argv) {
    int N = atoi(argv[2]), R = atoi(argv[4]), mode = atoi(argv[5]), radius = argc > 6 ? atoi(argv[6]) : 16;
    int W = argc > 7 ? atoi(argv[7]) : 2; int FIXED = 0; if (W < 0) { W = -W; FIXED = 1; }
    box = malloc(sizeof(float[6]) * N); FILE *f = fopen(argv[1], "rb"); if (fread(box, 24, N, f) != (size_t)N) return 1; fclose(f);
    float (*rays)[7] = malloc(sizeof(float[7]) * R); f = fopen(argv[3], "rb"); if (fread(rays, 28, R, f) != (size_t)R) return 1; fclose(f);
    nodes = malloc(sizeof(BNode) * N);
    MKey *keys = malloc(sizeof(MKey) * N);
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int g = 0; g < N; ++g) for (int k = 0; k < 3; ++k) { float c = 0.5f * (box[g][k] + box[g][3 + k]); clo[k] = fminf(clo[k], c); chi[k] = fmaxf(chi[k], c); }
    for (int g = 0; g < N; ++g) { uint32_t q[3]; for (int k = 0; k < 3; ++k) { float c = 0.5f * (box[g][k] + box[g][3 + k]); float u = (c - clo[k]) / (chi[k] - clo[k]); int v = (int)(u * 1024.f); if (v > 1023) v = 1023; if (v < 0) v = 0; q[k] = v; }
        keys[g].code = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]); keys[g].g = g; }
    qsort(keys, N, sizeof(MKey), mcmp);
    double t0 = omp_get_wtime();
    int root = mode == 0 ? build_rec(keys, 0, N, 29) : ploc(keys, N, radius);
    double tb = omp_get_wtime() - t0;
    double sah = 0; float ra = area1(nodes[root].lo, nodes[root].hi);
    for (int i = 0; i < n_nodes; ++i) sah += area1(nodes[i].lo, nodes[i].hi) / ra;
    double sah_leaf = 0; for (int g = 0; g < N; ++g) sah_leaf += area1(box[g], box[g] + 3) / ra;
    long long V = 0, P = 0, maxV = 0; int maxsp = 0;
    if (W == 2) {
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : V, P) reduction(max : maxV, maxsp)
    for (int r = 0; r < R; ++r) {
        const float *o = rays[r], *d = rays[r] + 3; float tmax = rays[r][6];
        float inv[3]; for (int k = 0; k < 3; ++k) inv[k] = 1.0f / (fabsf(d[k]) > 1e-30f ? d[k] : 1e-30f);
        int stack[128], sp = 0; int cur = root; long long v = 0;
        while (1) { ++v; const BNode *nd = &nodes[cur];
            const float *l0, *h0, *l1, *h1; child_box(nd->left, &l0, &h0); child_box(nd->right, &l1, &h1);
            float tl, tr; int hl = box_hit(l0, h0, o, inv, tmax, &tl), hr = box_hit(l1, h1, o, inv, tmax, &tr);
            int next = -1 << 30;
            int cn = nd->left, cf = nd->right, hn = hl, hf = hr; if (hr && (!hl || tr < tl)) { cn = nd->right; cf = nd->left; hn = hr; hf = hl; }
            if (hn) { if (cn < 0) ++P; else next = cn; }
            if (hf) { if (cf < 0) ++P; else if (next == (-1 << 30)) next = cf; else stack[sp++] = cf; }
            if (sp > maxsp) maxsp = sp;
            if (next == (-1 << 30)) { if (sp == 0) break; next = stack[--sp]; }
            cur = next; }
        V += v; if (v > maxV) maxV = v;
    }
    } else {
        wn = malloc(sizeof(WNode) * N); wbox = malloc(sizeof(float[8][6]) * N); n_w = 0;
        // collapse needs child bounds: recompute from binary nodes during collapse -> do it iteratively here
        // (simple approach: run collapse, then fill wbox by a second pass storing binary ids)
        // Re-implement collapse to record boxes:
        int *stackb = malloc(sizeof(int) * 2 * N), *stackw = malloc(sizeof(int) * 2 * N); int sp = 0;
        n_w = 1; stackb[0] = root; stackw[0] = 0; sp = 1;
        while (sp) { --sp; int b = stackb[sp], w = stackw[sp]; int ch[8]; int nch = 2; ch[0] = nodes[b].left; ch[1] = nodes[b].right;
            if (FIXED) { int c0 = ch[0], c1 = ch[1]; nch = 0; if (c0 >= 0) { ch[nch++] = nodes[c0].left; ch[nch++] = nodes[c0].right; } else ch[nch++] = c0;
                if (c1 >= 0) { ch[nch++] = nodes[c1].left; ch[nch++] = nodes[c1].right; } else ch[nch++] = c1; }
            else
            while (nch < W) { int best = -1; float ba = -1; for (int i = 0; i < nch; ++i) if (ch[i] >= 0) { float a = area1(nodes[ch[i]].lo, nodes[ch[i]].hi); if (a > ba) { ba = a; best = i; } }
                if (best < 0) break; int c = ch[best]; ch[best] = nodes[c].left; ch[nch++] = nodes[c].right; }
            wn[w].nch = nch;
            for (int i = 0; i < nch; ++i) { const float *lo, *hi; child_box(ch[i], &lo, &hi); memcpy(wbox[w][i], lo, 12); memcpy(wbox[w][i] + 3, hi, 12);
                if (ch[i] >= 0) { int id = n_w++; wn[w].ch[i] = id; stackb[sp] = ch[i]; stackw[sp] = id; ++sp; } else wn[w].ch[i] = ch[i]; } }
        long long slots = 0; for (int i = 0; i < n_w; ++i) slots += wn[i].nch;
        fprintf(stderr, "wide nodes %d avg children %.2f\n", n_w, (double)slots / n_w);
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : V, P) reduction(max : maxV, maxsp)
        for (int r = 0; r < R; ++r) {
            const float *o = rays[r], *d = rays[r] + 3; float tmax = rays[r][6];
            float inv[3]; for (int k = 0; k < 3; ++k) inv[k] = 1.0f / (fabsf(d[k]) > 1e-30f ? d[k] : 1e-30f);
            int stack[256], s = 0; stack[s++] = 0; long long v = 0;
            while (s) { int cur = stack[--s]; ++v; const WNode *nd = &wn[cur];
                for (int i = 0; i < nd->nch; ++i) { float tn; if (box_hit(wbox[cur][i], wbox[cur][i] + 3, o, inv, tmax, &tn)) { if (nd->ch[i] < 0) ++P; else stack[s++] = nd->ch[i]; } }
                if (s > maxsp) maxsp = s; }
            V += v; if (v > maxV) maxV = v;
        }
    }
    printf("mode %d radius %d W %d: build %.2fs  SAH internal %.1f leaf %.1f | node visits/ray %.1f  leaf tests/ray %.1f  max visits %lld max stack %d\n", mode, radius, W, tb, sah, sah_leaf,
           (double)V / R, (double)P / R, maxV, maxsp);
    return 0;
}
