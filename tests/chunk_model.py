"""A float64 model of what the REFERENCE tracer does on the rays the strict golden comparison leaves out (test infrastructure).

The rays with >= 16 proxy crossings are handled by the reference in several 16-hit chunks (gaussiantrace_forward.cu:27-107):
after a chunk the ray restarts at `ray_o + t_16 * ray_d` with tmin = FLT_EPSILON, and whether the proxy triangle of the chunk's
LAST hit is found again from the new origin (its distance is ~0 +- one float rounding) is decided by rounding inside the
closed-source OptiX runtime.  Everything else the reference does on such a ray is deterministic and restated here:

  * candidates = FRONT-FACING proxy triangles (OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES) of the 20-triangle squashed icosahedra
    of scene/gaussian_model.py:712-723, crossed in (FLT_EPSILON, 100), ordered by TRIANGLE depth, 16 per chunk; the proxy
    is larger than the alpha >= alpha_min ellipse, so crossings with alpha < alpha_min occupy buffer slots without compositing;
  * per hit: the clamped plane depth, alpha, SH colour and compositing of gaussiantrace_forward.cu:50-101.

`explain_ray` enumerates the 2^(chunk boundaries) possibilities "last hit of the chunk seen again / not seen again" and reports
whether ONE of them reproduces the reference's recorded outputs to 1e-4: the test then asserts that the restated math plus this
one documented rounding quirk explains (nearly) every ray of the golden files, instead of only bounding their error.
"""
import numpy as np

T_SCENE_MAX, T_EPS = 100.0, 1.1920929e-07
SH_C0, SH_C1 = 0.28209479177387814, 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def sh_basis(deg, d):
    x, y, z = d
    Y = [SH_C0]
    if deg > 0:
        Y += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        Y += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if deg > 2:
        Y += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
              SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy), SH_C3[5] * z * (xx - yy),
              SH_C3[6] * x * (xx - 3 * yy)]
    return np.array(Y)


class Scene64:
    def __init__(self, sc, inp, alpha_min):
        from irgs_b200 import synth
        f = lambda t: t.double().numpy()                                                    # noqa: E731
        self.mu, self.n, self.ru, self.rv = f(inp["means3D"]), f(inp["normals"]), f(inp["ru"]), f(inp["rv"])
        self.op, self.shs, self.feat = f(inp["opacity"]).reshape(-1), f(inp["shs"]), f(inp["features"])
        vb, fb, _ = synth.proxy_mesh(sc, alpha_min)
        self.tri = f(vb)[fb.numpy()].reshape(-1, 20, 3, 3)          # [N, 20, 3 vertices, xyz]
        # bounding sphere of each proxy (NaN proxies: opacity < alpha_min)
        r = np.linalg.norm(f(vb).reshape(-1, 12, 3) - self.mu[:, None], axis=-1).max(1)
        self.rad = np.where(np.isfinite(r), r, -1.0)
        self.alpha_min = alpha_min


def proxy_hits(S, o, d):
    """Front-facing proxy-triangle crossings of ray (o, d) in (eps, 100): (triangle depth, surfel id) sorted by depth."""
    rel = S.mu - o
    tc = rel @ d
    near = np.nonzero((((rel - tc[:, None] * d) ** 2).sum(-1) <= (S.rad * 1.001) ** 2) & (S.rad > 0))[0]
    if near.size == 0:
        return np.zeros(0), np.zeros(0, np.int64)
    tri = S.tri[near]                                               # [M, 20, 3, 3]
    v0, e1, e2 = tri[:, :, 0], tri[:, :, 1] - tri[:, :, 0], tri[:, :, 2] - tri[:, :, 0]
    pv = np.cross(np.broadcast_to(d, e2.shape), e2)
    det = (e1 * pv).sum(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / det
        tv = o - v0
        u = (tv * pv).sum(-1) * inv
        qv = np.cross(tv, e1)
        v = (qv @ d) * inv
        t = (e2 * qv).sum(-1) * inv
    front = (np.cross(e1, e2) @ d) < 0                              # counter-clockwise seen from the ray = facing it
    ok = front & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > T_EPS) & (t < T_SCENE_MAX) & np.isfinite(t)
    idx_s, idx_t = np.nonzero(ok)
    ts, gs = t[idx_s, idx_t], near[idx_s]
    # one crossing per surfel (a convex proxy is entered once; a shared edge can report two triangles at the same depth)
    order = np.lexsort((gs, ts))
    ts, gs = ts[order], gs[order]
    keep = np.ones(len(gs), bool)
    seen = set()
    for i, g in enumerate(gs):
        keep[i] = g not in seen
        seen.add(g)
    return ts[keep], gs[keep]


def composite(S, o, d, seq, deg, back_culling, T_min, n_feat):
    """gaussiantrace_forward.cu:50-101 over the surfel sequence `seq`; returns (outputs[8 + n_feat], index after which T < T_min or
    len(seq))."""
    Y = sh_basis(deg, d)
    out = np.zeros(8 + n_feat)
    T = 1.0
    for i, g in enumerate(seq):
        n = S.n[g]
        cos = -(d @ n)
        m = 1.0 if cos > 0 else -1.0
        if m < 0 and back_culling:
            continue
        og, dg = n @ (o - S.mu[g]), n @ d
        depth = -og * dg / max(1e-6, dg * dg)
        pos = o + depth * d - S.mu[g]
        pu, pv = S.ru[g] @ pos, S.rv[g] @ pos
        alpha = min(0.99, S.op[g] * np.exp(-0.5 * (pu * pu + pv * pv)))
        if alpha < S.alpha_min:
            continue
        c = np.maximum(Y @ S.shs[g, :len(Y)] + 0.5, 0.0)
        w = T * alpha
        out[0:3] += w * c
        out[3:6] += w * m * n
        out[6] += w * depth
        out[7] += w
        if n_feat:
            out[8:] += w * S.feat[g, :n_feat]
        T *= 1 - alpha
        if T < T_min:
            return out, i + 1
    return out, len(seq)


def explain_ray(S, o, d, want, deg, back_culling, T_min, n_feat, tol=1e-4, max_boundaries=10):
    """want = the reference's recorded [color(3), normal(3), depth, alpha, features...].  Returns (explained, n_chunks, best error)."""
    ts, gs = proxy_hits(S, o, d)
    scale = np.ones(8 + n_feat)
    scale[6] = max(1.0, abs(want[6]))                               # depth: relative to max(1, |depth|) as in golden_util
    best = np.inf

    def run(prefix, pos, boundaries):
        """prefix: surfels processed so far (with duplicates); pos: next index of the sorted crossing list."""
        nonlocal best
        seq = prefix + list(gs[pos:pos + 16])
        out, stop = composite(S, o, d, seq, deg, back_culling, T_min, n_feat)
        if stop < len(seq) or pos + 16 >= len(gs) or boundaries >= max_boundaries:
            # the ray terminated inside what has been processed, ran out of crossings, or the enumeration budget is spent
            if stop >= len(seq) and pos + 16 < len(gs):
                out, _ = composite(S, o, d, prefix + list(gs[pos:]), deg, back_culling, T_min, n_feat)
            best = min(best, float(np.max(np.abs(out - want) / scale)))
            return
        # chunk boundary after seq: the next chunk starts either with the following crossing, or sees the last hit again (and
        # then holds only 15 new crossings)
        run(seq, pos + 16, boundaries + 1)
        seq_dup = seq + [seq[-1]]
        run_dup(seq_dup, pos + 16, boundaries + 1)

    def run_dup(prefix, pos, boundaries):
        nonlocal best
        seq = prefix + list(gs[pos:pos + 15])
        out, stop = composite(S, o, d, seq, deg, back_culling, T_min, n_feat)
        if stop < len(seq) or pos + 15 >= len(gs) or boundaries >= max_boundaries:
            if stop >= len(seq) and pos + 15 < len(gs):
                out, _ = composite(S, o, d, prefix + list(gs[pos:]), deg, back_culling, T_min, n_feat)
            best = min(best, float(np.max(np.abs(out - want) / scale)))
            return
        run(seq, pos + 15, boundaries + 1)
        run_dup(seq + [seq[-1]], pos + 15, boundaries + 1)

    run([], 0, 0)
    return best <= tol, (len(gs) + 15) // 16, best
