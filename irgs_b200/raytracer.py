"""Python-facing tracer: the same surface as the reference's surfel_tracer/raytracer.py (121 lines), backed by the
C ABI of libirgs_b200.so instead of the OptiX extension.

Reference being mirrored (all /root/reference/submodules/surfel_tracer/surfel_tracer/raytracer.py):
  * class GaussianTracer(transmittance_min=0.001): .impl, .transmittance_min                    (:69-72)
  * build_bvh(vertices_b, faces_b, gs_idxs) / update_bvh(same)                                   (:74-82)
  * trace(rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs, alpha_min, deg=3,
          back_culling=False) -> (color, normal, feature, depth, alpha)                          (:84-122)
  * _GaussianTrace autograd.Function: grads for rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs (:5-66)

Differences that a caller cannot observe in the returned values: no intersection-test / mask / compaction pre-pass
(rays that hit nothing simply come back as zeros, :103-114), no host synchronisation anywhere on the trace path, and
the backward replays the hit list saved by the forward instead of re-tracing.  Extras (not in the reference):
`last_hit_count`, `build_from_surfels`/`update_from_surfels`, deferred gradient accumulation with a single
all-reduce for ray-sharded multi-GPU runs (`accumulate_grads`, `flush_grads`).
"""
import ctypes

import torch

from . import _lib

GRAD_STRIDE = 64  # IRGS_GRAD_STRIDE


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_f32(name, t, device):
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if t.device != device:
        raise ValueError(f"{name} is on {t.device}, the tracer's acceleration structure is on {device}")


def _alloc_outputs(B, S, dev):
    """The five outputs and the hit counts of a trace call as views of ONE allocation, back to back in the order the native
    forward zero-fills with a single memset (color, normal, feature, depth, alpha, hit_count)."""
    buf = torch.empty(B * (9 + S), device=dev, dtype=torch.float32)
    o = 0
    views = []
    for width in (3, 3, S, 1, 1, 1):
        views.append(buf[o:o + B * width])
        o += B * width
    color, normal, feature, depth, alpha, count = views
    return (color.view(B, 3), normal.view(B, 3), feature.view(B, S), depth, alpha, count.view(torch.int32))


class _Impl:
    """Owns the native handle (the reference's `_C.create_gaussiantracer()` object, bindings.cu:101-103)."""

    def __init__(self, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("irgs_b200 has no CPU path; a CUDA (sm_100a) device is required")
        h = ctypes.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        _lib.check(self.lib.irgs_tracer_create(ctypes.byref(h), idx))
        self.h = h
        # every forward of this package allocates its outputs through _alloc_outputs: one block, one memset
        _lib.check(self.lib.irgs_set_option(h, b"contiguous_outputs", 1))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.irgs_tracer_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _reuse_records(ctx, impl):
    """Backward of a node whose forward recorded ctx.pack_epoch: if nothing has packed or rebuilt since (autograd itself guarantees
    that the saved arrays are unmodified) the packed records are still the ones of these arrays and the next backward call skips
    its own pack."""
    if impl.lib.irgs_get_info(impl.h, b"pack_epoch") == getattr(ctx, "pack_epoch", -1):
        impl.lib.irgs_set_option(impl.h, b"skip_next_pack", 1)


class _GaussianTrace(torch.autograd.Function):
    """raytracer.py:5-66."""

    @staticmethod
    def forward(ctx, tracer, rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs, alpha_min, deg,
                back_culling):
        impl = tracer.impl
        dev = impl.device
        B = rays_o.shape[0]
        S = features.shape[-1]
        K = shs.shape[1]
        color, normal, feature, depth, alpha, hit_count = _alloc_outputs(B, S, dev)
        need_grad = any(ctx.needs_input_grad)
        cap = tracer.hit_cap if need_grad else 0
        hits = torch.empty(B, cap, device=dev, dtype=torch.int32) if cap > 0 else None
        _lib.check(impl.lib.irgs_trace_forward(
            impl.h, B, S, K, deg, _ptr(rays_o), _ptr(rays_d), _ptr(means3D), _ptr(opacity), _ptr(ru), _ptr(rv),
            _ptr(normals), _ptr(features), _ptr(shs), _ptr(color), _ptr(normal), _ptr(feature), _ptr(depth),
            _ptr(alpha), _ptr(hit_count), _ptr(hits), cap, alpha_min, tracer.transmittance_min, int(back_culling),
            _stream(dev)))
        tracer.last_hit_count = hit_count
        ctx.tracer = tracer
        ctx.pack_epoch = impl.lib.irgs_get_info(impl.h, b"pack_epoch") if need_grad else -1
        ctx.alpha_min, ctx.deg, ctx.back_culling, ctx.cap = alpha_min, deg, back_culling, cap
        ctx.transmittance_min = tracer.transmittance_min
        ctx.save_for_backward(rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs, color, normal,
                              feature, depth, alpha, hit_count, hits if hits is not None else hit_count)
        ctx.mark_non_differentiable(hit_count)
        return color, normal, feature, depth, alpha, hit_count

    @staticmethod
    def backward(ctx, g_color, g_normal, g_feature, g_depth, g_alpha, _g_count):
        (rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs, color, normal, feature, depth, alpha,
         hit_count, hits) = ctx.saved_tensors
        tracer = ctx.tracer
        impl = tracer.impl
        dev = impl.device
        B, N, S, K = rays_o.shape[0], means3D.shape[0], features.shape[-1], shs.shape[1]
        g_color, g_normal, g_feature = g_color.contiguous(), g_normal.contiguous(), g_feature.contiguous()
        g_depth, g_alpha = g_depth.contiguous(), g_alpha.contiguous()
        grad_rays_o = torch.empty_like(rays_o)
        grad_rays_d = torch.empty_like(rays_d)
        deferred = tracer.accumulate_grads
        if deferred:
            fused, gfeat = tracer._grad_buffers(N, S)
        else:
            fused = torch.zeros(N, GRAD_STRIDE, device=dev, dtype=torch.float32)
            gfeat = torch.zeros(N, S, device=dev, dtype=torch.float32)
        have_list = ctx.cap > 0
        _reuse_records(ctx, impl)
        _lib.check(impl.lib.irgs_trace_backward(
            impl.h, B, S, K, ctx.deg, _ptr(rays_o), _ptr(rays_d), _ptr(means3D), _ptr(opacity), _ptr(ru), _ptr(rv),
            _ptr(normals), _ptr(features), _ptr(shs), _ptr(color), _ptr(normal), _ptr(feature), _ptr(depth),
            _ptr(alpha), _ptr(hit_count) if have_list else ctypes.c_void_p(0),
            _ptr(hits) if have_list else ctypes.c_void_p(0), ctx.cap, _ptr(g_color), _ptr(g_normal), _ptr(g_feature),
            _ptr(g_depth), _ptr(g_alpha), _ptr(grad_rays_o), _ptr(grad_rays_d), _ptr(fused), _ptr(gfeat),
            ctx.alpha_min, ctx.transmittance_min, int(ctx.back_culling), _stream(dev)))
        if deferred:
            surf = (None,) * 7
        else:
            surf = tracer._unpack(fused, gfeat, opacity.shape, K)
        return (None, grad_rays_o, grad_rays_d) + surf + (None, None, None)


class GaussianTracer:
    """raytracer.py:69-122 (see the module docstring for the mapping)."""

    def __init__(self, transmittance_min=0.001, device=None, hit_cap=96):
        if not torch.cuda.is_available():
            raise RuntimeError("irgs_b200.GaussianTracer needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.impl = _Impl(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        self.transmittance_min = transmittance_min
        if hit_cap % 4 != 0 or hit_cap < 0:
            raise ValueError("hit_cap must be a non-negative multiple of 4")
        self.hit_cap = hit_cap
        self.last_hit_count = None
        self.accumulate_grads = False
        self._side_streams = None
        self._fused = None
        self._gfeat = None
        self._n_proxy = None
        self.faces_shape = None
        self.gs_idxs = None

    # ------------------------------------------------------------------ acceleration structure
    def _proxy_boxes(self, vertices_b, faces_b, gs_idxs):
        """Proxy mesh -> what irgs_build_from_proxy consumes: (vertex array, surfel count, verts per surfel)."""
        dev = self.impl.device
        _check_f32("vertices_b", vertices_b, dev)
        v = vertices_b.contiguous()
        n_tri = faces_b.shape[0]
        if v.shape[0] % 12 == 0 and n_tri == (v.shape[0] // 12) * 20 and gs_idxs.shape[0] == n_tri:
            # IRGS layout (scene/gaussian_model.py:712-723): 12 consecutive vertices / 20 faces per surfel
            return v, v.shape[0] // 12, 12
        # general proxy mesh: per-surfel AABB over its triangles' vertices, handed over as a 2-vertex "proxy"
        n = int(gs_idxs.max().item()) + 1
        tri = v[faces_b.reshape(-1)].reshape(n_tri, 3, 3)
        idx = gs_idxs.long()[:, None, None].expand(-1, 3, 3).reshape(-1, 3)
        flat = tri.reshape(-1, 3)
        lo = torch.full((n, 3), float("inf"), device=dev).scatter_reduce(0, idx, flat, "amin")
        hi = torch.full((n, 3), float("-inf"), device=dev).scatter_reduce(0, idx, flat, "amax")
        return torch.stack([lo, hi], 1).reshape(-1, 3).contiguous(), n, 2

    def build_bvh(self, vertices_b, faces_b, gs_idxs):
        v, n, vps = self._proxy_boxes(vertices_b, faces_b, gs_idxs)
        self.faces_shape = tuple(faces_b.shape)
        self.gs_idxs = gs_idxs
        _lib.check(self.impl.lib.irgs_build_from_proxy(self.impl.h, _ptr(v), n, vps, _stream(self.impl.device)))

    def update_bvh(self, vertices_b, faces_b, gs_idxs):
        # the reference compares every face index on the device and syncs (raytracer.py:80); the topology
        # contract is checked here on shapes only, the native refit re-checks the surfel count
        assert self.faces_shape == tuple(faces_b.shape), "Update bvh must keep the triangle id not change~"
        v, n, vps = self._proxy_boxes(vertices_b, faces_b, gs_idxs)
        self.gs_idxs = gs_idxs
        _lib.check(self.impl.lib.irgs_refit_from_proxy(self.impl.h, _ptr(v), n, vps, _stream(self.impl.device)))

    def _surfel_args(self, means3D, opacity, ru, rv, normals):
        dev = self.impl.device
        ts = [t.detach().contiguous() for t in (means3D, opacity, ru, rv, normals)]
        for name, t in zip(("means3D", "opacity", "ru", "rv", "normals"), ts):
            _check_f32(name, t, dev)
        return ts

    def build_from_surfels(self, means3D, opacity, ru, rv, normals, alpha_min):
        """Native build: analytic elliptical bounds from the tracer's own inputs (no proxy mesh)."""
        ts = self._surfel_args(means3D, opacity, ru, rv, normals)
        _lib.check(self.impl.lib.irgs_build_from_surfels(self.impl.h, *[_ptr(t) for t in ts], ts[0].shape[0],
                                                         alpha_min, _stream(self.impl.device)))

    def update_from_surfels(self, means3D, opacity, ru, rv, normals, alpha_min):
        ts = self._surfel_args(means3D, opacity, ru, rv, normals)
        _lib.check(self.impl.lib.irgs_refit_from_surfels(self.impl.h, *[_ptr(t) for t in ts], ts[0].shape[0],
                                                         alpha_min, _stream(self.impl.device)))

    def num_surfels(self):
        return int(self.impl.lib.irgs_num_surfels(self.impl.h))

    def bounds(self):
        """(per-surfel bounds [N,6], root bound [6]) as device tensors (tests)."""
        n = self.num_surfels()
        b = torch.empty(n, 6, device=self.impl.device)
        r = torch.empty(6, device=self.impl.device)
        _lib.check(self.impl.lib.irgs_get_bounds(self.impl.h, _ptr(b), _ptr(r), _stream(self.impl.device)))
        return b, r

    # ------------------------------------------------------------------ tracing
    def _prep(self, rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs):
        dev = self.impl.device
        rays_o, rays_d = rays_o.contiguous(), rays_d.contiguous()
        means3D, opacity = means3D.contiguous(), opacity.contiguous()
        ru, rv, normals, shs = ru.contiguous(), rv.contiguous(), normals.contiguous(), shs.contiguous()
        features = features.contiguous() if features is not None else torch.zeros_like(means3D[:, :0])
        for name, t in (("rays_o", rays_o), ("rays_d", rays_d), ("means3D", means3D), ("opacity", opacity),
                        ("ru", ru), ("rv", rv), ("normals", normals), ("features", features), ("shs", shs)):
            _check_f32(name, t, dev)
        n = means3D.shape[0]
        if n != self.num_surfels():
            raise ValueError(f"trace got {n} surfels but the acceleration structure holds {self.num_surfels()}; "
                             "call build_bvh after changing the surfel set")
        if opacity.numel() != n or ru.shape[0] != n or rv.shape[0] != n or normals.shape[0] != n or \
                shs.shape[0] != n or features.shape[0] != n:
            raise ValueError("per-surfel arrays disagree on the number of surfels")
        if rays_o.shape != rays_d.shape or rays_o.shape[-1] != 3:
            raise ValueError("rays_o and rays_d must both be [..., 3]")
        return rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs

    def trace(self, rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs, alpha_min, deg=3,
              back_culling=False):
        rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs = self._prep(
            rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs)
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.view(-1, 3)
        rays_d = rays_d.view(-1, 3)
        S = features.shape[-1]
        if rays_o.shape[0] == 0:
            z = lambda *s: torch.zeros(*s, device=rays_o.device, dtype=torch.float32)  # noqa: E731
            return z(*prefix, 3), z(*prefix, 3), z(*prefix, S), z(*prefix), z(*prefix)
        color, normal, feature, depth, alpha, hit_count = _GaussianTrace.apply(
            self, rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs, float(alpha_min), int(deg),
            bool(back_culling))
        self.last_hit_count = hit_count.view(*prefix)
        return (color.view(*prefix, 3), normal.view(*prefix, 3), feature.view(*prefix, S), depth.view(*prefix),
                alpha.view(*prefix))

    def trace_incident(self, position, normals_pt, sample_num, means3D, opacity, ru, rv, normals, features, shs, alpha_min,
                       azimuth=None, t_min=0.05, deg=3, back_culling=False):
        """`sample_incident_rays` + `trace` of the reference's rendering_equation in one call (gaussian_renderer/
        __init__.py:324-332,376): the P x sample_num incident rays are generated inside the kernels (irgs_b200/incident.py).
        position, normals_pt [P,3]; azimuth [P] = the training mode's `rand * 2 pi` (None: evaluation mode).
        Returns (color [P,S,3], normal [P,S,3], feature [P,S,F], depth [P,S], alpha [P,S]); differentiable w.r.t.
        position, normals_pt and the surfel parameters."""
        from . import incident
        dummy = position.new_zeros(1, 3)
        _, _, means3D, opacity, ru, rv, normals, features, shs = self._prep(dummy, dummy, means3D, opacity, ru, rv, normals,
                                                                            features, shs)
        position, normals_pt = position.contiguous(), normals_pt.contiguous()
        azimuth = azimuth.contiguous().view(-1) if azimuth is not None else None
        incident._check(position, normals_pt, azimuth, self.impl.device)
        P, S, F = position.shape[0], int(sample_num), features.shape[-1]
        if P == 0:
            z = lambda *s: torch.zeros(*s, device=position.device, dtype=torch.float32)  # noqa: E731
            return z(0, S, 3), z(0, S, 3), z(0, S, F), z(0, S), z(0, S)
        color, normal, feature, depth, alpha, hit_count = incident._IncidentTrace.apply(
            self, position, normals_pt, azimuth, S, float(t_min), means3D, opacity, ru, rv, normals, features, shs,
            float(alpha_min), int(deg), bool(back_culling))
        self.last_hit_count = hit_count.view(P, S)
        return color.view(P, S, 3), normal.view(P, S, 3), feature.view(P, S, F), depth.view(P, S), alpha.view(P, S)

    @torch.no_grad()
    def trace_with_hits(self, rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs, alpha_min, deg=3,
                        back_culling=False, hit_cap=None):
        """Forward only, no autograd; additionally returns the ordered hit lists.  dict(color, normal, feature, depth,
        alpha, hit_count [R] int32, hits [R, hit_cap] int32 surfel ids in compositing order)."""
        rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs = self._prep(
            rays_o, rays_d, means3D, opacity, ru, rv, normals, features, shs)
        impl, dev = self.impl, self.impl.device
        rays_o, rays_d = rays_o.view(-1, 3), rays_d.view(-1, 3)
        B, S, K = rays_o.shape[0], features.shape[-1], shs.shape[1]
        cap = self.hit_cap if hit_cap is None else hit_cap
        out = dict(zip(("color", "normal", "feature", "depth", "alpha", "hit_count"), _alloc_outputs(B, S, dev)))
        out["hits"] = torch.full((B, cap), -1, device=dev, dtype=torch.int32)
        if B > 0:
            _lib.check(impl.lib.irgs_trace_forward(
                impl.h, B, S, K, int(deg), _ptr(rays_o), _ptr(rays_d), _ptr(means3D), _ptr(opacity), _ptr(ru), _ptr(rv),
                _ptr(normals), _ptr(features), _ptr(shs), _ptr(out["color"]), _ptr(out["normal"]), _ptr(out["feature"]),
                _ptr(out["depth"]), _ptr(out["alpha"]), _ptr(out["hit_count"]), _ptr(out["hits"]), cap, float(alpha_min),
                self.transmittance_min, int(back_culling), _stream(dev)))
        return out

    def run_chunks(self, n_items, per_chunk, body, streams=None):
        """Chunk loop of the reference's renderer (gaussian_renderer/__init__.py:319-322) with consecutive chunks
        alternating between two CUDA streams: `body(begin, end)` is called for every chunk inside the stream context of
        its turn (the native tracer keeps one work counter + candidate scratch per stream), so that the
        drain of one chunk's persistent kernels overlaps the next chunk (+5 % on the C3 step).  `body` may call
        trace / trace_incident and run their backward.  Returns after everything has joined the current stream again.
        Pass the same `streams` list on every call to reuse them."""
        dev = self.impl.device
        if streams is None:
            if self._side_streams is None:
                self._side_streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
            streams = self._side_streams
        if not 1 <= len(streams) <= 2:
            raise ValueError("one or two streams")
        cur = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cur)
        try:
            for i, b in enumerate(range(0, n_items, per_chunk)):
                k = i % len(streams)
                with torch.cuda.stream(streams[k]):
                    body(b, min(b + per_chunk, n_items))
        finally:
            for st in streams:
                cur.wait_stream(st)

    def set_option(self, name, value):
        """Tuning knobs of the native tracer (never change results), e.g. set_option("sort_rays_min", 0)."""
        _lib.check(self.impl.lib.irgs_set_option(self.impl.h, name.encode(), int(value)))

    def get_info(self, name):
        """Introspection of the native tracer: "tree_depth", "ploc_iterations", "n_slots", "n_surfels"."""
        return int(self.impl.lib.irgs_get_info(self.impl.h, name.encode()))

    def set_stats(self, enable):
        _lib.check(self.impl.lib.irgs_set_stats(self.impl.h, int(enable)))

    def get_stats(self):
        """(node visits, surfel tests, composited hits, passes) summed over the rays of the last forward."""
        arr = (ctypes.c_int64 * 4)()
        _lib.check(self.impl.lib.irgs_get_stats(self.impl.h, arr))
        return tuple(int(x) for x in arr)

    def intersection_test(self, rays_o, rays_d, means3D, opacity, ru, rv, normals, alpha_min):
        """bool mask [*]: ray crosses any surfel's alpha >= alpha_min support (raytracer.py:103-104)."""
        dev = self.impl.device
        ts = self._surfel_args(means3D, opacity, ru, rv, normals)
        o, d = rays_o.detach().contiguous().view(-1, 3), rays_d.detach().contiguous().view(-1, 3)
        out = torch.zeros(o.shape[0], dtype=torch.uint8, device=dev)
        _lib.check(self.impl.lib.irgs_intersection_test(self.impl.h, o.shape[0], _ptr(o), _ptr(d),
                                                        *[_ptr(t) for t in ts], alpha_min, _ptr(out), _stream(dev)))
        return out.bool().view(rays_o.shape[:-1])

    # ------------------------------------------------------------------ gradients
    def _unpack(self, fused, gfeat, opacity_shape, K):
        dev = self.impl.device
        n = fused.shape[0]
        gm = torch.empty(n, 3, device=dev); go = torch.empty(opacity_shape, device=dev)
        gru = torch.empty(n, 3, device=dev); grv = torch.empty(n, 3, device=dev); gn = torch.empty(n, 3, device=dev)
        gsh = torch.empty(n, K, 3, device=dev)
        _lib.check(self.impl.lib.irgs_unpack_grads(_ptr(fused), n, K, _ptr(gm), _ptr(go), _ptr(gru), _ptr(grv), _ptr(gn),
                                                   _ptr(gsh), _stream(dev)))
        return gm, go, gru, grv, gn, gfeat, gsh

    def _grad_buffers(self, n, S):
        dev = self.impl.device
        fresh = False
        if self._fused is None or self._fused.shape[0] != n:
            self._fused = torch.zeros(n, GRAD_STRIDE, device=dev, dtype=torch.float32)
            fresh = True
        if self._gfeat is None or self._gfeat.shape != (n, S):
            self._gfeat = torch.zeros(n, S, device=dev, dtype=torch.float32)
            fresh = True
        if fresh:   # (once) the zero fill must be complete before a backward on ANOTHER stream may add to the buffers
            torch.cuda.current_stream(dev).synchronize()
        return self._fused, self._gfeat

    def flush_grads(self, K=16, opacity_shape=None, all_reduce=True, group=None):
        """Deferred mode (accumulate_grads=True): returns the per-surfel gradients accumulated by all backward
        calls since the last flush as dict(means3D, opacity, ru, rv, normals, features, shs).  When
        torch.distributed is initialised and all_reduce is set, the fused [N,64] buffer is summed over ranks with ONE
        NCCL all-reduce first -- the only collective of the ray-sharded multi-GPU path."""
        if self._fused is None:
            raise RuntimeError("flush_grads: no gradients accumulated")
        fused, gfeat = self._fused, self._gfeat
        if all_reduce and torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(group) > 1:
            torch.distributed.all_reduce(fused, group=group)
            if gfeat.numel() > 0:
                torch.distributed.all_reduce(gfeat, group=group)
        n = fused.shape[0]
        out = self._unpack(fused, gfeat.clone(), opacity_shape or (n, 1), K)
        fused.zero_()
        gfeat.zero_()
        return dict(zip(("means3D", "opacity", "ru", "rv", "normals", "features", "shs"), out))
