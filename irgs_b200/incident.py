"""Fused incident-ray generation around the tracer.

The reference shades every surface point with S secondary rays that it first materialises as two [P, S, 3] tensors
(/root/reference/gaussian_renderer/__init__.py:324-332 `sample_incident_rays`, :376
`pc.trace(position.unsqueeze(1) + incident_dirs * pipe.light_t_min, incident_dirs, ...)`;
utils/graphics_utils.py:19-47 `fibonacci_sphere_sampling`, :133-165 `rotation_between_z`).  Here the tracing kernels
generate those rays themselves from one (position, normal, azimuth) triple per point, so neither the origins nor the
directions ever exist in memory on the tracing path, and the backward reduces the per-ray gradients to dL/dposition and
dL/dnormal per point.

    color, normal, feature, depth, alpha = tracer.trace_incident(position, normals, sample_num, azimuth, t_min, ...)

replaces `sample_incident_rays` + `pc.trace(...)`; `incident_dirs(...)` returns the same directions for the shading
code that follows (BRDF, environment lookup), computed by the same device function the tracer uses.
"""
import ctypes

import torch

from . import _lib
from .raytracer import GRAD_STRIDE, _alloc_outputs, _ptr, _reuse_records, _stream


class IncidentDesc(ctypes.Structure):
    """irgs_incident_t of include/irgs_b200.h."""
    _fields_ = [("position", ctypes.c_void_p), ("normals", ctypes.c_void_p), ("azimuth", ctypes.c_void_p),
                ("n_points", ctypes.c_int64), ("sample_num", ctypes.c_int32), ("t_min", ctypes.c_float)]


def _desc(position, normals, azimuth, sample_num, t_min):
    return IncidentDesc(position.data_ptr(), normals.data_ptr(), azimuth.data_ptr() if azimuth is not None else None,
                        position.shape[0], int(sample_num), float(t_min))


def rotation_between_z(vec):
    """utils/graphics_utils.py:133-165 restated in torch (differentiable); the tracer's own backward does this chain in its
    kernel, this copy is for shading code and for the tests."""
    v1, v2 = -vec[..., 1], vec[..., 0]
    c = (vec[..., 2] + 1).clamp_min(1e-7)
    zero = torch.zeros_like(v1)
    R = torch.stack([1 + (-v2 * v2) / c, v1 * v2 / c, v2,
                     v1 * v2 / c, 1 + (-v1 * v1) / c, -v1,
                     -v2, v1, 1 + (-v2 * v2 - v1 * v1) / c], -1).reshape(vec.shape[:-1] + (3, 3))
    R = R + 0 * zero[..., None, None]
    flip = -torch.eye(3, dtype=vec.dtype, device=vec.device).expand_as(R)
    return torch.where((vec[..., 2] + 1 > 0)[..., None, None], R, flip)


def _check(position, normals, azimuth, dev):
    for name, t in (("position", position), ("normals", normals)) + ((("azimuth", azimuth),) if azimuth is not None else ()):
        if t.dtype != torch.float32 or t.device != dev:
            raise TypeError(f"{name} must be a float32 tensor on {dev}")
    if position.shape != normals.shape or position.dim() != 2 or position.shape[-1] != 3:
        raise ValueError("position and normals must both be [P, 3]")
    if azimuth is not None and azimuth.numel() != position.shape[0]:
        raise ValueError("azimuth must have one entry per shading point")


@torch.no_grad()
def incident_rays(position, normals, sample_num, azimuth=None, t_min=0.05):
    """(rays_o, rays_d) [P, S, 3] exactly as the tracing kernels generate them."""
    dev = position.device
    position, normals = position.contiguous(), normals.contiguous()
    azimuth = azimuth.contiguous().view(-1) if azimuth is not None else None
    _check(position, normals, azimuth, dev)
    P = position.shape[0]
    o = torch.empty(P, sample_num, 3, device=dev)
    d = torch.empty(P, sample_num, 3, device=dev)
    if P > 0:
        desc = _desc(position, normals, azimuth, sample_num, t_min)
        _lib.check(_lib.load().irgs_incident_rays(ctypes.byref(desc), _ptr(o), _ptr(d), _stream(dev)))
    return o, d


def incident_dirs(normals, sample_num, azimuth=None):
    """The reference's `sample_incident_rays` directions [P, S, 3] (no gradient; the tracer's own backward carries
    dL/dnormal)."""
    return incident_rays(torch.zeros_like(normals), normals, sample_num, azimuth, 0.0)[1]


class _IncidentTrace(torch.autograd.Function):
    """_GaussianTrace (raytracer.py) with the rays generated in the kernels."""

    @staticmethod
    def forward(ctx, tracer, position, normals_pt, azimuth, sample_num, t_min, means3D, opacity, ru, rv, normals,
                features, shs, alpha_min, deg, back_culling):
        impl, dev = tracer.impl, tracer.impl.device
        P = position.shape[0]
        B, S, K = P * sample_num, features.shape[-1], shs.shape[1]
        color, normal, feature, depth, alpha, hit_count = _alloc_outputs(B, S, dev)
        cap = tracer.hit_cap if any(ctx.needs_input_grad) else 0
        hits = torch.empty(B, cap, device=dev, dtype=torch.int32) if cap > 0 else None
        desc = _desc(position, normals_pt, azimuth, sample_num, t_min)
        _lib.check(impl.lib.irgs_trace_forward_incident(
            impl.h, ctypes.byref(desc), S, K, deg, _ptr(means3D), _ptr(opacity), _ptr(ru), _ptr(rv), _ptr(normals),
            _ptr(features), _ptr(shs), _ptr(color), _ptr(normal), _ptr(feature), _ptr(depth), _ptr(alpha),
            _ptr(hit_count), _ptr(hits), cap, alpha_min, tracer.transmittance_min, int(back_culling), _stream(dev)))
        tracer.last_hit_count = hit_count
        ctx.tracer, ctx.cap = tracer, cap
        ctx.pack_epoch = impl.lib.irgs_get_info(impl.h, b"pack_epoch") if any(ctx.needs_input_grad) else -1
        ctx.cfg = (sample_num, t_min, alpha_min, deg, back_culling, tracer.transmittance_min, azimuth is not None)
        ctx.save_for_backward(position, normals_pt, azimuth if azimuth is not None else position[:0], means3D, opacity, ru,
                              rv, normals, features, shs, color, normal, feature, depth, alpha, hit_count,
                              hits if hits is not None else hit_count)
        ctx.mark_non_differentiable(hit_count)
        return color, normal, feature, depth, alpha, hit_count

    @staticmethod
    def backward(ctx, g_color, g_normal, g_feature, g_depth, g_alpha, _g_count):
        (position, normals_pt, azimuth, means3D, opacity, ru, rv, normals, features, shs, color, normal, feature, depth,
         alpha, hit_count, hits) = ctx.saved_tensors
        sample_num, t_min, alpha_min, deg, back_culling, T_min, has_azim = ctx.cfg
        tracer = ctx.tracer
        impl, dev = tracer.impl, tracer.impl.device
        P, N, S, K = position.shape[0], means3D.shape[0], features.shape[-1], shs.shape[1]
        B = P * sample_num
        g = [t.contiguous() for t in (g_color, g_normal, g_feature, g_depth, g_alpha)]
        scratch_o = torch.empty(B, 3, device=dev)
        scratch_d = torch.empty(B, 3, device=dev)
        g_pos = torch.empty(P, 3, device=dev)
        g_nrm = torch.empty(P, 3, device=dev)
        deferred = tracer.accumulate_grads
        if deferred:
            fused, gfeat = tracer._grad_buffers(N, S)
        else:
            fused = torch.zeros(N, GRAD_STRIDE, device=dev)
            gfeat = torch.zeros(N, S, device=dev)
        have_list = ctx.cap > 0
        desc = _desc(position, normals_pt, azimuth if has_azim else None, sample_num, t_min)
        null = ctypes.c_void_p(0)
        _reuse_records(ctx, impl)
        _lib.check(impl.lib.irgs_trace_backward_incident(
            impl.h, ctypes.byref(desc), S, K, deg, _ptr(means3D), _ptr(opacity), _ptr(ru), _ptr(rv), _ptr(normals),
            _ptr(features), _ptr(shs), _ptr(color), _ptr(normal), _ptr(feature), _ptr(depth), _ptr(alpha),
            _ptr(hit_count) if have_list else null, _ptr(hits) if have_list else null, ctx.cap, *[_ptr(t) for t in g],
            _ptr(scratch_o), _ptr(scratch_d), _ptr(g_pos), _ptr(g_nrm), _ptr(fused), _ptr(gfeat), alpha_min, T_min,
            int(back_culling), _stream(dev)))
        surf = (None,) * 7 if deferred else tracer._unpack(fused, gfeat, opacity.shape, K)
        return (None, g_pos, g_nrm, None, None, None) + surf + (None, None, None)
