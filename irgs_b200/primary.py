"""Primary pass through the tracer: the G-buffer the reference takes from its 2DGS tile rasteriser
(/root/reference/gaussian_renderer/__init__.py:121-131 -> alpha, normal, depth, base colour + roughness, SH colour; consumed at
:134-183) produced by the SAME ray tracer that shoots the secondary rays, with the H x W pinhole rays of scene/cameras.py:87-100
generated inside the kernels (irgs_trace_forward_camera / irgs_trace_backward_camera): no [H*W,3] ray arrays, one launch for the
whole image, differentiable with respect to every surfel parameter.

    cam = Camera.from_reference(viewpoint_camera)            # or Camera(origin, cam_to_world, fx, fy, width, height)
    g = render_primary(tracer, cam, surfels, alpha_min)      # dict of [C,H,W] maps named like render_ir's results

Stated parity risk (not reproduced, by design): the rasteriser intersects splats per tile in screen space with a low-pass
filter and cuts Gaussians at 3 sigma, the tracer intersects the surfel planes exactly and cuts at alpha_min; the rasteriser's
median depth (allmap[5]) and distortion map (allmap[6]) have no counterpart here (`depth_ratio` = 0 is what this pass
implements: the expected depth).  What IS pinned: the generated rays against the reference's Camera arithmetic, and the traced
G-buffer against the CPU oracle on those rays (tests/test_gpu_primary.py).
"""
import ctypes
import math

import torch

from . import _lib
from .raytracer import GRAD_STRIDE, _alloc_outputs, _ptr, _reuse_records, _stream


class CameraDesc(ctypes.Structure):
    """irgs_camera_t of include/irgs_b200.h."""
    _fields_ = [("origin", ctypes.c_float * 3), ("cam_to_world", ctypes.c_float * 9), ("fx", ctypes.c_float),
                ("fy", ctypes.c_float), ("width", ctypes.c_int32), ("height", ctypes.c_int32)]


class Camera:
    """A pinhole camera as the tracer needs it: centre, camera-to-world rotation, focal lengths, image size."""

    def __init__(self, origin, cam_to_world, fx, fy, width, height):
        self.origin = [float(v) for v in torch.as_tensor(origin).reshape(-1).tolist()]
        self.cam_to_world = [float(v) for v in torch.as_tensor(cam_to_world, dtype=torch.float32).reshape(-1).tolist()]
        if len(self.origin) != 3 or len(self.cam_to_world) != 9:
            raise ValueError("origin must have 3 and cam_to_world 9 entries")
        self.fx, self.fy, self.width, self.height = float(fx), float(fy), int(width), int(height)

    @classmethod
    def from_reference(cls, cam):
        """From an IRGS `Camera` / `MiniCam` (scene/cameras.py): rays_d = rays_d_camera @ world_view_transform[:3,:3].T, focal
        lengths from FoVx / FoVy (:87-100)."""
        fx = cam.image_width / (2 * math.tan(cam.FoVx * 0.5))
        fy = cam.image_height / (2 * math.tan(cam.FoVy * 0.5))
        return cls(cam.camera_center, cam.world_view_transform[:3, :3], fx, fy, cam.image_width, cam.image_height)

    @classmethod
    def look_at(cls, eye, target, up, fov_x, width, height):
        """A synthetic camera (tests, benchmarks): OpenCV convention like the reference (x right, y down, z forward)."""
        eye, target, up = (torch.as_tensor(v, dtype=torch.float64) for v in (eye, target, up))
        z = torch.nn.functional.normalize(target - eye, dim=0)
        x = torch.nn.functional.normalize(torch.linalg.cross(z, up), dim=0)
        y = torch.linalg.cross(z, x)
        focal = width / (2 * math.tan(fov_x * 0.5))
        return cls(eye.float(), torch.stack([x, y, z], 1).float(), focal, focal, width, height)

    def desc(self):
        return CameraDesc((ctypes.c_float * 3)(*self.origin), (ctypes.c_float * 9)(*self.cam_to_world), self.fx, self.fy,
                          self.width, self.height)

    @torch.no_grad()
    def rays(self, device):
        """(rays_o, rays_d) [H*W,3] exactly as the tracing kernels generate them."""
        n = self.width * self.height
        o, d = torch.empty(n, 3, device=device), torch.empty(n, 3, device=device)
        d_ = self.desc()
        _lib.check(_lib.load().irgs_camera_rays(ctypes.byref(d_), _ptr(o), _ptr(d), _stream(torch.device(device))))
        return o, d


class _CameraTrace(torch.autograd.Function):
    """_GaussianTrace (raytracer.py) with the rays generated from a camera inside the kernels."""

    @staticmethod
    def forward(ctx, tracer, cam, means3D, opacity, ru, rv, normals, features, shs, alpha_min, deg, back_culling):
        impl, dev = tracer.impl, tracer.impl.device
        B, S, K = cam.width * cam.height, features.shape[-1], shs.shape[1]
        *outs, hit_count = _alloc_outputs(B, S, dev)
        cap = tracer.hit_cap if any(ctx.needs_input_grad) else 0
        hits = torch.empty(B, cap, device=dev, dtype=torch.int32) if cap > 0 else None
        desc = cam.desc()
        _lib.check(impl.lib.irgs_trace_forward_camera(
            impl.h, ctypes.byref(desc), S, K, deg, _ptr(means3D), _ptr(opacity), _ptr(ru), _ptr(rv), _ptr(normals), _ptr(features),
            _ptr(shs), *[_ptr(t) for t in outs], _ptr(hit_count), _ptr(hits), cap, alpha_min, tracer.transmittance_min,
            int(back_culling), _stream(dev)))
        tracer.last_hit_count = hit_count
        ctx.tracer, ctx.cam, ctx.cap = tracer, cam, cap
        ctx.pack_epoch = impl.lib.irgs_get_info(impl.h, b"pack_epoch") if any(ctx.needs_input_grad) else -1
        ctx.cfg = (alpha_min, deg, back_culling, tracer.transmittance_min)
        ctx.save_for_backward(means3D, opacity, ru, rv, normals, features, shs, *outs, hit_count, hits if hits is not None else hit_count)
        ctx.mark_non_differentiable(hit_count)
        return (*outs, hit_count)

    @staticmethod
    def backward(ctx, *g):
        means3D, opacity, ru, rv, normals, features, shs, color, normal, feature, depth, alpha, hit_count, hits = ctx.saved_tensors
        alpha_min, deg, back_culling, T_min = ctx.cfg
        tracer, cam = ctx.tracer, ctx.cam
        impl, dev = tracer.impl, tracer.impl.device
        B, N, S, K = cam.width * cam.height, means3D.shape[0], features.shape[-1], shs.shape[1]
        gs = [t.contiguous() for t in g[:5]]
        scratch_o, scratch_d = torch.empty(B, 3, device=dev), torch.empty(B, 3, device=dev)
        deferred = tracer.accumulate_grads
        if deferred:
            fused, gfeat = tracer._grad_buffers(N, S)
        else:
            fused, gfeat = torch.zeros(N, GRAD_STRIDE, device=dev), torch.zeros(N, S, device=dev)
        have, null = ctx.cap > 0, ctypes.c_void_p(0)
        desc = cam.desc()
        _reuse_records(ctx, impl)
        _lib.check(impl.lib.irgs_trace_backward_camera(
            impl.h, ctypes.byref(desc), S, K, deg, _ptr(means3D), _ptr(opacity), _ptr(ru), _ptr(rv), _ptr(normals), _ptr(features),
            _ptr(shs), _ptr(color), _ptr(normal), _ptr(feature), _ptr(depth), _ptr(alpha), _ptr(hit_count) if have else null,
            _ptr(hits) if have else null, ctx.cap, *[_ptr(t) for t in gs], _ptr(scratch_o), _ptr(scratch_d), _ptr(fused), _ptr(gfeat),
            alpha_min, T_min, int(back_culling), _stream(dev)))
        surf = (None,) * 7 if deferred else tracer._unpack(fused, gfeat, opacity.shape, K)
        return (None, None) + surf + (None, None, None)


def trace_camera(tracer, cam, means3D, opacity, ru, rv, normals, features, shs, alpha_min, deg=3, back_culling=False):
    """`tracer.trace(cam.rays ...)` without the ray arrays: (color [H,W,3], normal [H,W,3], feature [H,W,S], depth [H,W],
    alpha [H,W]) raw accumulations, differentiable w.r.t. the surfel arrays."""
    dummy = means3D.new_zeros(1, 3)
    _, _, means3D, opacity, ru, rv, normals, features, shs = tracer._prep(dummy, dummy, means3D, opacity, ru, rv, normals, features, shs)
    H, W, S = cam.height, cam.width, features.shape[-1]
    color, normal, feature, depth, alpha, hit_count = _CameraTrace.apply(tracer, cam, means3D, opacity, ru, rv, normals, features, shs,
                                                                         float(alpha_min), int(deg), bool(back_culling))
    tracer.last_hit_count = hit_count.view(H, W)
    return color.view(H, W, 3), normal.view(H, W, 3), feature.view(H, W, S), depth.view(H, W), alpha.view(H, W)


def render_primary(tracer, cam, surfels, alpha_min, deg=3, bg_color=None):
    """The primary G-buffer of render_ir (gaussian_renderer/__init__.py:121-183) by ray tracing.  `surfels` = (means3D, opacity,
    ru, rv, normals, features [N,4] = cat([base_color, roughness]), shs).  Returns [C,H,W] maps under render_ir's names:
      rend_alpha [1,H,W]; rend_normal [3,H,W] (world space, alpha-weighted, facing the camera); surf_depth [1,H,W] = the expected
      z-depth (allmap[0] / alpha, nan -> 0: `depth_ratio` 0); normal_map [H,W,3] = normalize(rend_normal / alpha); base_color
      [3,H,W], roughness [1,H,W] (alpha-weighted accumulations, like rendered_features); render [3,H,W] = SH colour
      (+ (1 - alpha) * bg_color if given); points [H,W,3] = surf_depth * rays_d_hw_unnormalized + camera_center (:153)."""
    means3D, opacity, ru, rv, normals, features, shs = surfels
    if features is None or features.shape[-1] != 4:
        raise ValueError("render_primary traces cat([base_color, roughness]) as features: surfels[5] must be [N,4]")
    color, normal, feature, depth, alpha = trace_camera(tracer, cam, means3D, opacity, ru, rv, normals, features, shs, alpha_min, deg)
    dev = color.device
    _, d = cam.rays(dev)
    m = torch.tensor(cam.cam_to_world, device=dev).view(3, 3)
    z_of_ray = (d.view(cam.height, cam.width, 3) @ m[:, 2])          # cos between the ray and the optical axis: z-depth = t * cos
    a_ = alpha.clamp_min(1e-6)
    depth_expected = torch.nan_to_num(depth * z_of_ray / alpha, 0.0, 0.0)
    normal_map = torch.nn.functional.normalize(normal / a_[..., None], dim=-1)
    d_unnorm = d.view(cam.height, cam.width, 3) / z_of_ray[..., None]
    origin = torch.tensor(cam.origin, device=dev)
    render = color if bg_color is None else color + (1 - alpha[..., None]) * torch.as_tensor(bg_color, device=dev, dtype=color.dtype)
    return {
        "render": render.permute(2, 0, 1), "rend_alpha": alpha[None], "rend_normal": normal.permute(2, 0, 1),
        "surf_depth": depth_expected[None], "normal_map": normal_map, "base_color": feature[..., :3].permute(2, 0, 1),
        "roughness": feature[..., 3:4].permute(2, 0, 1), "points": depth_expected[..., None] * d_unnorm + origin,
        "rays_d_hw": d.view(cam.height, cam.width, 3),
    }
