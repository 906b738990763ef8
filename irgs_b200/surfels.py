"""Tracing straight from surfel parameters (means, scales, rotations, opacities): the input form BASELINE.json's
north_star names, and the caller glue of the reference restated once so that users need not copy it.

Mirrors scene/gaussian_model.py:712-756 of the reference:
  * `surfel_frames`     :733-747  ru = R[:,:,0] / s_u, rv = R[:,:,1] / s_v, normals = R[:,:,2] flipped towards the camera
  * `SurfelScene.build / refit` :725-731 (get_boundings + build_bvh / update_bvh) -- here the bounds are derived
    analytically from the parameters by the native library, no 12N-vertex proxy mesh is materialised
  * `SurfelScene.trace` :733-765 incl. the normalisation by alpha where alpha >= 1 - T_min (:751-756)
  * `SurfelScene.rendering_equation`: the reference's rendering_equation (gaussian_renderer/__init__.py:334-415) from the
    same parameters -- incident rays generated in the kernels, traced, shaded by the epilogue kernels (irgs_b200/shading.py)
`SurfelScene.trace` runs that glue as kernels (csrc/surfel_params.cu: irgs_surfel_frames, irgs_normalize_outputs and their
hand-derived backward, irgs_unpack_grads_params) inside ONE autograd Function, so a trace + backward from parameters launches
no element-wise torch kernels; `surfel_frames` is the same math in differentiable torch (tests, `rendering_equation`).
"""
import ctypes

import torch

from . import _lib
from .raytracer import GRAD_STRIDE, GaussianTracer, _alloc_outputs, _check_f32, _ptr, _stream


def _cam_ptr(camera_center):
    """HOST float[3] for the native calls (or NULL), kept alive by the caller."""
    if camera_center is None:
        return None, None
    cc = [float(v) for v in (camera_center.detach().cpu().tolist() if torch.is_tensor(camera_center) else camera_center)]
    arr = (ctypes.c_float * 3)(*cc)
    return arr, ctypes.cast(arr, ctypes.POINTER(ctypes.c_float))


class _SurfelTrace(torch.autograd.Function):
    """scene/gaussian_model.py:733-756 + raytracer.py:5-66 in one node: parameters -> frames (kernel) -> trace -> normalisation
    of saturated rays (kernel); backward: normalisation backward (kernel) -> trace backward -> parameter gradients (kernel)."""

    @staticmethod
    def forward(ctx, scene, rays_o, rays_d, means, scales, rotations, opacities, shs, features, camera_center, deg,
                back_culling, normalize):
        tracer = scene.tracer
        impl, dev, lib = tracer.impl, tracer.impl.device, tracer.impl.lib
        N, B, S, K = means.shape[0], rays_o.shape[0], features.shape[-1], shs.shape[1]
        cam_arr, cam = _cam_ptr(camera_center)
        ru, rv, normals = (torch.empty(N, 3, device=dev) for _ in range(3))
        st = _stream(dev)
        _lib.check(lib.irgs_surfel_frames(_ptr(means), _ptr(scales), _ptr(rotations), cam, N, _ptr(ru), _ptr(rv), _ptr(normals), st))
        *raw, hit_count = _alloc_outputs(B, S, dev)
        cap = tracer.hit_cap if any(ctx.needs_input_grad) else 0
        hits = torch.empty(B, cap, device=dev, dtype=torch.int32) if cap > 0 else None
        _lib.check(lib.irgs_trace_forward(
            impl.h, B, S, K, deg, _ptr(rays_o), _ptr(rays_d), _ptr(means), _ptr(opacities), _ptr(ru), _ptr(rv), _ptr(normals),
            _ptr(features), _ptr(shs), *[_ptr(t) for t in raw], _ptr(hit_count), _ptr(hits), cap, scene.alpha_min,
            tracer.transmittance_min, int(back_culling), st))
        tracer.last_hit_count = hit_count
        if normalize:
            outs = [torch.empty_like(t) for t in raw]
            _lib.check(lib.irgs_normalize_outputs(B, S, 1.0 - tracer.transmittance_min, *[_ptr(t) for t in raw],
                                                  *[_ptr(t) for t in outs], st))
        else:
            outs = raw
        ctx.scene, ctx.cap, ctx.cam = scene, cap, camera_center
        ctx.cfg = (deg, back_culling, normalize, scene.alpha_min, tracer.transmittance_min)
        ctx.save_for_backward(rays_o, rays_d, means, scales, rotations, opacities, shs, features, ru, rv, normals, *raw, hit_count,
                              hits if hits is not None else hit_count)
        ctx.mark_non_differentiable(hit_count, normals)
        return (*outs, hit_count, normals)

    @staticmethod
    def backward(ctx, g_color, g_normal, g_feature, g_depth, g_alpha, _g_count, _g_normals):
        (rays_o, rays_d, means, scales, rotations, opacities, shs, features, ru, rv, normals, color, normal, feature, depth, alpha,
         hit_count, hits) = ctx.saved_tensors
        deg, back_culling, normalize, alpha_min, T_min = ctx.cfg
        scene = ctx.scene
        tracer = scene.tracer
        impl, dev, lib = tracer.impl, tracer.impl.device, tracer.impl.lib
        N, B, S, K = means.shape[0], rays_o.shape[0], features.shape[-1], shs.shape[1]
        st = _stream(dev)
        g = [t.contiguous() for t in (g_color, g_normal, g_feature, g_depth, g_alpha)]
        if normalize:
            g = [t.clone() for t in g]      # rewritten in place as gradients of the raw accumulations
            _lib.check(lib.irgs_normalize_outputs_backward(B, S, 1.0 - T_min, _ptr(color), _ptr(normal), _ptr(feature), _ptr(depth),
                                                           _ptr(alpha), *[_ptr(t) for t in g], st))
        grad_rays_o, grad_rays_d = torch.empty_like(rays_o), torch.empty_like(rays_d)
        deferred = tracer.accumulate_grads
        if deferred:
            fused, gfeat = tracer._grad_buffers(N, S)
        else:
            fused = torch.zeros(N, GRAD_STRIDE, device=dev)
            gfeat = torch.zeros(N, S, device=dev)
        null = ctypes.c_void_p(0)
        have = ctx.cap > 0
        _lib.check(lib.irgs_trace_backward(
            impl.h, B, S, K, deg, _ptr(rays_o), _ptr(rays_d), _ptr(means), _ptr(opacities), _ptr(ru), _ptr(rv), _ptr(normals),
            _ptr(features), _ptr(shs), _ptr(color), _ptr(normal), _ptr(feature), _ptr(depth), _ptr(alpha),
            _ptr(hit_count) if have else null, _ptr(hits) if have else null, ctx.cap, *[_ptr(t) for t in g], _ptr(grad_rays_o),
            _ptr(grad_rays_d), _ptr(fused), _ptr(gfeat), alpha_min, T_min, int(back_culling), st))
        if deferred:
            surf = (None,) * 6       # collected by SurfelScene.flush_grads after the (single) all-reduce
        else:
            surf = scene._unpack_params(fused, gfeat, means, scales, rotations, opacities.shape, K, ctx.cam)
        gm, gs, gr, go, gsh, gf = surf
        return (None, grad_rays_o, grad_rays_d, gm, gs, gr, go, gsh, gf, None, None, None, None)


def quat_to_rot(q):
    """utils/general_utils.py:78-99: (w, x, y, z), not necessarily normalised -> R [N,3,3]."""
    q = q / q.norm(dim=-1, keepdim=True)
    r, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(*q.shape[:-1], 3, 3)


def surfel_frames(means, scales, rotations, camera_center=None):
    """(ru, rv, normals) as the tracer takes them (scene/gaussian_model.py:738-747)."""
    R = quat_to_rot(rotations)
    s = 1.0 / scales
    ru = R[:, :, 0] * s[:, 0:1]
    rv = R[:, :, 1] * s[:, 1:2]
    normals = R[:, :, 2]
    if camera_center is not None:
        cc = torch.as_tensor(camera_center, dtype=means.dtype, device=means.device)
        dotp = (normals * -(means - cc)).sum(-1, keepdim=True)          # utils/general_utils.py:135-146 flip_align_view
        normals = normals * torch.where(dotp >= 0, 1.0, -1.0)
    normals = normals / normals.norm(dim=-1, keepdim=True).clamp_min(1e-20)  # safe_normalize
    return ru, rv, normals


class SurfelScene:
    """A tracer bound to surfel parameters.  `alpha_min` / `transmittance_min` default to IRGS's hard-coded values
    (scene/gaussian_model.py:118-119)."""

    def __init__(self, transmittance_min=0.03, alpha_min=1.0 / 255.0, device=None):
        self.tracer = GaussianTracer(transmittance_min=transmittance_min, device=device)
        self.alpha_min = alpha_min
        self._built_n = None

    @torch.no_grad()
    def frames(self, means, scales, rotations, camera_center=None):
        """(ru, rv, normals) [N,3] from the parameters, one kernel (no gradient; `surfel_frames` is the differentiable twin)."""
        dev = self.tracer.impl.device
        means, scales, rotations = means.detach().contiguous(), scales.detach().contiguous(), rotations.detach().contiguous()
        n = means.shape[0]
        ru, rv, normals = (torch.empty(n, 3, device=dev) for _ in range(3))
        cam_arr, cam = _cam_ptr(camera_center)
        _lib.check(self.tracer.impl.lib.irgs_surfel_frames(_ptr(means), _ptr(scales), _ptr(rotations), cam, n, _ptr(ru), _ptr(rv),
                                                           _ptr(normals), _stream(dev)))
        return ru, rv, normals

    @torch.no_grad()
    def build(self, means, scales, rotations, opacities, camera_center=None):
        ru, rv, normals = self.frames(means, scales, rotations, camera_center)
        self.tracer.build_from_surfels(means, opacities, ru, rv, normals, self.alpha_min)
        self._built_n = means.shape[0]

    @torch.no_grad()
    def refit(self, means, scales, rotations, opacities, camera_center=None):
        """Per-iteration update with frozen topology (train.py:150-154 calls update_bvh when geometry moves)."""
        if self._built_n != means.shape[0]:
            return self.build(means, scales, rotations, opacities, camera_center)
        ru, rv, normals = self.frames(means, scales, rotations, camera_center)
        self.tracer.update_from_surfels(means, opacities, ru, rv, normals, self.alpha_min)

    def _unpack_params(self, fused, gfeat, means, scales, rotations, opacity_shape, K, camera_center):
        """Fused [N,64] rows -> (d/dmeans, d/dscales, d/drotations, d/dopacities, d/dshs, d/dfeatures), one kernel."""
        dev = self.tracer.impl.device
        n = fused.shape[0]
        gm, gs, gr = torch.empty(n, 3, device=dev), torch.empty(n, 2, device=dev), torch.empty(n, 4, device=dev)
        go, gsh = torch.empty(opacity_shape, device=dev), torch.empty(n, K, 3, device=dev)
        cam_arr, cam = _cam_ptr(camera_center)
        _lib.check(self.tracer.impl.lib.irgs_unpack_grads_params(
            _ptr(fused), n, K, _ptr(means), _ptr(scales), _ptr(rotations), cam, _ptr(gm), _ptr(go), _ptr(gs), _ptr(gr), _ptr(gsh),
            _stream(dev)))
        return gm, gs, gr, go, gsh, gfeat

    def flush_grads(self, means, scales, rotations, opacities, K=16, camera_center=None, all_reduce=True, group=None):
        """Deferred mode (`tracer.accumulate_grads = True`): the parameter gradients accumulated by all backward calls since the
        last flush, after ONE all-reduce of the fused buffer over the ranks: dict(means, scales, rotations, opacities, shs,
        features)."""
        tr = self.tracer
        if tr._fused is None:
            raise RuntimeError("flush_grads: no gradients accumulated")
        if all_reduce and torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size(group) > 1:
            torch.distributed.all_reduce(tr._fused, group=group)
            if tr._gfeat.numel() > 0:
                torch.distributed.all_reduce(tr._gfeat, group=group)
        out = self._unpack_params(tr._fused, tr._gfeat.clone(), means.detach().contiguous(), scales.detach().contiguous(),
                                  rotations.detach().contiguous(), tuple(opacities.shape), K, camera_center)
        tr._fused.zero_()
        tr._gfeat.zero_()
        return dict(zip(("means", "scales", "rotations", "opacities", "shs", "features"), out))

    def trace(self, rays_o, rays_d, means, scales, rotations, opacities, shs, features=None, camera_center=None,
              deg=3, back_culling=False, normalize=True):
        """dict(color, normal, feature, depth, alpha, hit_count, normals).  With `normalize` the accumulations of rays
        that saturated (alpha >= 1 - T_min) are divided by alpha and their alpha set to 1, as GaussianModel.trace does.
        Differentiable w.r.t. rays_o, rays_d, means, scales, rotations, opacities, shs, features."""
        dev = self.tracer.impl.device
        f = lambda t: t.contiguous()                                                        # noqa: E731
        rays_o, rays_d, means, scales, rotations, opacities, shs = map(f, (rays_o, rays_d, means, scales, rotations, opacities, shs))
        features = f(features) if features is not None else torch.zeros_like(means[:, :0])
        for name, t in (("rays_o", rays_o), ("rays_d", rays_d), ("means", means), ("scales", scales), ("rotations", rotations),
                        ("opacities", opacities), ("shs", shs), ("features", features)):
            _check_f32(name, t, dev)
        n = means.shape[0]
        if n != self.tracer.num_surfels():
            raise ValueError(f"trace got {n} surfels but the acceleration structure holds {self.tracer.num_surfels()}")
        if scales.shape != (n, 2) or rotations.shape != (n, 4) or opacities.numel() != n or shs.shape[0] != n or \
                features.shape[0] != n:
            raise ValueError("scales [N,2], rotations [N,4], opacities [N(,1)], shs [N,K,3], features [N,S] expected")
        if rays_o.shape != rays_d.shape or rays_o.shape[-1] != 3:
            raise ValueError("rays_o and rays_d must both be [..., 3]")
        prefix = rays_o.shape[:-1]
        S = features.shape[-1]
        if rays_o.numel() == 0:
            z = lambda *sh: torch.zeros(*sh, device=dev)                                     # noqa: E731
            _, _, normals = surfel_frames(means, scales, rotations, camera_center)
            return dict(color=z(*prefix, 3), normal=z(*prefix, 3), feature=z(*prefix, S), depth=z(*prefix), alpha=z(*prefix),
                        hit_count=torch.zeros(*prefix, device=dev, dtype=torch.int32), normals=normals)
        color, normal, feature, depth, alpha, hit_count, normals = _SurfelTrace.apply(
            self, rays_o.view(-1, 3), rays_d.view(-1, 3), means, scales, rotations, opacities, shs, features, camera_center,
            int(deg), bool(back_culling), bool(normalize))
        return dict(color=color.view(*prefix, 3), normal=normal.view(*prefix, 3), feature=feature.view(*prefix, S),
                    depth=depth.view(*prefix), alpha=alpha.view(*prefix), hit_count=hit_count.view(*prefix), normals=normals)

    def rendering_equation(self, base_color, roughness, normals_pt, position, viewdirs, means, scales, rotations, opacities,
                           shs, envmap, sample_num, training=False, azimuth=None, camera_center=None, deg=3,
                           light_t_min=0.05, wo_indirect=False, detach_indirect=False):
        """The reference's `rendering_equation(base_color, roughness, normals, position, viewdirs, pc, pipe, training)` with
        `pc` spelled out as its parameters: dict(diffuse, specular, light_direct) when training, plus visibility, light and
        light_indirect otherwise.  Differentiable w.r.t. the shading-point inputs, the environment texels and the surfel
        parameters (scales and rotations included)."""
        from . import shading
        ru, rv, normals = surfel_frames(means, scales, rotations, camera_center)
        return shading.rendering_equation(base_color, roughness, normals_pt, position, viewdirs, self.tracer,
                                          (means, opacities, ru, rv, normals, None, shs), envmap, sample_num,
                                          training=training, azimuth=azimuth, light_t_min=light_t_min,
                                          alpha_min=self.alpha_min, deg=deg, wo_indirect=wo_indirect,
                                          detach_indirect=detach_indirect)
