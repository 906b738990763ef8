"""ctypes stand-in for the reference's pybind11 extension `surfel_tracer._C` (src/bindings.cu:24-116), bound to the C ABI of
libirgs_b200.so (include/irgs_b200.h).

With this module in place the reference's own `surfel_tracer/raytracer.py` (121 lines) runs UNCHANGED on the B200-native
tracer: it does `from surfel_tracer import _C`, `_C.create_gaussiantracer()` and calls the five methods below with the
argument lists of bindings.cu.  tests/test_gpu_dropin.py executes the unmodified file from baseline/_ref against this stub.
(The package's own `GaussianTracer`, irgs_b200/raytracer.py, is the faster way in: no mask / compaction pre-pass, saved
hit lists instead of a re-trace.)
"""
import ctypes

import torch

from irgs_b200 import _lib as _binding

_vp = ctypes.c_void_p


def _p(t):
    return _vp(t.data_ptr() if t is not None and t.numel() else 0)


def _s(t):
    return _vp(torch.cuda.current_stream(t.device).cuda_stream)


class GaussianTracer:
    """bindings.cu:24-98.  Tensors are CUDA float32 and contiguous, as the reference assumes without checking."""

    def __init__(self):
        self._lib = _binding.load()
        self.h = _vp()
        _binding.check(self._lib.irgs_tracer_create(ctypes.byref(self.h), torch.cuda.current_device()))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self._lib.irgs_tracer_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @staticmethod
    def _soup(triangles):
        # triangles [n_tri, 3, 3]: the gathered soup `vertices_b[faces_b]` of raytracer.py:77, 20 triangles = 60 vertices
        # per surfel (scene/gaussian_model.py:712-723); the bound of a surfel is the AABB of its 60 soup vertices
        v = triangles.reshape(-1, 3).contiguous().float()
        if v.shape[0] % 60:
            raise RuntimeError("build_bvh expects the 20-triangles-per-surfel proxy soup of IRGS")
        return v, v.shape[0] // 60

    def build_bvh(self, triangles):                      # bindings.cu:30-34
        v, n = self._soup(triangles)
        _binding.check(self._lib.irgs_build_from_proxy(self.h, _p(v), n, 60, _s(v)))

    def update_bvh(self, triangles):                     # bindings.cu:36-40
        v, n = self._soup(triangles)
        _binding.check(self._lib.irgs_refit_from_proxy(self.h, _p(v), n, 60, _s(v)))

    def intersection_test(self, rays_o, rays_d, gs_idxs, means3D, opacity, ru, rv, normals, intersection):   # bindings.cu:61-71
        # the reference tests the proxy triangles and has no alpha_min here; alpha_min = 0 makes the native test accept every
        # plane crossing inside a surfel's bound: a superset of the rays that can composite anything, like the proxy test
        out = torch.empty(rays_o.shape[0], dtype=torch.uint8, device=rays_o.device)
        _binding.check(self._lib.irgs_intersection_test(self.h, rays_o.shape[0], _p(rays_o), _p(rays_d), _p(means3D),
                                                        _p(opacity), _p(ru), _p(rv), _p(normals), 0.0, _p(out), _s(rays_o)))
        intersection.copy_(out.bool())

    def trace_forward(self, rays_o, rays_d, gs_idxs, means3D, opacity, ru, rv, normals, features, shs, color, normal,
                      feature, depth, alpha, alpha_min, transmittance_min, deg, back_culling):   # bindings.cu:42-59
        _binding.check(self._lib.irgs_trace_forward(
            self.h, rays_o.shape[0], features.shape[1], shs.shape[1], int(deg), _p(rays_o), _p(rays_d), _p(means3D),
            _p(opacity), _p(ru), _p(rv), _p(normals), _p(features), _p(shs), _p(color), _p(normal), _p(feature), _p(depth),
            _p(alpha), _vp(0), _vp(0), 0, float(alpha_min), float(transmittance_min), int(bool(back_culling)), _s(rays_o)))

    def trace_backward(self, rays_o, rays_d, gs_idxs, means3D, opacity, ru, rv, normals, features, shs, color, normal,
                       feature, depth, alpha, grad_rays_o, grad_rays_d, grad_means3D, grad_opacity, grad_ru, grad_rv,
                       grad_normals, grad_features, grad_shs, grad_out_color, grad_out_normal, grad_out_feature,
                       grad_out_depth, grad_out_alpha, alpha_min, transmittance_min, deg, back_culling):   # bindings.cu:73-94
        n, K = means3D.shape[0], shs.shape[1]
        fused = torch.zeros(n, 64, device=means3D.device, dtype=torch.float32)
        gouts = [g.contiguous() for g in (grad_out_color, grad_out_normal, grad_out_feature, grad_out_depth, grad_out_alpha)]
        # no saved hit list on this path (the reference's autograd Function saves none): re-trace mode, like the reference
        _binding.check(self._lib.irgs_trace_backward(
            self.h, rays_o.shape[0], features.shape[1], K, int(deg), _p(rays_o), _p(rays_d), _p(means3D), _p(opacity),
            _p(ru), _p(rv), _p(normals), _p(features), _p(shs), _p(color), _p(normal), _p(feature), _p(depth), _p(alpha),
            _vp(0), _vp(0), 0, *[_p(g) for g in gouts], _p(grad_rays_o), _p(grad_rays_d), _p(fused), _p(grad_features),
            float(alpha_min), float(transmittance_min), int(bool(back_culling)), _s(rays_o)))
        _binding.check(self._lib.irgs_unpack_grads(_p(fused), n, K, _p(grad_means3D), _p(grad_opacity), _p(grad_ru),
                                                   _p(grad_rv), _p(grad_normals), _p(grad_shs), _s(rays_o)))


def create_gaussiantracer():                             # bindings.cu:101-103
    return GaussianTracer()
