"""Shading epilogue around the incident-ray trace (SURVEY.md 8f rank 1 + rank 3): the reference's `rendering_equation`
as two CUDA kernels (forward / backward) instead of some thirty element-wise torch kernels over [P, S, 3] tensors.

Reference being mirrored (paths under /root/reference):
  * gaussian_renderer/__init__.py:334-415  rendering_equation(base_color, roughness, normals, position, viewdirs, pc, pipe,
        training=...) -> {"diffuse", "specular", "light_direct"} (training) + {"visibility", "light", "light_indirect"}
        (evaluation); the non-relight path with `diffuse_sample_num > 0`: pure Fibonacci sampling (`light_sample_num == 0`,
        the stage-2 training path and BASELINE config C3) and the mix with light-importance samples (`light_sample_num > 0`,
        :340-357, what render.py uses after `env_map.update_pdf()`).  The relight path (nvdiffrast cube-map mips + FG LUT)
        is not built and raises NotImplementedError.
  * gaussian_renderer/__init__.py:417-457  GGX_specular
  * scene/light.py:132-172,245-246,287-297,315  EnvLight (base, activation, transform, __call__(mode='pure_env'));
        :174-223 update_pdf, sample_light_directions, light_pdf
  * scene/gaussian_model.py:748-752  GaussianModel.trace's normalisation of saturated rays (folded into the kernels)

    env = EnvLight(resolution=(256, 512), activation="exp")                   # or any object with .base / .activation_name / .transform
    out = rendering_equation(base_color, roughness, normals, position, viewdirs, tracer, surfels, env,
                             sample_num=256, training=True)
    out["diffuse"], out["specular"], out["light_direct"]                      # [P,3] each, differentiable

`surfels` is the tuple (means3D, opacity, ru, rv, normals, features, shs) that `GaussianTracer.trace` takes.  Gradients
reach base_color, roughness, normals (n_d_i, GGX and the sampled directions), position, viewdirs, the environment texels
and -- through the tracer's backward -- the surfel parameters.  No CPU path: everything runs in libirgs_b200.so.
"""
import ctypes
import math

import torch

from . import _lib
from .incident import _check, _desc
from .raytracer import _ptr, _stream

ACTIVATIONS = {"none": 0, "exp": 1, "sigmoid": 2}
OUT_SLICES = {"diffuse": slice(0, 3), "specular": slice(3, 6), "light_direct": slice(6, 9), "visibility": slice(9, 10),
              "light": slice(10, 13), "light_indirect": slice(13, 16)}


class EnvDesc(ctypes.Structure):
    """irgs_envmap_t of include/irgs_b200.h."""
    _fields_ = [("base", ctypes.c_void_p), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
                ("activation", ctypes.c_int32), ("has_transform", ctypes.c_int32), ("transform", ctypes.c_float * 9)]


def _env_desc(base, activation, transform):
    if base.dim() != 3 or base.shape[-1] != 3 or base.dtype != torch.float32 or not base.is_cuda:
        raise ValueError("environment base must be a float32 CUDA tensor [H, W, 3]")
    if activation not in ACTIVATIONS:
        raise NotImplementedError(f"environment activation {activation!r} (light.py:160-168 knows exp / sigmoid / none)")
    d = EnvDesc(base.data_ptr(), base.shape[0], base.shape[1], ACTIVATIONS[activation], int(transform is not None))
    if transform is not None:
        t = [float(x) for x in torch.as_tensor(transform, dtype=torch.float32).reshape(-1).tolist()]
        if len(t) != 9:
            raise ValueError("environment transform must be 3 x 3")
        for j in range(9):
            d.transform[j] = t[j]
    return d


class SamplingDesc(ctypes.Structure):
    """irgs_shade_sampling_t of include/irgs_b200.h."""
    _fields_ = [("dirs", ctypes.c_void_p), ("pdf", ctypes.c_void_p), ("p_diffuse", ctypes.c_float), ("p_light", ctypes.c_float),
                ("total_samples", ctypes.c_int32)]


def _sampling_desc(dirs, pdf, p_diffuse, p_light, total):
    if dirs is None and pdf is None:
        return None
    return SamplingDesc(dirs.data_ptr() if dirs is not None else None, pdf.data_ptr() if pdf is not None else None,
                        float(p_diffuse), float(p_light), int(total))


def _env_fields(envmap):
    """(base, activation name, transform) of a reference-style EnvLight object (duck-typed: scene/light.py:132-172)."""
    return envmap.base, getattr(envmap, "activation_name", "exp"), getattr(envmap, "transform", None)


class _EnvLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dirs, base, activation, transform):
        dirs_c, base_c = dirs.reshape(-1, 3).contiguous(), base.contiguous()
        out = torch.empty_like(dirs_c)
        desc = _env_desc(base_c, activation, transform)
        _lib.check(_lib.load().irgs_env_lookup_forward(ctypes.byref(desc), _ptr(dirs_c), dirs_c.shape[0], _ptr(out),
                                                       _stream(dirs_c.device)))
        ctx.save_for_backward(dirs_c, base_c)
        ctx.cfg = (activation, transform, dirs.shape)
        return out.view(dirs.shape)

    @staticmethod
    def backward(ctx, g):
        dirs_c, base_c = ctx.saved_tensors
        activation, transform, shape = ctx.cfg
        g = g.reshape(-1, 3).contiguous()
        g_dirs = torch.empty_like(dirs_c) if ctx.needs_input_grad[0] else None
        g_env = torch.zeros_like(base_c) if ctx.needs_input_grad[1] else None
        desc = _env_desc(base_c, activation, transform)
        _lib.check(_lib.load().irgs_env_lookup_backward(ctypes.byref(desc), _ptr(dirs_c), _ptr(g), dirs_c.shape[0],
                                                        _ptr(g_dirs), _ptr(g_env), _stream(dirs_c.device)))
        return (g_dirs.view(shape) if g_dirs is not None else None), g_env, None, None


class EnvLight(torch.nn.Module):
    """The part of the reference's EnvLight (scene/light.py:132-172,245-246,287-297,315) that the tracing path uses: a
    learnable lat-long map `base` [H, W, 3] stored pre-activation, `activation_name`, an optional `transform`, and
    `__call__(dirs, mode='pure_env')`.  Cube-map mips (modes 'diffuse' / 'specular', nvdiffrast) are not built."""

    def __init__(self, resolution=(256, 512), activation="exp", init_value=0.5, device="cuda"):
        super().__init__()
        if activation not in ACTIVATIONS:
            raise NotImplementedError(activation)
        base = torch.full((resolution[0], resolution[1], 3), float(init_value), dtype=torch.float32, device=device)
        if activation == "sigmoid":
            base = torch.log(base / (1 - base))
        elif activation == "exp":
            base = torch.log(base)
        self.base = torch.nn.Parameter(base, requires_grad=True)
        self.activation_name = activation
        self.transform = None

    def set_transform(self, transform):
        self.transform = transform

    @torch.no_grad()
    def update_pdf(self):
        """light.py:174-179: texel sampling probabilities `_pdf` [H,W] from the current map (max over channels x sin(theta))."""
        H, W = self.base.shape[:2]
        y = ((torch.arange(H, dtype=torch.float32, device=self.base.device) + 0.5) / H)[:, None]
        act = {"exp": torch.exp, "sigmoid": torch.sigmoid, "none": lambda x: x}[self.activation_name]
        pdf = act(self.base).clamp_min(0.0).max(dim=-1)[0] * torch.sin(y * math.pi)
        self._pdf = (pdf / pdf.sum()).contiguous()

    @torch.no_grad()
    def sample_light_directions(self, B, sample_num, training=False):
        """light.py:181-205: B x sample_num directions drawn texel-wise from `_pdf` (jittered inside the texel when
        training) and their solid-angle densities -> ([B, sample_num, 3], [B, sample_num, 1])."""
        H, W = self._pdf.shape
        idx = torch.multinomial(self._pdf.reshape(-1), B * sample_num, replacement=True)
        gx = ((idx % W + 0.5) / W) * 2 - 1
        gy = (idx // W + 0.5) / H
        if training:
            gx = gx + (torch.rand_like(gx) - 0.5) / W * 2
            gy = gy + (torch.rand_like(gy) - 0.5) / H
        st, ct = torch.sin(gy * math.pi), torch.cos(gy * math.pi)
        sp, cp = torch.sin(gx * math.pi), torch.cos(gx * math.pi)
        direction = torch.stack((st * sp, ct, -st * cp), dim=-1)
        if self.transform is not None:
            direction = direction @ torch.as_tensor(self.transform, dtype=torch.float32, device=direction.device)
        direction = direction.reshape(B, sample_num, 3).contiguous()
        return direction, self.light_pdf(direction)

    @torch.no_grad()
    def light_pdf(self, direction):
        """light.py:207-223 (the shading kernels evaluate the same expression per sample; this copy serves callers that
        want the densities themselves)."""
        H, W = self._pdf.shape
        flat = direction.reshape(-1, 3)
        if self.transform is not None:
            flat = flat @ torch.as_tensor(self.transform, dtype=torch.float32, device=flat.device).T
        u = torch.atan2(flat[..., 0], -flat[..., 2]).nan_to_num() / (2.0 * math.pi) + 0.5
        v = torch.acos(flat[..., 1].clamp(-1.0 + 1e-6, 1.0 - 1e-6)) / math.pi
        idx = (u * W).clamp(0, W - 1).long() + (v * H).clamp(0, H - 1).long() * W
        weight = H * W / (2.0 * math.pi ** 2 * torch.sin(v * math.pi).clamp_min(1e-6))
        return (self._pdf.reshape(-1)[idx] * weight).reshape(*direction.shape[:-1], 1)

    def __call__(self, l, mode="pure_env", roughness=None):
        if mode != "pure_env":
            raise NotImplementedError("only mode='pure_env' is built (the cube-map modes need nvdiffrast's mip chain)")
        return _EnvLookup.apply(l, self.base, self.activation_name, self.transform)


def env_lookup(dirs, base, activation="exp", transform=None):
    """EnvLight.__call__(dirs, mode='pure_env') as a function: dirs [...,3] -> radiance [...,3] (differentiable w.r.t.
    dirs and base)."""
    return _EnvLookup.apply(dirs, base, activation, transform)


class _ShadeIncident(torch.autograd.Function):
    """out [P,16] = means over the S incident samples (layout: OUT_SLICES)."""

    @staticmethod
    def forward(ctx, normals_pt, azimuth, sample_num, base_color, roughness, viewdirs, env_base, activation, transform,
                trace_color, trace_alpha, saturate_alpha, dirs, pdf, p_diffuse, p_light, total_samples):
        dev = normals_pt.device
        P = normals_pt.shape[0]
        out = torch.empty(P, 16, device=dev)
        gen = _desc(normals_pt, normals_pt, azimuth, sample_num, 0.0)
        env = _env_desc(env_base, activation, transform)
        smp = _sampling_desc(dirs, pdf, p_diffuse, p_light, total_samples)
        _lib.check(_lib.load().irgs_shade_forward(ctypes.byref(gen), ctypes.byref(env), ctypes.byref(smp) if smp else None,
                                                  _ptr(base_color), _ptr(roughness), _ptr(viewdirs), _ptr(trace_color),
                                                  _ptr(trace_alpha), saturate_alpha, _ptr(out), _stream(dev)))
        none = normals_pt[:0]
        ctx.save_for_backward(normals_pt, azimuth if azimuth is not None else none, base_color, roughness, viewdirs,
                              env_base, trace_color, trace_alpha, dirs if dirs is not None else none,
                              pdf if pdf is not None else none)
        ctx.cfg = (sample_num, activation, transform, saturate_alpha, azimuth is not None, dirs is not None, pdf is not None,
                   p_diffuse, p_light, total_samples)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (normals_pt, azimuth, base_color, roughness, viewdirs, env_base, trace_color, trace_alpha, dirs,
         pdf) = ctx.saved_tensors
        (sample_num, activation, transform, saturate_alpha, has_azim, has_dirs, has_pdf, p_diffuse, p_light,
         total_samples) = ctx.cfg
        dev = normals_pt.device
        P = normals_pt.shape[0]
        g_out = g_out.contiguous()
        g_color = torch.empty_like(trace_color)
        g_alpha = torch.empty_like(trace_alpha)
        g_point = torch.empty(P, 16, device=dev)
        g_env = torch.zeros_like(env_base) if ctx.needs_input_grad[6] else None
        gen = _desc(normals_pt, normals_pt, azimuth if has_azim else None, sample_num, 0.0)
        env = _env_desc(env_base, activation, transform)
        smp = _sampling_desc(dirs if has_dirs else None, pdf if has_pdf else None, p_diffuse, p_light, total_samples)
        _lib.check(_lib.load().irgs_shade_backward(ctypes.byref(gen), ctypes.byref(env), ctypes.byref(smp) if smp else None,
                                                   _ptr(base_color), _ptr(roughness), _ptr(viewdirs), _ptr(trace_color),
                                                   _ptr(trace_alpha), saturate_alpha, _ptr(g_out), _ptr(g_color),
                                                   _ptr(g_alpha), _ptr(g_point), _ptr(g_env), _stream(dev)))
        return (g_point[:, 4:7], None, None, g_point[:, 0:3], g_point[:, 3], g_point[:, 7:10], g_env, None, None,
                g_color, g_alpha, None, None, None, None, None, None)


def shade_incident(normals, sample_num, base_color, roughness, viewdirs, env_base, trace_color, trace_alpha, azimuth=None,
                   activation="exp", transform=None, transmittance_min=None, dirs=None, pdf=None, p_diffuse=1.0, p_light=0.0,
                   total_samples=None):
    """The rendering-equation epilogue on its own: given the tracer's RAW colour [P,S,3] / alpha [P,S] of the incident rays
    that `GaussianTracer.trace_incident(position, normals, sample_num, ..., azimuth=azimuth)` traced, returns the dict
    of rendering_equation (all six keys; [P,3] each, visibility [P,1]).  transmittance_min: apply GaussianModel.trace's
    normalisation of saturated rays (None: the colour / alpha are used as they are).
    Mixed sampling (light_sample_num > 0): `dirs` [P,S,3] = explicit directions of these samples (the light samples; None:
    the generated Fibonacci ones), `pdf` [H,W] = the map's texel probabilities (EnvLight._pdf), p_diffuse / p_light = the
    two sample fractions, total_samples = diffuse + light samples (the divisor of the means): the results of the two
    calls, one per kind of sample, add up to rendering_equation's."""
    dev = normals.device
    f = lambda t: t.contiguous()                                                        # noqa: E731
    normals, base_color, viewdirs = f(normals), f(base_color), f(viewdirs)
    roughness = f(roughness).view(-1)
    azimuth = f(azimuth).view(-1) if azimuth is not None else None
    _check(normals, normals, azimuth, dev)
    P, S = normals.shape[0], int(sample_num)
    for name, t, shape in (("base_color", base_color, (P, 3)), ("roughness", roughness, (P,)), ("viewdirs", viewdirs, (P, 3)),
                           ("trace_color", trace_color, (P, S, 3)), ("trace_alpha", trace_alpha, (P, S))):
        if t.dtype != torch.float32 or t.device != dev:
            raise TypeError(f"{name} must be a float32 tensor on {dev}")
        if tuple(t.shape) != shape:
            raise ValueError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
    sat = -1.0 if transmittance_min is None else 1.0 - float(transmittance_min)
    if dirs is not None:
        dirs = f(dirs.detach())
        if dirs.dtype != torch.float32 or dirs.device != dev or tuple(dirs.shape) != (P, S, 3):
            raise ValueError(f"dirs must be a float32 tensor of shape {(P, S, 3)} on {dev}")
    if pdf is not None:
        pdf = f(pdf.detach())
        if pdf.dtype != torch.float32 or pdf.device != dev or tuple(pdf.shape) != tuple(env_base.shape[:2]):
            raise ValueError("pdf must be a float32 tensor with the environment map's [H, W]")
    total = S if total_samples is None else int(total_samples)
    if P == 0:
        out = torch.zeros(0, 16, device=dev)
    else:
        out = _ShadeIncident.apply(normals, azimuth, S, base_color, roughness, viewdirs, f(env_base), activation, transform,
                                   f(trace_color).view(P * S, 3), f(trace_alpha).view(P * S), sat, dirs, pdf, float(p_diffuse),
                                   float(p_light), total)
    return {k: out[:, s] for k, s in OUT_SLICES.items()}


@torch.no_grad()
def relight_local_lights(dirs, trace_normal, trace_feature, trace_alpha, envmap, fg_lut, f0=0.04, transmittance_min=None,
                         wo_indirect_relight=False):
    """gaussian_renderer/__init__.py:365-379: the light a secondary ray brings back under NOVEL lighting -- its hit point shaded
    with the environment's diffuse / specular prefilter and the FG table -- from the tracer's RAW normal [...,3], feature
    [...,4] = (base colour, roughness) and alpha [...] of rays with directions dirs [...,3].  `envmap(l, mode='diffuse')` and
    `envmap(l, roughness=r, mode='specular')` are the caller's (scene/light.py:264-328: cube-map mips, out of scope here); the
    arithmetic around them runs in two kernels (irgs_relight_hit / irgs_relight_combine).  fg_lut: pc.FG_LUT, [1,H,W,2] or
    [H,W,2].  Returns (local_incident_lights [...,3], alpha [...] after GaussianModel.trace's normalisation)."""
    dev = dirs.device
    shape = dirs.shape[:-1]
    f = lambda t, c: t.reshape(-1, c).contiguous() if c else t.reshape(-1).contiguous()      # noqa: E731
    d, n, ft, a = f(dirs, 3), f(trace_normal, 3), f(trace_feature, 4), f(trace_alpha, 0)
    for name, t in (("dirs", d), ("trace_normal", n), ("trace_feature", ft), ("trace_alpha", a)):
        if t.dtype != torch.float32 or t.device != dev or not t.is_cuda:
            raise TypeError(f"{name} must be a float32 CUDA tensor on {dev}")
    R = d.shape[0]
    if n.shape[0] != R or ft.shape[0] != R or a.shape[0] != R:
        raise ValueError("dirs, trace_normal, trace_feature [...,4] and trace_alpha must describe the same rays")
    lut = fg_lut.reshape(fg_lut.shape[-3], fg_lut.shape[-2], 2).contiguous().float()
    hit_n, refl, rough = torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev), torch.empty(R, device=dev)
    pack, local, alpha_n = torch.empty(R, 8, device=dev), torch.empty(R, 3, device=dev), torch.empty(R, device=dev)
    if R == 0:
        return local.view(*shape, 3), alpha_n.view(*shape)
    lib, st = _lib.load(), _stream(dev)
    sat = -1.0 if transmittance_min is None else 1.0 - float(transmittance_min)
    _lib.check(lib.irgs_relight_hit(R, _ptr(d), _ptr(n), _ptr(ft), _ptr(a), sat, _ptr(hit_n), _ptr(refl), _ptr(rough), _ptr(pack), st))
    env_d = envmap(hit_n, mode="diffuse").reshape(R, 3).contiguous().float()                   # __init__.py:370
    env_s = envmap(refl, roughness=rough[:, None], mode="specular").reshape(R, 3).contiguous().float()   # :376
    _lib.check(lib.irgs_relight_combine(R, _ptr(pack), _ptr(env_d), _ptr(env_s), _ptr(lut), lut.shape[0], lut.shape[1], float(f0),
                                        int(bool(wo_indirect_relight)), _ptr(local), _ptr(alpha_n), st))
    return local.view(*shape, 3), alpha_n.view(*shape)


def _rendering_equation_relight(base_color, roughness, normals, position, viewdirs, tracer, surfels, envmap, sample_num,
                                light_sample_num, light_t_min, alpha_min, deg, fg_lut, f0, wo_indirect_relight):
    """The relight branch (evaluation only, like eval_relighting_*.py: under torch.no_grad()).  `surfels`' feature slot holds
    cat([pc.get_base_color, pc.get_rough], 1) [N,4] (__init__.py:363)."""
    from . import incident
    if fg_lut is None:
        raise ValueError("relight=True needs fg_lut (pc.FG_LUT, the split-sum table assets/bsdf_256_256.bin)")
    means3D, opacity, ru, rv, surf_normals, features, shs = surfels
    if features is None or features.shape[-1] != 4:
        raise ValueError("relight=True traces cat([base_color, roughness]) as features: surfels[5] must be [N,4]")
    base, activation, transform = _env_fields(envmap)
    P = base_color.shape[0]
    with torch.no_grad():
        _, t_normal, t_feature, _, t_alpha = tracer.trace_incident(position, normals, sample_num, means3D, opacity, ru, rv,
                                                                   surf_normals, features, shs, alpha_min, t_min=light_t_min, deg=deg)
        dirs = incident.incident_dirs(normals.contiguous(), sample_num)
        local, alpha_n = relight_local_lights(dirs, t_normal, t_feature, t_alpha, envmap, fg_lut, f0, tracer.transmittance_min,
                                              wo_indirect_relight)
        common = dict(activation=activation, transform=transform, transmittance_min=None)
        if light_sample_num == 0:
            return shade_incident(normals, sample_num, base_color, roughness, viewdirs, base, local, alpha_n, **common)
        pdf = getattr(envmap, "_pdf", None)
        if pdf is None:
            raise RuntimeError("light_sample_num > 0 needs the texel probabilities: call envmap.update_pdf() first")
        total = sample_num + light_sample_num
        mix = dict(pdf=pdf, p_diffuse=sample_num / total, p_light=light_sample_num / total, total_samples=total)
        out_d = shade_incident(normals, sample_num, base_color, roughness, viewdirs, base, local, alpha_n, **common, **mix)
        light_dirs, _ = envmap.sample_light_directions(P, light_sample_num, False)
        light_dirs = light_dirs.detach().contiguous()
        _, l_normal, l_feature, _, l_alpha = tracer.trace(position[:, None] + light_dirs * light_t_min, light_dirs, means3D, opacity,
                                                          ru, rv, surf_normals, features, shs, alpha_min, deg=deg)
        local_l, alpha_l = relight_local_lights(light_dirs, l_normal, l_feature, l_alpha, envmap, fg_lut, f0,
                                                tracer.transmittance_min, wo_indirect_relight)
        out_l = shade_incident(normals, light_sample_num, base_color, roughness, viewdirs, base, local_l, alpha_l, dirs=light_dirs,
                               **common, **mix)
        return {k: out_d[k] + out_l[k] for k in out_d}


def rendering_equation(base_color, roughness, normals, position, viewdirs, tracer, surfels, envmap, sample_num,
                       training=False, azimuth=None, light_sample_num=0, light_t_min=0.05, alpha_min=1 / 255, deg=3,
                       relight=False, wo_indirect=False, detach_indirect=False, fg_lut=None, f0=0.04,
                       wo_indirect_relight=False):
    """gaussian_renderer/__init__.py:334-415.  `tracer` + `surfels` stand for the reference's `pc.trace`, `envmap` for
    `pc.get_envmap`, sample_num / light_sample_num / light_t_min / wo_indirect / detach_indirect for the `pipe` fields of
    the same names (arguments/__init__.py:92-101).  training=True draws the per-point random azimuth like
    fibonacci_sphere_sampling(random_rotate=True) unless `azimuth` [P] is given.  light_sample_num > 0 additionally draws
    that many directions per point from `envmap` (its `_pdf` / `sample_light_directions`, like the reference)."""
    if sample_num <= 0 or light_sample_num < 0:
        raise NotImplementedError("diffuse_sample_num > 0 is required (the reference raises likewise, __init__.py:358-359)")
    P = base_color.shape[0]
    keys = ("diffuse", "specular", "light_direct") if training else tuple(OUT_SLICES)
    if P == 0:
        return {k: base_color.new_zeros(0, 1 if k == "visibility" else 3) for k in keys}
    if relight:
        # __init__.py:362-381; evaluation only (eval_relighting_*.py run it under no_grad): fg_lut = pc.FG_LUT, f0 as passed by
        # rendering_equation_chunk, wo_indirect_relight = pipe.wo_indirect_relight; the Fibonacci samples are not rotated
        if training:
            raise NotImplementedError("the relight branch is forward-only (the reference evaluates it under torch.no_grad())")
        out = _rendering_equation_relight(base_color, roughness, normals, position, viewdirs, tracer, surfels, envmap, sample_num,
                                          light_sample_num, light_t_min, alpha_min, deg, fg_lut, f0, wo_indirect_relight)
        return {k: out[k] for k in keys}
    if training and azimuth is None:
        azimuth = torch.rand(P, device=normals.device) * (2 * math.pi)                   # graphics_utils.py:31
    means3D, opacity, ru, rv, surf_normals, features, shs = surfels
    base, activation, transform = _env_fields(envmap)

    def indirect(color, alpha):
        if wo_indirect:
            color = torch.zeros_like(color)                                              # __init__.py:384-385
        if detach_indirect:
            color, alpha = color.detach(), alpha.detach()                                # __init__.py:386-388
        return color, alpha

    color, _, _, _, alpha = tracer.trace_incident(position, normals, sample_num, means3D, opacity, ru, rv, surf_normals,
                                                  features, shs, alpha_min, azimuth=azimuth, t_min=light_t_min, deg=deg)
    color, alpha = indirect(color, alpha)
    if light_sample_num == 0:
        out = shade_incident(normals, sample_num, base_color, roughness, viewdirs, base, color, alpha, azimuth=azimuth,
                             activation=activation, transform=transform, transmittance_min=tracer.transmittance_min)
        return {k: out[k] for k in keys}
    # __init__.py:340-357: Fibonacci samples + directions drawn from the environment map, every sample weighted by the
    # mixed density (evaluated per sample inside the shading kernels); one shading call per kind of sample, results add up
    pdf = getattr(envmap, "_pdf", None)
    if pdf is None:
        raise RuntimeError("light_sample_num > 0 needs the texel probabilities: call envmap.update_pdf() first (render.py:89)")
    total = sample_num + light_sample_num
    mix = dict(pdf=pdf, p_diffuse=sample_num / total, p_light=light_sample_num / total, total_samples=total)
    light_dirs, _ = envmap.sample_light_directions(P, light_sample_num, training)
    light_dirs = light_dirs.detach().contiguous()
    out_d = shade_incident(normals, sample_num, base_color, roughness, viewdirs, base, color, alpha, azimuth=azimuth,
                           activation=activation, transform=transform, transmittance_min=tracer.transmittance_min, **mix)
    color_l, _, _, _, alpha_l = tracer.trace(position[:, None] + light_dirs * light_t_min, light_dirs, means3D, opacity, ru,
                                             rv, surf_normals, features, shs, alpha_min, deg=deg)
    color_l, alpha_l = indirect(color_l, alpha_l)
    out_l = shade_incident(normals, light_sample_num, base_color, roughness, viewdirs, base, color_l, alpha_l,
                           activation=activation, transform=transform, transmittance_min=tracer.transmittance_min,
                           dirs=light_dirs, **mix)
    return {k: out_d[k] + out_l[k] for k in keys}
