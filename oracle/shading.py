"""CPU restatement of the reference's shading around the incident-ray trace -- TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Differentiable torch code on the CPU (any dtype; the tests use float64 for gradient checks), each function citing the
reference lines it follows (paths under /root/reference):

  texture_linear_wrap   nvdiffrast.torch.texture(filter_mode='linear', boundary_mode='wrap')  -- the package (pinned by the
                        reference's environment to nvdiffrast 0.3.x) is NOT in this image, so its published bilinear scheme
                        (texture.cu: indexTextureLinear) is restated: texel centres at (i + 0.5) / size, wrap-around
                        neighbours.  This one function is therefore "parity unpinned"; everything below it is pinned
                        against golden vectors recorded from the UNMODIFIED reference functions with this lookup plugged
                        in as `dr.texture` (tests/golden/ref_shading.npz, generator oracle/gen_golden_shading.py).
  env_pure              scene/light.py:287-297,315   EnvLight.__call__(mode='pure_env')
  normalise_trace       scene/gaussian_model.py:748-752   GaussianModel.trace post-processing of colour and alpha
  ggx_specular          gaussian_renderer/__init__.py:417-457
  rendering_equation    gaussian_renderer/__init__.py:334-415 (diffuse_sample_num > 0, relight=False; with light_sample_num > 0
                        through mis_areas)
  update_pdf, light_dirs_from_texels, light_pdf, mis_areas    scene/light.py:174-223 + __init__.py:340-357 (light-importance
                        sampling mixed with the Fibonacci samples)
"""
import math

import torch
import torch.nn.functional as F


def texture_linear_wrap(tex, uv):
    """tex [H, W, C], uv [..., 2] in texture units -> [..., C]."""
    H, W = tex.shape[:2]
    u, v = uv[..., 0], uv[..., 1]
    u = u - torch.floor(u)
    v = v - torch.floor(v)
    u = u * W - 0.5
    v = v * H - 0.5
    iu0, iv0 = torch.floor(u).long(), torch.floor(v).long()
    fu, fv = (u - iu0)[..., None], (v - iv0)[..., None]
    iu1, iv1 = iu0 + 1, iv0 + 1
    iu0 = torch.where(iu0 < 0, iu0 + W, iu0)
    iv0 = torch.where(iv0 < 0, iv0 + H, iv0)
    iu1 = torch.where(iu1 >= W, iu1 - W, iu1)
    iv1 = torch.where(iv1 >= H, iv1 - H, iv1)
    a00, a10, a01, a11 = tex[iv0, iu0], tex[iv0, iu1], tex[iv1, iu0], tex[iv1, iu1]
    top = a00 + (a10 - a00) * fu
    bot = a01 + (a11 - a01) * fu
    return top + (bot - top) * fv


ACTIVATIONS = {"exp": torch.exp, "sigmoid": torch.sigmoid, "none": lambda x: x}


def env_pure(base, dirs, activation="exp", transform=None):
    """light.py:287-297,315.  base [H,W,3] pre-activation, dirs [...,3] -> radiance [...,3]."""
    l = dirs
    if transform is not None:
        l = l @ transform.T
    uv = torch.cat([torch.atan2(l[..., :1], -l[..., 2:3]).nan_to_num() / (2.0 * math.pi) + 0.5,
                    torch.acos(l[..., 1:2].clamp(-1.0 + 1e-6, 1.0 - 1e-6)) / math.pi], dim=-1).clamp(0, 1)
    return ACTIVATIONS[activation](texture_linear_wrap(base, uv)).clamp_min(0.0)


def normalise_trace(color, alpha, transmittance_min):
    """gaussian_model.py:748-752.  color [...,3], alpha [...] raw tracer outputs."""
    a_ = alpha[..., None]
    sat = 1 - transmittance_min
    color = torch.where(a_ < sat, color, color / a_)
    alpha = torch.where(alpha < sat, alpha, torch.ones_like(alpha))
    return color, alpha


def ggx_specular(normal, pts2c, pts2l, roughness, fresnel):
    """__init__.py:417-457 (out-of-place clamps: same values and gradients as the reference's clamp_)."""
    L = F.normalize(pts2l, dim=-1)
    V = F.normalize(pts2c, dim=-1)
    H = F.normalize((L + V[:, None, :]) / 2.0, dim=-1)
    N = F.normalize(normal, dim=-1)
    NoV = torch.sum(V * N, dim=-1, keepdim=True)
    N = N * NoV.sign()
    NoL = torch.sum(N[:, None, :] * L, dim=-1, keepdim=True).clamp(1e-6, 1)
    NoV = torch.sum(N * V, dim=-1, keepdim=True).clamp(1e-6, 1)
    NoH = torch.sum(N[:, None, :] * H, dim=-1, keepdim=True).clamp(1e-6, 1)
    VoH = torch.sum(V[:, None, :] * H, dim=-1, keepdim=True).clamp(1e-6, 1)
    alpha = roughness * roughness
    alpha2 = alpha * alpha
    k = (alpha + 2 * roughness + 1.0) / 8.0
    FMi = ((-5.55473) * VoH - 6.98316) * VoH
    frac0 = fresnel + (1 - fresnel) * torch.pow(2.0, FMi)
    frac = frac0 * alpha2[:, None, :]
    nom0 = NoH * NoH * (alpha2[:, None, :] - 1) + 1
    nom1 = NoV * (1 - k) + k
    nom2 = NoL * (1 - k[:, None, :]) + k[:, None, :]
    nom = (4 * math.pi * nom0 * nom0 * nom1[:, None, :] * nom2).clamp(1e-6, 4 * math.pi)
    return frac / nom


def update_pdf(base, activation="exp"):
    """light.py:174-179 EnvLight.update_pdf: per-texel sampling probability [H,W] (no gradient)."""
    with torch.no_grad():
        H, W = base.shape[:2]
        Y = ((torch.arange(0, H, dtype=torch.float32, device=base.device) + 0.5) / H)[:, None].expand(H, W)   # pixel_grid(...)[..., 1]
        pdf = torch.max(ACTIVATIONS[activation](base).clamp_min(0.0), dim=-1)[0] * torch.sin(Y * math.pi)
        return pdf / torch.sum(pdf)


def light_dirs_from_texels(idx, H, W, jitter_x=None, jitter_y=None, transform=None):
    """light.py:186-202 EnvLight.sample_light_directions after the multinomial draw: texel index -> direction.
    jitter_* = the training mode's `rand - 0.5` terms (None: texel centres)."""
    gx = ((idx % W + 0.5) / W) * 2 - 1
    gy = (idx // W + 0.5) / H
    if jitter_x is not None:
        gx = gx + jitter_x / W * 2
        gy = gy + jitter_y / H
    sintheta, costheta = torch.sin(gy * math.pi), torch.cos(gy * math.pi)
    sinphi, cosphi = torch.sin(gx * math.pi), torch.cos(gx * math.pi)
    d = torch.stack((sintheta * sinphi, costheta, -sintheta * cosphi), dim=-1)
    if transform is not None:
        d = d @ transform
    return d


def light_pdf(pdf, direction, transform=None):
    """light.py:207-223 EnvLight.light_pdf: solid-angle density of the texel sampling at `direction` [...,3] -> [...,1]."""
    H, W = pdf.shape[:2]
    flat = direction.reshape(-1, 3)
    if transform is not None:
        flat = flat @ transform.T
    u = torch.atan2(flat[..., 0], -flat[..., 2]).nan_to_num() / (2.0 * math.pi) + 0.5
    v = torch.acos(flat[..., 1].clamp(-1.0 + 1e-6, 1.0 - 1e-6)) / math.pi
    u_idx = (u * W).clamp(0, W - 1).long()
    v_idx = (v * H).clamp(0, H - 1).long()
    weight = H * W / (2.0 * math.pi ** 2 * torch.sin(v * math.pi).clamp_min(1e-6))
    return (pdf.reshape(-1)[u_idx + v_idx * W] * weight).reshape(*direction.shape[:-1], 1)


def mis_areas(incident_dirs, pdf, n_diffuse, n_light, transform=None):
    """__init__.py:340-357: per-sample `incident_areas` [P,S,1] of the mixed diffuse / light-importance sampling, S = n_diffuse +
    n_light (the same expression for both kinds of samples: 1/(2 pi) * p_diffuse + light_pdf(dir) * p_light)."""
    p_d = n_diffuse / (n_diffuse + n_light)
    p_l = n_light / (n_diffuse + n_light)
    mix = 1 / (2 * math.pi) * p_d + light_pdf(pdf, incident_dirs, transform) * p_l
    return 1 / mix.clamp_min(1e-6)


def rendering_equation(base_color, roughness, normals, viewdirs, incident_dirs, trace_color, trace_alpha, env_base,
                       activation="exp", transform=None, transmittance_min=None, incident_areas=None):
    """__init__.py:334-415 given the incident directions [P,S,3] and the tracer's RAW colour [P,S,3] / alpha [P,S] of those
    rays (transmittance_min: GaussianModel.trace's normalisation, None to skip).  incident_areas: [P,S,1] from mis_areas
    for the light_sample_num > 0 branch (None: the 2 pi of pure Fibonacci sampling).  Returns the evaluation-mode dict (the
    training-mode dict is its subset diffuse / specular / light_direct)."""
    if transmittance_min is not None:
        trace_color, trace_alpha = normalise_trace(trace_color, trace_alpha, transmittance_min)
    global_incident_lights = env_pure(env_base, incident_dirs, activation, transform)
    incident_visibility = 1 - trace_alpha[..., None]
    local_incident_lights = trace_color
    incident_lights = incident_visibility * global_incident_lights + local_incident_lights
    if incident_areas is None:
        incident_areas = 2 * math.pi                               # graphics_utils.py:43
    n_d_i = (normals[:, None] * incident_dirs).sum(-1, keepdim=True).clamp(min=0)
    f_d = base_color[:, None] / math.pi
    f_s = ggx_specular(normals, viewdirs, incident_dirs, roughness, fresnel=0.04)
    transport = incident_lights * incident_areas * n_d_i
    return {
        "diffuse": (f_d * transport).mean(dim=-2),
        "specular": (f_s * transport).mean(dim=-2),
        "visibility": incident_visibility.mean(dim=1),
        "light": incident_lights.mean(dim=1),
        "light_indirect": local_incident_lights.mean(dim=1),
        "light_direct": global_incident_lights.mean(dim=1),
    }


def fibonacci_dirs(normals, sample_num, azimuth=None):
    """Differentiable torch restatement of utils/graphics_utils.py:19-47 + :133-165 (the numpy twin, pinned against the
    reference's golden vectors, is oracle/incident.py)."""
    P = normals.shape[0]
    dt = normals.dtype
    idx = torch.arange(sample_num, dtype=dt)[None]
    z = (1 - 2 * idx / (2 * sample_num - 1)).clamp_min(math.sin(10 / 180 * math.pi))
    rad = torch.sqrt(1 - z ** 2)
    theta = math.pi * (3.0 - math.sqrt(5.0)) * idx
    if azimuth is not None:
        theta = azimuth.reshape(-1, 1).to(dt) + theta
    zs = torch.stack([(torch.sin(theta) * rad).expand(P, sample_num), (torch.cos(theta) * rad).expand(P, sample_num),
                      z.expand(P, sample_num)], -2)
    v1, v2 = -normals[..., 1], normals[..., 0]
    c = (normals[..., 2] + 1).clamp_min(1e-7)
    R = torch.stack([1 + (-v2 * v2) / c, v1 * v2 / c, v2,
                     v1 * v2 / c, 1 + (-v1 * v1) / c, -v1,
                     -v2, v1, 1 + (-v2 * v2 - v1 * v1) / c], -1).reshape(P, 3, 3)
    flip = -torch.eye(3, dtype=dt).expand(P, 3, 3)
    R = torch.where((normals[..., 2] + 1 > 0)[:, None, None], R, flip)
    return F.normalize(R @ zs, dim=-2).transpose(-1, -2)


# ------------------------------------------------------------------------------------------------ relight branch
def texture_linear_clamp(tex, uv):
    """nvdiffrast.torch.texture(filter_mode='linear', boundary_mode='clamp') restated like texture_linear_wrap (texel centres
    at (i + 0.5) / size, neighbour indices clamped to the texture): tex [H, W, C], uv [..., 2] -> [..., C].  Used for the FG
    table lookup of gaussian_renderer/__init__.py:375; "parity unpinned" like the wrap lookup (nvdiffrast is not in this image)."""
    H, W = tex.shape[:2]
    u = uv[..., 0].clamp(0, 1) * W - 0.5
    v = uv[..., 1].clamp(0, 1) * H - 0.5
    iu0, iv0 = torch.floor(u).long(), torch.floor(v).long()
    fu, fv = (u - iu0)[..., None], (v - iv0)[..., None]
    iu1, iv1 = (iu0 + 1).clamp(0, W - 1), (iv0 + 1).clamp(0, H - 1)
    iu0, iv0 = iu0.clamp(0, W - 1), iv0.clamp(0, H - 1)
    a00, a10, a01, a11 = tex[iv0, iu0], tex[iv0, iu1], tex[iv1, iu0], tex[iv1, iu1]
    top = a00 + (a10 - a00) * fu
    bot = a01 + (a11 - a01) * fu
    return top + (bot - top) * fv


class RelightEnvStandIn(torch.nn.Module):
    """A test double for the reference's EnvLight with ALL THREE lookup modes (scene/light.py:264-328).  The real modes
    'diffuse' / 'specular' read cube-map mips prefiltered by nvdiffrec's renderutils through nvdiffrast -- neither is in this
    image and both are out of scope -- so this stand-in answers them from lat-long maps: 'diffuse' from
    `base_diffuse`, 'specular' from a roughness-weighted blend of `base_spec0` / `base_spec1`.  'pure_env' is env_pure, the
    reference's own lat-long path.  The reference's rendering_equation (relight=True) runs unmodified around it when the golden
    vectors are recorded, and the same object (moved to the GPU) is handed to irgs_b200.shading.rendering_equation in the tests."""

    def __init__(self, base, base_diffuse, base_spec0, base_spec1, activation="exp", transform=None):
        super().__init__()
        self.base = torch.nn.Parameter(base.clone(), requires_grad=False)
        self.base_diffuse, self.base_spec0, self.base_spec1 = base_diffuse, base_spec0, base_spec1
        self.activation_name, self.transform = activation, transform

    def to(self, device):
        self.base.data = self.base.data.to(device)
        self.base_diffuse, self.base_spec0, self.base_spec1 = (t.to(device) for t in (self.base_diffuse, self.base_spec0, self.base_spec1))
        return self

    def _uv(self, l):
        if self.transform is not None:
            l = l @ torch.as_tensor(self.transform, dtype=l.dtype, device=l.device).T
        return torch.cat([torch.atan2(l[..., :1], -l[..., 2:3]).nan_to_num() / (2.0 * math.pi) + 0.5,
                          torch.acos(l[..., 1:2].clamp(-1.0 + 1e-6, 1.0 - 1e-6)) / math.pi], dim=-1).clamp(0, 1)

    def __call__(self, l, mode="pure_env", roughness=None):
        act = ACTIVATIONS[self.activation_name]
        if mode == "pure_env":
            return env_pure(self.base, l, self.activation_name, self.transform)
        if mode == "diffuse":
            return act(texture_linear_wrap(self.base_diffuse, self._uv(l))).clamp_min(0.0)
        r = roughness.reshape(*l.shape[:-1], 1).clamp(0, 1)
        uv = self._uv(l)
        return act((1 - r) * texture_linear_wrap(self.base_spec0, uv) + r * texture_linear_wrap(self.base_spec1, uv)).clamp_min(0.0)


def relight_local(dirs, normal_raw, feature_raw, alpha_raw, envmap, fg_lut, f0=0.04, transmittance_min=None,
                  wo_indirect_relight=False):
    """__init__.py:363-379 given the tracer's RAW normal [...,3] / feature [...,4] / alpha [...] (GaussianModel.trace's
    normalisation, scene/gaussian_model.py:751-756, applied here when transmittance_min is given).  Returns
    (local_incident_lights [...,3], trace_alpha [...])."""
    alpha = alpha_raw
    normal, feature = normal_raw, feature_raw
    if transmittance_min is not None:
        a_ = alpha[..., None]
        sat = a_ < 1 - transmittance_min
        normal = torch.where(sat, normal, normal / a_)
        feature = torch.where(sat, feature, feature / a_)
        alpha = torch.where(alpha < 1 - transmittance_min, alpha, torch.ones_like(alpha))
    trace_alpha = alpha[..., None]
    trace_feature = feature / trace_alpha.clamp_min(1e-6)
    trace_normal = F.normalize(normal, dim=-1)
    trace_base_color, trace_roughness = trace_feature.split([3, 1], dim=-1)
    trace_diffuse = trace_base_color * envmap(trace_normal, mode="diffuse")
    trace_wi = -dirs
    trace_NdotV = (trace_normal * trace_wi).sum(-1, keepdim=True)
    trace_reflected = F.normalize(trace_NdotV * trace_normal * 2 - trace_wi, dim=-1)
    fg_uv = torch.cat([trace_NdotV, trace_roughness], -1).clamp(0, 1)
    fg = texture_linear_clamp(fg_lut.reshape(fg_lut.shape[-3], fg_lut.shape[-2], 2), fg_uv)
    trace_specular = envmap(trace_reflected, roughness=trace_roughness, mode="specular") * (f0 * fg[..., 0:1] + fg[..., 1:2])
    local = (trace_diffuse + trace_specular) * trace_alpha
    if wo_indirect_relight:
        local = torch.zeros_like(local)
    return local, alpha
