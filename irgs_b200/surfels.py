"""Tracing straight from surfel parameters (means, scales, rotations, opacities): the input form BASELINE.json's
north_star names, and the caller glue of the reference restated once so that users need not copy it.

Mirrors scene/gaussian_model.py:712-756 of the reference:
  * `surfel_frames`     :733-747  ru = R[:,:,0] / s_u, rv = R[:,:,1] / s_v, normals = R[:,:,2] flipped towards the camera
  * `SurfelScene.build / refit` :725-731 (get_boundings + build_bvh / update_bvh) -- here the bounds are derived
    analytically from the parameters by the native library, no 12N-vertex proxy mesh is materialised
  * `SurfelScene.trace` :733-765 incl. the normalisation by alpha where alpha >= 1 - T_min (:751-756)
  * `SurfelScene.rendering_equation`: the reference's rendering_equation (gaussian_renderer/__init__.py:334-415) from the
    same parameters -- incident rays generated in the kernels, traced, shaded by the epilogue kernels (irgs_b200/shading.py)
Everything between the parameters and the tracer is plain differentiable torch, so gradients reach scales and
rotations (quaternions) through autograd exactly as they do in IRGS.
"""
import torch

from .raytracer import GaussianTracer


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
    def build(self, means, scales, rotations, opacities, camera_center=None):
        ru, rv, normals = surfel_frames(means, scales, rotations, camera_center)
        self.tracer.build_from_surfels(means, opacities, ru, rv, normals, self.alpha_min)
        self._built_n = means.shape[0]

    @torch.no_grad()
    def refit(self, means, scales, rotations, opacities, camera_center=None):
        """Per-iteration update with frozen topology (train.py:150-154 calls update_bvh when geometry moves)."""
        if self._built_n != means.shape[0]:
            return self.build(means, scales, rotations, opacities, camera_center)
        ru, rv, normals = surfel_frames(means, scales, rotations, camera_center)
        self.tracer.update_from_surfels(means, opacities, ru, rv, normals, self.alpha_min)

    def trace(self, rays_o, rays_d, means, scales, rotations, opacities, shs, features=None, camera_center=None,
              deg=3, back_culling=False, normalize=True):
        """dict(color, normal, feature, depth, alpha, hit_count, normals).  With `normalize` the accumulations of rays
        that saturated (alpha >= 1 - T_min) are divided by alpha and their alpha set to 1, as GaussianModel.trace does."""
        ru, rv, normals = surfel_frames(means, scales, rotations, camera_center)
        color, normal, feature, depth, alpha = self.tracer.trace(
            rays_o, rays_d, means, opacities, ru, rv, normals, features, shs, self.alpha_min, deg=deg,
            back_culling=back_culling)
        hit_count = self.tracer.last_hit_count
        if normalize:
            sat = alpha >= 1 - self.tracer.transmittance_min
            a_ = alpha[..., None]
            color = torch.where(sat[..., None], color / a_, color)
            normal = torch.where(sat[..., None], normal / a_, normal)
            feature = torch.where(sat[..., None], feature / a_, feature)
            depth = torch.where(sat, depth / alpha, depth)
            alpha = torch.where(sat, torch.ones_like(alpha), alpha)
        return dict(color=color, normal=normal, feature=feature, depth=depth, alpha=alpha, hit_count=hit_count,
                    normals=normals)

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
