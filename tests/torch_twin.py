"""Differentiable float64 pure-torch twin of the tracer's compositing math, used to pin the oracle's (and through it
the CUDA kernels') gradients with torch autograd -- the nvdiffrec test pattern of the reference
(scene/renderutils/tests/test_bsdf.py: CUDA op vs use_python=True twin on random tensors).

The twin consumes the ORDERED HIT LISTS produced by the oracle (the discrete part) and re-computes everything
continuous.  The reference backward's documented quirks are reproduced with straight-through
estimators so that plain autograd yields the reference's formulas:
  * min(0.99, .) and max(., 0) are applied in the value but ignored in the derivative,
  * the SH colour gets no gradient with respect to the ray direction.
"""
import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def sh_basis(deg, d):
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    Y = [torch.full_like(x, SH_C0)]
    if deg > 0:
        Y += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        Y += [SH_C2[0] * xy, SH_C2[1] * yz, SH_C2[2] * (2 * zz - xx - yy), SH_C2[3] * xz, SH_C2[4] * (xx - yy)]
    if deg > 2:
        Y += [SH_C3[0] * y * (3 * xx - yy), SH_C3[1] * xy * z, SH_C3[2] * y * (4 * zz - xx - yy),
              SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy), SH_C3[4] * x * (4 * zz - xx - yy), SH_C3[5] * z * (xx - yy),
              SH_C3[6] * x * (xx - 3 * yy)]
    return torch.stack(Y, -1)


def composite(rays_o, rays_d, means, opacity, ru, rv, normals, features, shs, hits, hit_count, deg=3):
    """hits [R,H] long (padded with anything), hit_count [R].  Returns color, normal, feature, depth, alpha."""
    R, H = hits.shape
    valid = torch.arange(H)[None] < hit_count[:, None]
    g = torch.where(valid, hits, torch.zeros_like(hits))
    o, d = rays_o[:, None], rays_d[:, None]
    mu, n, a, b = means[g], normals[g], ru[g], rv[g]
    rel = o - mu
    og = (n * rel).sum(-1)
    dg = (n * d).sum(-1)
    t = -og * dg / torch.clamp(dg * dg, min=1e-6)
    pos = rel + t[..., None] * d
    pu, pv = (a * pos).sum(-1), (b * pos).sum(-1)
    G = torch.exp(-0.5 * (pu * pu + pv * pv))
    a_raw = opacity.reshape(-1)[g] * G
    alpha = a_raw + (torch.clamp(a_raw, max=0.99) - a_raw).detach()
    alpha = torch.where(valid, alpha, torch.zeros_like(alpha))
    Y = sh_basis(deg, rays_d.detach())  # [R,nb]; no direction gradient through the colour
    nb = Y.shape[-1]
    c_raw = 0.5 + (Y[:, None, :, None] * shs[g][:, :, :nb, :]).sum(-2)
    c = c_raw + (torch.clamp(c_raw, min=0.0) - c_raw).detach()
    T = torch.cumprod(torch.cat([torch.ones(R, 1, dtype=alpha.dtype), 1 - alpha[:, :-1]], 1), 1)
    w = T * alpha
    m = torch.where(-dg > 0, 1.0, -1.0).detach()
    color = (w[..., None] * c).sum(1)
    normal = (w[..., None] * m[..., None] * n).sum(1)
    depth = (w * t).sum(1)
    alpha_out = w.sum(1)
    feature = (w[..., None] * features[g]).sum(1)
    return color, normal, feature, depth, alpha_out
