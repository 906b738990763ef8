"""CPU restatement of the reference's incident-ray sampling -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/utils/graphics_utils.py:19-47 (fibonacci_sphere_sampling), :133-165 (rotation_between_z) and
the ray construction of /root/reference/gaussian_renderer/__init__.py:324-332,376 (sample_incident_rays; origins
position + dir * light_t_min), in numpy float32 with the reference's order of operations.  Pinned against vectors
recorded from the unmodified reference functions: tests/golden/ref_incident.npz (generator: oracle/gen_golden_incident.py).
"""
import numpy as np

F32 = np.float32
DELTA = F32(np.pi * (3.0 - np.sqrt(5.0)))     # graphics_utils.py:23 (a Python double, applied to a float32 tensor)
SIN10 = F32(np.sin(10 / 180 * np.pi))         # graphics_utils.py:27


def rotation_between_z(vec):
    """graphics_utils.py:133-165: rotation taking +z to `vec` [...,3] (unit); -identity where vec.z + 1 <= 0."""
    vec = np.asarray(vec, F32)
    v1, v2 = -vec[..., 1], vec[..., 0]
    v3 = np.zeros_like(v1)
    v11, v22, v33 = v1 * v1, v2 * v2, v3 * v3
    v12, v13, v23 = v1 * v2, v1 * v3, v2 * v3
    c = np.maximum(vec[..., 2] + F32(1), F32(1e-7))
    R = np.zeros(vec.shape[:-1] + (3, 3), F32)
    R[..., 0, 0] = F32(1) + (-v33 - v22) / c
    R[..., 0, 1] = -v3 + v12 / c
    R[..., 0, 2] = v2 + v13 / c
    R[..., 1, 0] = v3 + v12 / c
    R[..., 1, 1] = F32(1) + (-v33 - v11) / c
    R[..., 1, 2] = -v1 + v23 / c
    R[..., 2, 0] = -v2 + v13 / c
    R[..., 2, 1] = v1 + v23 / c
    R[..., 2, 2] = F32(1) + (-v22 - v11) / c
    flip = ~(vec[..., 2] + F32(1) > 0)
    R[flip] = -np.eye(3, dtype=F32)
    return R


def incident_dirs(normals, sample_num, azimuth=None):
    """graphics_utils.py:19-47.  normals [P,3] float32 unit vectors; azimuth [P] = the `rand * 2 * pi` term of the
    training mode (None: evaluation mode, no random rotation).  Returns unit directions [P, S, 3]."""
    n = np.asarray(normals, F32).reshape(-1, 3)
    idx = np.arange(sample_num, dtype=F32)[None]
    z = np.maximum(F32(1) - F32(2) * idx / F32(2 * sample_num - 1), SIN10)
    rad = np.sqrt(F32(1) - z * z)
    theta = DELTA * idx
    if azimuth is not None:
        theta = np.asarray(azimuth, F32).reshape(-1, 1) + theta
    y = np.cos(theta) * rad
    x = np.sin(theta) * rad
    zs = np.stack([np.broadcast_to(x, y.shape) if azimuth is None else x, y, np.broadcast_to(z, y.shape)], -2)  # [P or 1, 3, S]
    zs = np.broadcast_to(zs, (n.shape[0], 3, sample_num)).astype(F32)
    R = rotation_between_z(n)
    v = np.einsum("pij,pjs->pis", R, zs).astype(F32)
    nrm = np.maximum(np.sqrt((v * v).sum(-2, keepdims=True)), F32(1e-12))
    return np.ascontiguousarray(np.swapaxes(v / nrm, -1, -2)).astype(F32)


def incident_rays(position, normals, sample_num, azimuth=None, t_min=0.05):
    """gaussian_renderer/__init__.py:376: origins position + dirs * light_t_min.  Returns (rays_o, rays_d) [P,S,3]."""
    d = incident_dirs(normals, sample_num, azimuth)
    o = (np.asarray(position, F32).reshape(-1, 1, 3) + d * F32(t_min)).astype(F32)
    return o, d
