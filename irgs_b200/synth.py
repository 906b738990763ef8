"""Deterministic synthetic surfel scenes and ray batches of the shapes BASELINE.md section 3 names.

Everything is generated with seeded torch CPU generators (scene 1234, rays 5678, grad_out 9012) and then moved to
the requested device, so the same inputs can be handed to the CUDA tracer, the CPU oracle and the reference tracer.

The derived per-surfel inputs follow the caller glue of the reference, scene/gaussian_model.py:733-747
(ru = R[:,:,0]/s_u, rv = R[:,:,1]/s_v, normals = R[:,:,2] flipped towards the camera), and the proxy mesh follows
scene/gaussian_model.py:111-116,712-723 (icosahedron with unit in-sphere, squashed by (s_u, s_v, 1e-6) and scaled by
sqrt(2 ln(opacity/alpha_min))).  The secondary rays follow utils/graphics_utils.py:19-47,133-165 (Fibonacci
hemisphere around the shading normal) and gaussian_renderer/__init__.py:382 (origin = x + light_t_min * dir).
"""
import math

import numpy as np
import torch

SCENE_SEED, RAY_SEED, GRAD_SEED = 1234, 5678, 9012
ALPHA_MIN = 1.0 / 255.0      # scene/gaussian_model.py:119
T_MIN = 0.03                 # scene/gaussian_model.py:118
LIGHT_T_MIN = 0.05           # arguments/__init__.py:96


def quat_to_rot(q):
    """utils/general_utils.py:78-99 (w, x, y, z) -> R[N,3,3]."""
    q = q / q.norm(dim=-1, keepdim=True)
    r, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(*q.shape[:-1], 3, 3)


def _lego_surface_samples(n, gen):
    """Points + outward normals on a lego-like union of axis-aligned box shells and studs inside [-1,1]^3."""
    boxes = [  # centre, half-extent
        ((0.0, 0.0, -0.55), (0.90, 0.60, 0.20)),
        ((-0.30, 0.0, -0.10), (0.55, 0.45, 0.25)),
        ((0.45, 0.0, 0.15), (0.30, 0.40, 0.50)),
        ((-0.55, 0.0, 0.45), (0.25, 0.30, 0.30)),
        ((0.0, 0.0, 0.80), (0.70, 0.15, 0.10)),
    ]
    studs = [((x, y, -0.30), 0.08, 0.05) for x in (-0.7, -0.35, 0.0, 0.35, 0.7) for y in (-0.4, 0.0, 0.4)]
    areas = []
    for _, h in boxes:
        areas.append(8.0 * (h[0] * h[1] + h[1] * h[2] + h[0] * h[2]))
    for _, r, hh in studs:
        areas.append(2 * math.pi * r * (2 * hh) + math.pi * r * r)
    areas = torch.tensor(areas, dtype=torch.float64)
    prim = torch.multinomial(areas / areas.sum(), n, replacement=True, generator=gen)
    pts = torch.zeros(n, 3, dtype=torch.float64)
    nrm = torch.zeros(n, 3, dtype=torch.float64)
    u = torch.rand(n, 3, generator=gen, dtype=torch.float64)
    for i, (c, h) in enumerate(boxes):
        m = prim == i
        k = int(m.sum())
        if k == 0:
            continue
        c = torch.tensor(c, dtype=torch.float64)
        h = torch.tensor(h, dtype=torch.float64)
        fa = torch.stack([h[1] * h[2], h[0] * h[2], h[0] * h[1]])  # area weight of the faces normal to x, y, z
        axis = torch.multinomial(fa / fa.sum(), k, replacement=True, generator=gen)
        sign = (torch.rand(k, generator=gen, dtype=torch.float64) < 0.5).double() * 2 - 1
        p = (u[m] * 2 - 1) * h
        p[torch.arange(k), axis] = sign * h[axis]
        nn = torch.zeros(k, 3, dtype=torch.float64)
        nn[torch.arange(k), axis] = sign
        pts[m] = p + c
        nrm[m] = nn
    for j, (c, r, hh) in enumerate(studs):
        m = prim == len(boxes) + j
        k = int(m.sum())
        if k == 0:
            continue
        c = torch.tensor(c, dtype=torch.float64)
        side_area, cap_area = 2 * math.pi * r * 2 * hh, math.pi * r * r
        on_cap = u[m][:, 0] < cap_area / (side_area + cap_area)
        ang = u[m][:, 1] * 2 * math.pi
        rad = torch.where(on_cap, r * u[m][:, 2].sqrt(), torch.full((k,), r, dtype=torch.float64))
        z = torch.where(on_cap, torch.full((k,), hh, dtype=torch.float64), (u[m][:, 2] * 2 - 1) * hh)
        p = torch.stack([rad * ang.cos(), rad * ang.sin(), z], -1)
        nn = torch.stack([ang.cos(), ang.sin(), torch.zeros(k, dtype=torch.float64)], -1)
        nn[on_cap] = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
        pts[m] = p + c
        nrm[m] = nn
    return pts.float(), nrm.float()


def _rot_from_z(n):
    """Rotation taking +z to unit vector n (Rodrigues form of utils/graphics_utils.py:133-165)."""
    v1, v2 = -n[..., 1], n[..., 0]
    c1 = (n[..., 2] + 1).clamp_min(1e-7)
    zero = torch.zeros_like(v1)
    R = torch.stack([
        1 - v2 * v2 / c1, v1 * v2 / c1, v2,
        v1 * v2 / c1, 1 - v1 * v1 / c1, -v1,
        -v2, v1, 1 - (v1 * v1 + v2 * v2) / c1], dim=-1).reshape(*n.shape[:-1], 3, 3)
    del zero
    flip = -torch.eye(3, dtype=n.dtype, device=n.device).expand_as(R)
    return torch.where((n[..., 2] + 1 > 0)[..., None, None], R, flip)


def _rot_to_quat(R):
    """Rotation matrices -> (w,x,y,z), numerically safe branch selection."""
    m = R
    t = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    q = torch.zeros(R.shape[0], 4, dtype=R.dtype)
    c0 = t > 0
    s = torch.sqrt(t.clamp_min(-0.999) + 1.0) * 2
    q[c0] = torch.stack([0.25 * s, (m[:, 2, 1] - m[:, 1, 2]) / s, (m[:, 0, 2] - m[:, 2, 0]) / s,
                         (m[:, 1, 0] - m[:, 0, 1]) / s], -1)[c0]
    c1 = (~c0) & (m[:, 0, 0] > m[:, 1, 1]) & (m[:, 0, 0] > m[:, 2, 2])
    s1 = torch.sqrt((1.0 + m[:, 0, 0] - m[:, 1, 1] - m[:, 2, 2]).clamp_min(1e-12)) * 2
    q[c1] = torch.stack([(m[:, 2, 1] - m[:, 1, 2]) / s1, 0.25 * s1, (m[:, 0, 1] + m[:, 1, 0]) / s1,
                         (m[:, 0, 2] + m[:, 2, 0]) / s1], -1)[c1]
    c2 = (~c0) & (~c1) & (m[:, 1, 1] > m[:, 2, 2])
    s2 = torch.sqrt((1.0 + m[:, 1, 1] - m[:, 0, 0] - m[:, 2, 2]).clamp_min(1e-12)) * 2
    q[c2] = torch.stack([(m[:, 0, 2] - m[:, 2, 0]) / s2, (m[:, 0, 1] + m[:, 1, 0]) / s2, 0.25 * s2,
                         (m[:, 1, 2] + m[:, 2, 1]) / s2], -1)[c2]
    c3 = (~c0) & (~c1) & (~c2)
    s3 = torch.sqrt((1.0 + m[:, 2, 2] - m[:, 0, 0] - m[:, 1, 1]).clamp_min(1e-12)) * 2
    q[c3] = torch.stack([(m[:, 1, 0] - m[:, 0, 1]) / s3, (m[:, 0, 2] + m[:, 2, 0]) / s3,
                         (m[:, 1, 2] + m[:, 2, 1]) / s3, 0.25 * s3], -1)[c3]
    return q


def make_scene(n, seed=SCENE_SEED, n_features=0, sh_coeffs=16, device="cpu", scale_mult=1.0):
    """Raw surfel parameters of a lego-shaped scene (BASELINE.md section 3 / SURVEY.md 8d).

    Returns dict(means[N,3], scales[N,2], rotations[N,4] (w,x,y,z), opacity[N,1], shs[N,K,3], features[N,S]).
    scale_mult > 1 enlarges the surfels so that a small-N test scene has the per-ray hit density of the 300k one."""
    gen = torch.Generator().manual_seed(seed)
    pts, nrm = _lego_surface_samples(n, gen)
    means = pts + 0.002 * torch.randn(n, 3, generator=gen)
    # surface normal perturbed by ~5 degrees, in-plane angle uniform
    tilt = math.radians(5.0) * torch.randn(n, 2, generator=gen)
    Rn = _rot_from_z(nrm)
    local = torch.stack([tilt[:, 0], tilt[:, 1], torch.ones(n)], -1)
    local = local / local.norm(dim=-1, keepdim=True)
    normal = (Rn @ local[..., None])[..., 0]
    phi = 2 * math.pi * torch.rand(n, generator=gen)
    Rz = torch.zeros(n, 3, 3)
    Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = phi.cos(), -phi.sin(), phi.sin(), phi.cos(), 1.0
    R = _rot_from_z(normal) @ Rz
    rotations = _rot_to_quat(R)
    scales = torch.exp(math.log(0.004) + (math.log(0.02) - math.log(0.004)) * torch.rand(n, 2, generator=gen)) * scale_mult
    opacity = torch.sigmoid(3.0 + 2.0 * torch.randn(n, 1, generator=gen)).clamp(0.02, 0.999)
    shs = torch.cat([torch.randn(n, 1, 3, generator=gen), 0.1 * torch.randn(n, sh_coeffs - 1, 3, generator=gen)], 1)
    features = torch.rand(n, n_features, generator=gen)
    out = dict(means=means, scales=scales, rotations=rotations, opacity=opacity, shs=shs, features=features)
    return {k: v.contiguous().to(device) for k, v in out.items()}


def derive_tracer_inputs(scene, camera_center=None):
    """scene/gaussian_model.py:733-747: (means, opacity, ru, rv, normals, features, shs) as the tracer takes them."""
    R = quat_to_rot(scene["rotations"])
    s = 1.0 / scene["scales"]
    ru = R[:, :, 0] * s[:, 0:1]
    rv = R[:, :, 1] * s[:, 1:2]
    normals = R[:, :, 2]
    if camera_center is not None:
        cc = torch.as_tensor(camera_center, dtype=normals.dtype, device=normals.device)
        dotp = (normals * -(scene["means"] - cc)).sum(-1, keepdim=True)
        normals = normals * torch.where(dotp >= 0, 1.0, -1.0)
    normals = normals / normals.norm(dim=-1, keepdim=True).clamp_min(1e-20)
    return dict(means3D=scene["means"].contiguous(), opacity=scene["opacity"].contiguous(), ru=ru.contiguous(),
                rv=rv.contiguous(), normals=normals.contiguous(), features=scene["features"].contiguous(),
                shs=scene["shs"].contiguous())


_ICO = None


def unit_icosahedron():
    """12 vertices (scaled so the in-sphere has radius 1: x1.2584, scene/gaussian_model.py:115) and 20
    outward-facing counter-clockwise faces.  Stands in for trimesh.creation.icosahedron() (not installed)."""
    global _ICO
    if _ICO is None:
        phi = (1 + 5 ** 0.5) / 2
        v = []
        for a in (-1, 1):
            for b in (-phi, phi):
                v += [(0, a, b), (a, b, 0), (b, 0, a)]
        v = np.array(v, dtype=np.float64)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        from scipy.spatial import ConvexHull
        faces = ConvexHull(v).simplices.copy()
        for f in faces:  # orient outward
            a, b, c = v[f]
            if np.dot(np.cross(b - a, c - a), a + b + c) < 0:
                f[1], f[2] = f[2], f[1]
        _ICO = (torch.from_numpy(v).float() * 1.2584, torch.from_numpy(faces.astype(np.int64)))
    return _ICO


def proxy_mesh(scene, alpha_min=ALPHA_MIN):
    """GaussianModel.get_boundings (scene/gaussian_model.py:712-723): vertices_b[12N,3], faces_b[20N,3], gs_id[20N]."""
    dev = scene["means"].device
    verts, faces = unit_icosahedron()
    verts, faces = verts.to(dev), faces.to(dev)
    n = scene["means"].shape[0]
    scale3 = torch.cat([scene["scales"], torch.full_like(scene["scales"][:, :1], 1e-6)], -1)
    L = quat_to_rot(scene["rotations"]) * scale3[:, None, :]
    r = (2 * (scene["opacity"] / alpha_min).log()).sqrt()
    vertices_b = r[:, None] * (verts[None] @ L.transpose(-1, -2)) + scene["means"][:, None]
    faces_b = faces[None] + torch.arange(n, device=dev)[:, None, None] * 12
    gs_id = torch.arange(n, device=dev)[:, None].expand(-1, 20)
    return vertices_b.reshape(-1, 3).contiguous(), faces_b.reshape(-1, 3).contiguous(), gs_id.reshape(-1).contiguous()


CAMERA_CENTER = (0.0, -4.0, 0.0)  # radius 4, looking at the origin along +y, z up


def primary_rays(h, w, camera_angle_x=0.6911, device="cpu"):
    """Pinhole rays (scene/cameras.py:87-100 convention: pixel centres, unit directions)."""
    focal = 0.5 * w / math.tan(0.5 * camera_angle_x)
    j, i = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    x = (i + 0.5 - 0.5 * w) / focal
    z = -(j + 0.5 - 0.5 * h) / focal
    d = torch.stack([x, torch.ones_like(x), z], -1)
    d = d / d.norm(dim=-1, keepdim=True)
    o = torch.tensor(CAMERA_CENTER).expand_as(d)
    return o.reshape(-1, 3).contiguous().to(device), d.reshape(-1, 3).contiguous().to(device)


def fibonacci_hemisphere(normals, sample_num, random_rotate, gen=None):
    """utils/graphics_utils.py:19-47.  normals[B,3] -> dirs[B,S,3]."""
    dev = normals.device
    delta = math.pi * (3.0 - math.sqrt(5.0))
    idx = torch.arange(sample_num, dtype=torch.float32, device=dev)[None]
    z = (1 - 2 * idx / (2 * sample_num - 1)).clamp_min(math.sin(10 / 180 * math.pi))
    rad = torch.sqrt(1 - z ** 2)
    theta = delta * idx
    if random_rotate:
        rnd = torch.rand(normals.shape[0], 1, generator=gen) if gen is not None else torch.rand(normals.shape[0], 1)
        theta = rnd.to(dev) * 2 * math.pi + theta
    y = torch.cos(theta) * rad
    x = torch.sin(theta) * rad
    zs = torch.stack([x, y, z.expand_as(y)], dim=-2)  # [B,3,S]
    dirs = _rot_from_z(normals) @ zs
    dirs = dirs / dirs.norm(dim=-2, keepdim=True)
    return dirs.transpose(-1, -2).contiguous()


def secondary_rays(points, normals, sample_num, random_rotate=True, seed=RAY_SEED, t_min=LIGHT_T_MIN):
    """gaussian_renderer/__init__.py:382: origins = x + dir * light_t_min.  Returns ([B,S,3], [B,S,3])."""
    gen = torch.Generator().manual_seed(seed)
    dirs = fibonacci_hemisphere(normals, sample_num, random_rotate, gen)
    return (points[:, None] + dirs * t_min).contiguous(), dirs


def shading_points_from_primary(o, d, depth, alpha, normal, seed=RAY_SEED):
    """Shading points / normals for every pixel from a primary trace.  Pixels whose primary ray missed the object
    are re-assigned the shading point of a randomly chosen hit pixel, so that every secondary bundle starts on the
    surface (the hard case); see DESIGN.md 'workload'."""
    hit = alpha > 0.5
    t = depth / alpha.clamp_min(1e-6)
    pts = o + d * t[:, None]
    nrm = normal / normal.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    idx_hit = torch.nonzero(hit)[:, 0]
    if idx_hit.numel() == 0:
        raise RuntimeError("primary pass hit nothing")
    gen = torch.Generator().manual_seed(seed + 1)
    pick = torch.randint(0, idx_hit.numel(), (o.shape[0],), generator=gen).to(o.device)
    src = torch.where(hit, torch.arange(o.shape[0], device=o.device), idx_hit[pick])
    return pts[src].contiguous(), nrm[src].contiguous()
