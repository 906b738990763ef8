"""Primary pass through the tracer: camera rays generated in the kernels against the reference's Camera
arithmetic (scene/cameras.py:87-100, restated below), the traced G-buffer against the CPU oracle on those rays, gradients
against the materialised-ray path."""
import math

import numpy as np
import pytest
import torch

import oracle
from irgs_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = ("means3D", "opacity", "ru", "rv", "normals", "features", "shs")


def _reference_camera_rays(world_view_transform, fovx, fovy, W, H):
    """scene/cameras.py:87-100 verbatim in its arithmetic (float32 torch on the CPU)."""
    v, u = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    focal_x = W / (2 * np.tan(fovx * 0.5))
    focal_y = H / (2 * np.tan(fovy * 0.5))
    rays_d_camera = torch.stack([(u - W / 2 + 0.5) / focal_x, (v - H / 2 + 0.5) / focal_y, torch.ones_like(u)], dim=-1).reshape(-1, 3)
    rays_d = rays_d_camera.float() @ world_view_transform[:3, :3].T
    return torch.nn.functional.normalize(rays_d, dim=-1), rays_d


def _camera(W=72, H=56, fovx=0.7):
    from irgs_b200.primary import Camera
    cam = Camera.look_at(synth.CAMERA_CENTER, (0.05, 0.0, -0.1), (0.0, 0.0, -1.0), fovx, W, H)
    return cam


def test_generated_camera_rays_match_the_reference_camera():
    from irgs_b200.primary import Camera
    W, H, fovx = 97, 61, 0.8
    fovy = 2 * math.atan(math.tan(fovx / 2) * H / W)
    g = torch.Generator().manual_seed(3)
    q = torch.randn(4, generator=g)
    R = synth.quat_to_rot(q[None])[0]                      # an arbitrary rotation: world_view_transform[:3,:3] of the reference
    wvt = torch.eye(4)
    wvt[:3, :3] = R
    centre = torch.tensor([0.3, -2.0, 0.7])

    class RefCam:
        image_width, image_height, FoVx, FoVy = W, H, fovx, fovy
        world_view_transform, camera_center = wvt, centre
    cam = Camera.from_reference(RefCam)
    o, d = cam.rays(DEV)
    want, _ = _reference_camera_rays(wvt, fovx, fovy, W, H)
    assert np.abs(d.cpu().numpy() - want.numpy()).max() <= 3e-7
    assert torch.equal(o.cpu(), centre[None].expand(W * H, 3))


def test_camera_trace_equals_trace_of_the_generated_rays_oracle_and_gradients(small_scene):
    from irgs_b200.primary import render_primary, trace_camera
    from irgs_b200.raytracer import GaussianTracer
    sc, inp = small_scene
    g = {k: v.to(DEV) for k, v in inp.items()}
    tr = GaussianTracer(transmittance_min=synth.T_MIN)
    tr.build_from_surfels(g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], synth.ALPHA_MIN)
    cam = _camera()
    H, W = cam.height, cam.width
    o, d = cam.rays(DEV)
    args = [g[k] for k in KEYS]
    with torch.no_grad():
        fused = trace_camera(tr, cam, *args, synth.ALPHA_MIN)
        hc = tr.last_hit_count.clone()
        plain = tr.trace(o, d, *args, synth.ALPHA_MIN)
    for a, b in zip(fused, plain):
        assert torch.equal(a.reshape(b.shape), b)                  # the very same rays: bit-identical
    assert torch.equal(hc.reshape(-1), tr.last_hit_count) and float((hc > 0).float().mean()) > 0.2
    S = oracle.Scene(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], inp["shs"], inp["features"])
    ref = oracle.trace_forward(S, o.cpu(), d.cpu())
    safe = (ref["margin"][:, 0] > 2e-5) & (ref["margin"][:, 1] > 2e-5)
    assert safe.mean() > 0.97 and np.array_equal(hc.cpu().numpy().reshape(-1)[safe], ref["hit_count"][safe])
    for name, t in zip(("color", "normal", "feature", "depth", "alpha"), fused):
        assert np.abs(t.cpu().numpy().reshape(ref[name].shape) - ref[name])[safe].max() <= 1e-4, name
    # gradients: generated rays vs the same rays materialised
    gen = torch.Generator().manual_seed(5)
    w = [torch.randn(H * W, c, generator=gen).to(DEV) for c in (3, 3, inp["features"].shape[1])] + \
        [torch.randn(H * W, generator=gen).to(DEV) for _ in range(2)]
    res = []
    for use_cam in (True, False):
        leaf = {k: g[k].clone().requires_grad_(True) for k in KEYS}
        outs = trace_camera(tr, cam, *[leaf[k] for k in KEYS], synth.ALPHA_MIN) if use_cam else \
            tr.trace(o, d, *[leaf[k] for k in KEYS], synth.ALPHA_MIN)
        sum((t.reshape(wi.shape) * wi).sum() for t, wi in zip(outs, w)).backward()
        res.append({k: leaf[k].grad for k in KEYS})
    for k in KEYS:
        a, b = res[0][k], res[1][k]
        assert bool(b.any()) and float((a - b).abs().max()) <= 2e-4 * float(b.abs().max()), k
    # the G-buffer dict
    feats4 = torch.cat([g["features"][:, :3], g["features"][:, 3:4].clamp(0.02, 1)], 1)
    gb = render_primary(tr, cam, (g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], feats4, g["shs"]), synth.ALPHA_MIN)
    assert gb["rend_alpha"].shape == (1, H, W) and gb["rend_normal"].shape == (3, H, W) and gb["base_color"].shape == (3, H, W)
    assert gb["roughness"].shape == (1, H, W) and gb["render"].shape == (3, H, W) and gb["points"].shape == (H, W, 3)
    solid = gb["rend_alpha"][0] > 0.9
    assert bool(solid.any())
    # the surface points lie along the pixel's ray at the expected depth, in front of the camera and inside the scene bounds
    assert float(gb["points"][solid].abs().max()) < 1.5 and float(gb["surf_depth"][0][solid].min()) > 0.5
    # the shading normal faces the camera
    assert float((gb["normal_map"][solid] * gb["rays_d_hw"][solid]).sum(-1).max()) < 0.2
