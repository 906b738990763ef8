"""Records golden vectors of the reference's shading around the incident-ray trace by EXECUTING THE UNMODIFIED SOURCE of

    gaussian_renderer/__init__.py : sample_incident_rays, rendering_equation, GGX_specular
    scene/light.py                : inverse_sigmoid, pixel_grid, EnvLight   (EnvLight.__call__(mode='pure_env'), update_pdf,
                                    sample_light_directions, light_pdf -- the light_sample_num > 0 branch)
    utils/graphics_utils.py       : fibonacci_sphere_sampling, rotation_between_z

on the CPU of the build container.  The functions are cut out of the reference files with `ast` (the modules themselves
import packages that are not in this image: diff_surfel_rasterization, nvdiffrast, kornia, trimesh, pyexr ...) and run
as they are; nothing is edited.  Three things are supplied from outside, as the reference's own callers would:
  * `dr.texture` -- nvdiffrast is absent, so its 'linear' / 'wrap' lookup is the restatement oracle/shading.py
    texture_linear_wrap (the one "parity unpinned" piece, see that file);
  * `pc.trace` -- returns recorded raw tracer outputs (random colour / alpha leaves) passed through the normalisation of
    scene/gaussian_model.py:748-752, so the golden gradients include that step;
  * torch factory calls with device='cuda' lose that argument (same shim as gen_golden_incident.py).
Outputs AND torch-autograd gradients (of a fixed random linear functional of the outputs) are recorded.

    python oracle/gen_golden_shading.py      ->  tests/golden/ref_shading.npz
"""
import ast
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen_golden_incident as ggi            # noqa: E402  (device shim + graphics_utils loader)
from oracle.shading import texture_linear_wrap           # noqa: E402

REF = "/root/reference"
T_MIN = 0.03   # GaussianTracer.transmittance_min as IRGS constructs it (scene/gaussian_model.py:118-119)


def cut(path, names):
    """Source text of the named top-level functions / classes of a reference file, unmodified."""
    src = open(path).read()
    tree = ast.parse(src)
    out = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            out.append(ast.get_source_segment(src, node))
    assert len(out) == len(names), (path, names)
    return "\n\n".join(out)


class _Dr:
    @staticmethod
    def texture(tex, uv, filter_mode="auto", boundary_mode="wrap"):
        assert filter_mode == "linear" and boundary_mode == "wrap" and tex.shape[0] == 1
        return texture_linear_wrap(tex[0], uv)


def load_reference():
    gu = ggi.load_reference()
    torch.full = ggi._strip_device(torch.full)
    ns = {"torch": torch, "F": F, "np": np, "math": math, "dr": _Dr,
          "fibonacci_sphere_sampling": gu.fibonacci_sphere_sampling, "rotation_between_z": gu.rotation_between_z}
    exec(cut(f"{REF}/gaussian_renderer/__init__.py", ["sample_incident_rays", "rendering_equation", "GGX_specular"]), ns)
    exec(cut(f"{REF}/scene/light.py", ["inverse_sigmoid", "pixel_grid", "EnvLight"]), ns)
    return types.SimpleNamespace(**ns)


class _PC:
    """What rendering_equation needs of GaussianModel: .get_envmap and .trace."""

    def __init__(self, env, color_raw, alpha_raw):
        self.get_envmap, self.color_raw, self.alpha_raw = env, color_raw, alpha_raw
        self.rays = None

    def trace(self, rays_o, rays_d, camera_center=None):
        self.rays = (rays_o.detach().clone(), rays_d.detach().clone())
        color, alpha = self.color_raw, self.alpha_raw
        alpha_ = alpha[..., None]                                                      # gaussian_model.py:748-752
        color = torch.where(alpha_ < 1 - T_MIN, color, color / alpha_)
        alpha = torch.where(alpha < 1 - T_MIN, alpha, torch.ones_like(alpha))
        return {"color": color, "alpha": alpha}


def make_case(ref, seed, P, S, training, activation, res, with_transform, n_light=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)          # noqa: E731
    n = torch.randn(P, 3, generator=g)
    n = n / n.norm(dim=-1, keepdim=True)
    n[0] = torch.tensor([0.0, 0.0, 1.0])
    n[1] = torch.tensor([0.0, 0.0, -1.0])               # the -identity branch of rotation_between_z
    view = torch.randn(P, 3, generator=g)
    view = torch.where(((view * n).sum(-1, keepdim=True) < 0) & (r(P, 1) < 0.8), -view, view)   # mostly front-facing
    view = view * (0.5 + r(P, 1))                        # not unit length: GGX normalises
    inp = {
        "base_color": r(P, 3), "roughness": 0.05 + 0.9 * r(P, 1), "normals": n, "viewdirs": view,
        "position": torch.randn(P, 3, generator=g),
        "color_raw": r(P, S + n_light, 3) * (r(P, S + n_light, 1) < 0.6),
        "alpha_raw": (r(P, S + n_light) * 1.15).clamp(0, 0.999) * (r(P, S + n_light) < 0.7),
    }
    inp["alpha_raw"][:, 0] = 0.985                       # saturated rays (>= 1 - T_MIN): normalised by pc.trace
    inp["color_raw"][:, 0] = r(P, 3)
    env = ref.EnvLight(path=None, device="cpu", resolution=list(res), activation=activation, init_value=0.5)
    base = torch.randn(res[0], res[1], 3, generator=g) * (0.8 if activation != "none" else 0.5)
    if activation == "none":
        base = base + 0.3                                # some texels negative: clamp_min(0) is exercised
    env.base.data = base.clone()
    transform = None
    if with_transform:
        a = 0.7
        transform = torch.tensor([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]])
        env.set_transform(transform)
    leaves = {k: v.clone().requires_grad_(True) for k, v in inp.items() if k != "position"}
    pc = _PC(env, leaves["color_raw"], leaves["alpha_raw"])
    pipe = types.SimpleNamespace(diffuse_sample_num=S, light_sample_num=n_light, light_t_min=0.05, wo_indirect=False,
                                 detach_indirect=False)
    torch.manual_seed(seed + 1)
    light_log = []
    if n_light > 0:
        env.update_pdf()                                  # render.py:89 does this before rendering with light samples
        sample = env.sample_light_directions

        def logged(*a, **k):
            out = sample(*a, **k)
            light_log.append(tuple(t.detach().clone() for t in out))
            return out
        env.sample_light_directions = logged
    ggi._rand_log.clear()
    out = ref.rendering_equation(leaves["base_color"], leaves["roughness"], leaves["normals"], inp["position"],
                                 leaves["viewdirs"], pc, pipe, training=training)
    azimuth = (ggi._rand_log[0] * 2 * np.pi).reshape(-1) if training else None      # the first rand call: the Fibonacci rotation
    keys = sorted(out)
    w = {k: torch.randn(out[k].shape, generator=g) for k in keys}
    sum((out[k] * w[k]).sum() for k in keys).backward()
    rec = {f"in_{k}": v.numpy() for k, v in inp.items()}
    rec["in_env_base"] = base.numpy()
    rec["activation"] = np.array(activation)
    rec["training"] = np.array(training)
    rec["S"] = np.array(S)
    rec["n_light"] = np.array(n_light)
    if n_light > 0:
        rec["in_light_dirs"], rec["light_pdfs"] = light_log[0][0].numpy(), light_log[0][1].numpy()
        rec["in_pdf"] = env._pdf.numpy()
    if transform is not None:
        rec["in_transform"] = transform.numpy()
    if azimuth is not None:
        rec["in_azimuth"] = azimuth.numpy()
    rec["rays_o"], rec["rays_d"] = pc.rays[0].numpy(), pc.rays[1].numpy()
    for k in keys:
        rec[f"out_{k}"] = out[k].detach().numpy()
        rec[f"w_{k}"] = w[k].numpy()
    for k, v in leaves.items():
        rec[f"grad_{k}"] = v.grad.numpy()
    rec["grad_env_base"] = env.base.grad.numpy()
    return rec


def main():
    ref = load_reference()
    cases = {
        "eval24": make_case(ref, 11, 40, 24, False, "exp", (16, 32), False),
        "train64": make_case(ref, 12, 33, 64, True, "exp", (32, 64), False),
        "train40_sigmoid_xf": make_case(ref, 13, 24, 40, True, "sigmoid", (8, 16), True),
        "eval33_none": make_case(ref, 14, 16, 33, False, "none", (16, 16), False),
        "eval24_light12": make_case(ref, 15, 24, 24, False, "exp", (16, 32), False, n_light=12),
        "train32_light16_xf": make_case(ref, 16, 20, 32, True, "exp", (8, 16), True, n_light=16),
    }
    flat = {f"{c}/{k}": v for c, rec in cases.items() for k, v in rec.items()}
    path = os.path.join(ROOT, "tests", "golden", "ref_shading.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, os.path.getsize(path), "bytes;", {c: sorted(k for k in rec if k.startswith("out_")) for c, rec in cases.items()})


if __name__ == "__main__":
    main()
