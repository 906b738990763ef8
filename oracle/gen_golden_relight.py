"""Golden vectors of the RELIGHT branch of the reference's rendering_equation (gaussian_renderer/__init__.py:362-381), recorded
by EXECUTING THE UNMODIFIED SOURCE of rendering_equation / GGX_specular / sample_incident_rays / fibonacci_sphere_sampling on
the CPU, cut out of the reference files with `ast` exactly as oracle/gen_golden_shading.py does.  Supplied from outside:
  * `pc.get_envmap`  -- oracle.shading.RelightEnvStandIn: the reference's EnvLight answers mode 'diffuse' / 'specular' from
    cube-map mips built by nvdiffrec + nvdiffrast (absent here, out of scope); the stand-in answers them from lat-long maps.
    The test hands the SAME object to the product, so what is pinned is everything around the two lookups;
  * `dr.texture`     -- 'linear' / 'wrap' (pure_env) and 'linear' / 'clamp' (FG table) restated in oracle/shading.py
    ("parity unpinned" pieces, nvdiffrast is not in this image);
  * `pc.trace`       -- returns recorded raw tracer outputs (normal, feature, alpha) passed through the normalisation of
    scene/gaussian_model.py:751-756, and logs the rays it was asked to trace.

    python oracle/gen_golden_relight.py      ->  tests/golden/ref_relight.npz
"""
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen_golden_incident as ggi                 # noqa: E402
from oracle.gen_golden_shading import REF, T_MIN, cut         # noqa: E402
from oracle.shading import RelightEnvStandIn, texture_linear_clamp, texture_linear_wrap, update_pdf   # noqa: E402


class _Dr:
    @staticmethod
    def texture(tex, uv, filter_mode="auto", boundary_mode="wrap"):
        assert filter_mode == "linear" and tex.shape[0] == 1
        if boundary_mode == "clamp":
            return texture_linear_clamp(tex[0], uv)
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
    """What the relight branch needs of GaussianModel: get_envmap, get_base_color, get_rough, FG_LUT, trace(features=...)."""

    def __init__(self, env, fg_lut, n_surf, normal_raw, feature_raw, alpha_raw, g):
        self.get_envmap, self.FG_LUT = env, fg_lut
        self.get_base_color, self.get_rough = torch.rand(n_surf, 3, generator=g), torch.rand(n_surf, 1, generator=g)
        self.raw = (normal_raw, feature_raw, alpha_raw)
        self.rays = self.features = None

    def trace(self, rays_o, rays_d, features=None, camera_center=None, back_culling=False):
        self.rays, self.features = (rays_o.detach().clone(), rays_d.detach().clone()), features
        normal, feature, alpha = self.raw
        alpha_ = alpha[..., None]                                                              # gaussian_model.py:751-756
        normal = torch.where(alpha_ < 1 - T_MIN, normal, normal / alpha_)
        feature = torch.where(alpha_ < 1 - T_MIN, feature, feature / alpha_)
        alpha = torch.where(alpha < 1 - T_MIN, alpha, torch.ones_like(alpha))
        return {"normal": normal, "feature": feature, "alpha": alpha}


def make_case(ref, seed, P, S, n_light, activation, res, with_transform, wo_indirect_relight, f0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)          # noqa: E731
    n = torch.randn(P, 3, generator=g)
    n = n / n.norm(dim=-1, keepdim=True)
    n[0] = torch.tensor([0.0, 0.0, 1.0])
    n[1] = torch.tensor([0.0, 0.0, -1.0])
    view = torch.randn(P, 3, generator=g)
    T = S + n_light
    hit = r(P, T) < 0.7
    alpha_raw = (r(P, T) * 1.15).clamp(0, 0.999) * hit
    alpha_raw[:, 0] = 0.985                                                   # saturated: normalised by pc.trace
    nrm = torch.randn(P, T, 3, generator=g)
    nrm = nrm / nrm.norm(dim=-1, keepdim=True)
    inp = {
        "base_color": r(P, 3), "roughness": 0.05 + 0.9 * r(P, 1), "normals": n, "viewdirs": view,
        "position": torch.randn(P, 3, generator=g),
        "normal_raw": nrm * alpha_raw[..., None] * (0.6 + 0.4 * r(P, T, 1)),   # composited normals: shorter than alpha
        "feature_raw": torch.cat([r(P, T, 3), 0.02 + 0.96 * r(P, T, 1)], -1) * alpha_raw[..., None],
        "alpha_raw": alpha_raw,
    }
    maps = {k: torch.randn(res[0], res[1], 3, generator=g) * 0.7 for k in ("base", "base_diffuse", "base_spec0", "base_spec1")}
    transform = None
    if with_transform:
        a = 0.4
        transform = torch.tensor([[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]])
    env = RelightEnvStandIn(maps["base"], maps["base_diffuse"], maps["base_spec0"], maps["base_spec1"], activation, transform)
    light_log = []
    if n_light > 0:
        env._pdf = update_pdf(env.base.data, activation)
        helper = ref.EnvLight(path=None, device="cpu", resolution=list(res), activation=activation, init_value=0.5)
        helper.base.data = maps["base"].clone()
        if transform is not None:
            helper.set_transform(transform)
        helper.update_pdf()
        assert torch.equal(helper._pdf, env._pdf)

        def sample(B, num, training=False):           # the reference's own sampler and density, logged
            out = helper.sample_light_directions(B, num, training)
            light_log.append(tuple(t.detach().clone() for t in out))
            return out
        env.sample_light_directions, env.light_pdf = sample, helper.light_pdf
    fg_lut = (0.05 + 0.9 * torch.rand(1, 24, 32, 2, generator=g))
    pc = _PC(env, fg_lut, 50, inp["normal_raw"], inp["feature_raw"], inp["alpha_raw"], g)
    pipe = types.SimpleNamespace(diffuse_sample_num=S, light_sample_num=n_light, light_t_min=0.05, wo_indirect=False,
                                 detach_indirect=False, wo_indirect_relight=wo_indirect_relight)
    torch.manual_seed(seed + 1)
    with torch.no_grad():
        out = ref.rendering_equation(inp["base_color"], inp["roughness"], inp["normals"], inp["position"], inp["viewdirs"], pc,
                                     pipe, training=False, f0=f0, relight=True)
    assert pc.features.shape == (50, 4)
    rec = {f"in_{k}": v.numpy() for k, v in inp.items()}
    rec.update({f"in_{k}": v.numpy() for k, v in maps.items()})
    rec["in_fg_lut"] = fg_lut.numpy()
    rec["activation"], rec["S"], rec["n_light"] = np.array(activation), np.array(S), np.array(n_light)
    rec["f0"], rec["wo_indirect_relight"] = np.array(f0, np.float32), np.array(wo_indirect_relight)
    if transform is not None:
        rec["in_transform"] = transform.numpy()
    if n_light > 0:
        rec["in_light_dirs"], rec["in_pdf"] = light_log[0][0].numpy(), env._pdf.numpy()
    rec["rays_o"], rec["rays_d"] = pc.rays[0].numpy(), pc.rays[1].numpy()
    for k in sorted(out):
        rec[f"out_{k}"] = out[k].numpy()
    return rec


def main():
    ref = load_reference()
    cases = {
        "eval32": make_case(ref, 21, 30, 32, 0, "exp", (16, 32), False, False, 0.04),
        "eval24_light12_xf": make_case(ref, 22, 20, 24, 12, "exp", (8, 16), True, False, 0.02),
        "eval16_sigmoid_wo": make_case(ref, 23, 12, 16, 0, "sigmoid", (8, 16), False, True, 0.04),
    }
    flat = {f"{c}/{k}": v for c, rec in cases.items() for k, v in rec.items()}
    path = os.path.join(ROOT, "tests", "golden", "ref_relight.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, os.path.getsize(path), "bytes;", {c: sorted(k for k in rec if k.startswith("out_")) for c, rec in cases.items()})


if __name__ == "__main__":
    main()
