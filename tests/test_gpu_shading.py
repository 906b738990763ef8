"""GPU parity tests of the shading epilogue (SURVEY.md 8f rank 1 + 3): the CUDA kernels of irgs_b200/csrc/shade.cu through
the C ABI against (1) golden vectors recorded from the UNMODIFIED reference functions (outputs and autograd gradients),
(2) the torch restatement oracle/shading.py composed with the un-fused tracer, end to end, and (3) size-independent
properties at a full-size chunk."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import shading as osh
from irgs_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_shading.npz")
CASES = ("eval24", "train64", "train40_sigmoid_xf", "eval33_none", "eval24_light12", "train32_light16_xf")
T_MIN = 0.03


def load_case(name):
    z = np.load(GOLDEN)
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def _close(a, b, what, tol=2e-4, cos_min=0.999999):
    a = a.detach().double().cpu().numpy().ravel() if torch.is_tensor(a) else np.asarray(a, np.float64).ravel()
    b = b.detach().double().cpu().numpy().ravel() if torch.is_tensor(b) else np.asarray(b, np.float64).ravel()
    assert np.isfinite(a).all(), what
    a, b = a[np.isfinite(b)], b[np.isfinite(b)]      # the reference's 0/0 gradients of rays with alpha == 0, see test_shading_cpu
    cos = a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300)
    assert cos > cos_min and np.abs(a - b).max() <= tol * np.abs(b).max(), (what, cos, np.abs(a - b).max(), np.abs(b).max())


@pytest.mark.parametrize("name", CASES)
def test_shade_kernels_match_reference_golden(name):
    from irgs_b200 import shading
    case = load_case(name)
    t = lambda k: torch.from_numpy(case[k]).to(DEV)                       # noqa: E731
    leaves = {k: t("in_" + k).requires_grad_(True) for k in ("base_color", "roughness", "normals", "viewdirs", "color_raw",
                                                            "alpha_raw", "env_base")}
    az = t("in_azimuth") if "in_azimuth" in case else None
    tr = t("in_transform") if "in_transform" in case else None
    S, Sl = int(case["S"]), int(case["n_light"]) if "n_light" in case else 0
    common = dict(activation=str(case["activation"]), transform=tr, transmittance_min=T_MIN)
    if Sl == 0:
        out = shading.shade_incident(leaves["normals"], S, leaves["base_color"], leaves["roughness"], leaves["viewdirs"],
                                     leaves["env_base"], leaves["color_raw"], leaves["alpha_raw"], azimuth=az, **common)
    else:   # light_sample_num > 0: one call per kind of sample, the rows add up
        mix = dict(pdf=t("in_pdf"), p_diffuse=S / (S + Sl), p_light=Sl / (S + Sl), total_samples=S + Sl)
        od = shading.shade_incident(leaves["normals"], S, leaves["base_color"], leaves["roughness"], leaves["viewdirs"],
                                    leaves["env_base"], leaves["color_raw"][:, :S], leaves["alpha_raw"][:, :S], azimuth=az,
                                    **common, **mix)
        ol = shading.shade_incident(leaves["normals"], Sl, leaves["base_color"], leaves["roughness"], leaves["viewdirs"],
                                    leaves["env_base"], leaves["color_raw"][:, S:], leaves["alpha_raw"][:, S:],
                                    dirs=t("in_light_dirs"), **common, **mix)
        out = {k: od[k] + ol[k] for k in od}
    keys = [k for k in shading.OUT_SLICES if f"out_{k}" in case]
    assert len(keys) == (3 if bool(case["training"]) else 6)
    for k in keys:
        ref = case[f"out_{k}"]
        assert out[k].shape == ref.shape
        assert np.abs(out[k].detach().cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), k
    sum((out[k] * t(f"w_{k}")).sum() for k in keys).backward()
    for k in ("base_color", "roughness", "viewdirs", "color_raw", "alpha_raw", "env_base"):
        _close(leaves[k].grad, case[f"grad_{k}"], k)
    _close(leaves["normals"].grad, case["grad_normals"], "normals", tol=5e-4)


def test_env_lookup_forward_backward_against_the_oracle():
    from irgs_b200 import shading
    g = torch.Generator().manual_seed(8)
    d = torch.randn(20000, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    w = torch.randn(20000, 3, generator=g)
    a = 0.4
    xf = torch.tensor([[1.0, 0.0, 0.0], [0.0, math.cos(a), -math.sin(a)], [0.0, math.sin(a), math.cos(a)]])
    for act, res, transform in (("exp", (16, 32), None), ("sigmoid", (7, 9), xf), ("none", (64, 128), None)):
        base = torch.randn(res[0], res[1], 3, generator=g) * 0.7
        bo, do = base.clone().double().requires_grad_(True), d.clone().double().requires_grad_(True)
        ref = osh.env_pure(bo, do, act, transform.double() if transform is not None else None)
        (ref * w.double()).sum().backward()
        bg, dg = base.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
        env = shading.EnvLight(resolution=res, activation=act, device=DEV)
        env.base.data = bg.data
        env.set_transform(transform)
        out = env(dg.view(100, 200, 3))
        assert out.shape == (100, 200, 3)
        (out * w.to(DEV).view(100, 200, 3)).sum().backward()
        # float32 atan2 / acos against float64: a texel-boundary crossing can move single samples, so medians + tails
        ref = ref.detach()
        err = (out.detach().cpu().view(-1, 3).double() - ref).abs()
        assert err.median() <= 1e-6 and (err > 1e-3 * max(1.0, float(ref.abs().max()))).float().mean() < 2e-3
        _close(env.base.grad, bo.grad, f"texels {act}", tol=2e-3, cos_min=0.99999)
        gd, gr = dg.grad.cpu().double(), do.grad
        bad = ((gd - gr).abs().amax(-1) > 1e-3 * gr.abs().max()).float().mean()
        assert bad < 2e-3, (act, float(bad))
    # empty input, wrong dtype
    env = shading.EnvLight(resolution=(4, 8), device=DEV)
    assert env(torch.zeros(0, 3, device=DEV)).shape == (0, 3)
    with pytest.raises(NotImplementedError):
        env(torch.zeros(4, 3, device=DEV), mode="specular")


def _torch_dirs(normals, S, azim):
    from irgs_b200.incident import rotation_between_z
    idx = torch.arange(S, dtype=torch.float32, device=normals.device)[None]
    z = (1 - 2 * idx / (2 * S - 1)).clamp_min(math.sin(10 / 180 * math.pi))
    rad = torch.sqrt(1 - z ** 2)
    theta = math.pi * (3.0 - math.sqrt(5.0)) * idx
    if azim is not None:
        theta = azim[:, None] + theta
    P = normals.shape[0]
    zs = torch.stack([(torch.sin(theta) * rad).expand(P, S), (torch.cos(theta) * rad).expand(P, S), z.expand(P, S)], -2)
    return torch.nn.functional.normalize(rotation_between_z(normals) @ zs, dim=-2).transpose(-1, -2)


@pytest.mark.parametrize("training", [False, True])
def test_rendering_equation_end_to_end_against_the_unfused_composition(small_scene, training):
    """rendering_equation (fused ray generation + tracer + shading kernels) against: torch directions -> tracer.trace on
    materialised rays -> the oracle's torch restatement of the reference's rendering_equation (run on the GPU tensors)."""
    from irgs_b200 import shading
    from irgs_b200.raytracer import GaussianTracer
    sc, inp = small_scene
    g = {k: v.to(DEV) for k, v in inp.items()}
    tr = GaussianTracer(transmittance_min=synth.T_MIN)
    tr.build_from_surfels(g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], synth.ALPHA_MIN)
    gen = torch.Generator().manual_seed(17)
    P, S = 96, 48
    idx = torch.randint(0, inp["means3D"].shape[0], (P,), generator=gen)
    nrm = inp["normals"][idx].contiguous()
    pos = (inp["means3D"][idx] + 0.01 * nrm).contiguous()
    view = torch.nn.functional.normalize(torch.tensor(synth.CAMERA_CENTER, dtype=torch.float32)[None] - pos, dim=-1)
    azim = (torch.rand(P, generator=gen) * 2 * math.pi).to(DEV) if training else None
    base_color, rough = torch.rand(P, 3, generator=gen), 0.1 + 0.8 * torch.rand(P, 1, generator=gen)
    env_base = torch.randn(16, 32, 3, generator=gen) * 0.5
    keys_s = ("means3D", "opacity", "ru", "rv", "normals", "features", "shs")
    keys = ("diffuse", "specular", "light_direct") if training else tuple(shading.OUT_SLICES)
    w = {k: torch.randn(P, 1 if k == "visibility" else 3, generator=gen).to(DEV) for k in keys}
    res = []
    for fused in (True, False):
        leaf = {k: g[k].clone().requires_grad_(True) for k in keys_s}
        pt = {"position": pos.to(DEV).requires_grad_(True), "normals_pt": nrm.to(DEV).requires_grad_(True),
              "viewdirs": view.to(DEV).requires_grad_(True), "base_color": base_color.to(DEV).requires_grad_(True),
              "roughness": rough.to(DEV).requires_grad_(True)}
        env = shading.EnvLight(resolution=(16, 32), activation="exp", device=DEV)
        env.base.data = env_base.to(DEV)
        surf = tuple(leaf[k] for k in keys_s)
        if fused:
            out = shading.rendering_equation(pt["base_color"], pt["roughness"], pt["normals_pt"], pt["position"], pt["viewdirs"],
                                             tr, surf, env, S, training=training, azimuth=azim, light_t_min=0.05,
                                             alpha_min=synth.ALPHA_MIN)
            assert set(out) == set(keys)
        else:
            d = _torch_dirs(pt["normals_pt"], S, azim)
            color, _, _, _, alpha = tr.trace(pt["position"][:, None] + d * 0.05, d, *surf, synth.ALPHA_MIN)
            out = osh.rendering_equation(pt["base_color"], pt["roughness"], pt["normals_pt"], pt["viewdirs"], d, color, alpha,
                                         env.base, "exp", None, synth.T_MIN)
        sum((out[k] * w[k]).sum() for k in keys).backward()
        res.append(({k: out[k].detach() for k in keys},
                    {**{k: v.grad for k, v in pt.items()}, **{k: leaf[k].grad for k in keys_s}, "env_base": env.base.grad}))
    for k in keys:
        a, b = res[0][0][k], res[1][0][k]
        assert float((a - b).abs().max()) <= 1e-4 * max(1.0, float(b.abs().max())), k
    for k in res[0][1]:
        a, b = res[0][1][k], res[1][1][k]
        assert a is not None and b is not None, k
        bf = torch.nan_to_num(b.double().flatten(), 0.0, 0.0, 0.0)     # the un-fused torch path back-propagates 0/0 at alpha == 0
        af = a.double().flatten()
        assert torch.isfinite(af).all(), k
        if float(bf.norm()) == 0.0:                 # e.g. `features`: the rendering equation does not read the feature output
            assert float(af.norm()) == 0.0, k
            continue
        cos = float(af @ bf / (af.norm() * bf.norm() + 1e-300))
        # ulp-level direction differences move individual threshold decisions of the tracer: cosine is the criterion
        assert cos >= 0.9999, (k, cos)


def test_flags_empty_input_and_errors(small_scene):
    from irgs_b200 import shading
    from irgs_b200.raytracer import GaussianTracer
    sc, inp = small_scene
    g = {k: v.to(DEV) for k, v in inp.items()}
    tr = GaussianTracer(transmittance_min=synth.T_MIN)
    tr.build_from_surfels(g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], synth.ALPHA_MIN)
    surf = tuple(g[k] for k in ("means3D", "opacity", "ru", "rv", "normals", "features", "shs"))
    env = shading.EnvLight(resolution=(8, 16), device=DEV)
    P, S = 20, 16
    pos, nrm = g["means3D"][:P] + 0.01 * g["normals"][:P], g["normals"][:P].contiguous()
    view, bc, ro = -nrm, torch.rand(P, 3, device=DEV), torch.rand(P, 1, device=DEV)
    args = (bc, ro, nrm, pos, view, tr, surf, env, S)
    full = shading.rendering_equation(*args)
    wo = shading.rendering_equation(*args, wo_indirect=True)
    assert float(wo["light_indirect"].detach().abs().max()) == 0.0 and torch.equal(wo["visibility"], full["visibility"])
    assert torch.equal(wo["light_direct"], full["light_direct"])
    # a constant environment of radiance 0.5 (the initial value): light_direct is exactly that
    assert float((full["light_direct"].detach() - 0.5).abs().max()) <= 1e-6
    det = shading.rendering_equation(bc.clone().requires_grad_(True), ro, nrm, pos.clone().requires_grad_(True), view, tr,
                                     tuple(t.clone().requires_grad_(True) for t in surf), env, S, detach_indirect=True)
    assert torch.equal(det["diffuse"], full["diffuse"])
    e = shading.rendering_equation(bc[:0], ro[:0], nrm[:0], pos[:0], view[:0], tr, surf, env, S, training=True)
    assert e["diffuse"].shape == (0, 3) and set(e) == {"diffuse", "specular", "light_direct"}
    with pytest.raises(ValueError):                     # the relight branch needs the FG table and [N,4] features
        shading.rendering_equation(*args, relight=True)
    with pytest.raises(NotImplementedError):            # ... and is forward-only
        shading.rendering_equation(*args, relight=True, training=True, fg_lut=torch.zeros(1, 4, 4, 2, device=DEV))
    with pytest.raises(ValueError):
        shading.shade_incident(nrm, S, bc, ro, view, env.base, torch.zeros(P, S + 1, 3, device=DEV), torch.zeros(P, S, device=DEV))
    with pytest.raises(TypeError):
        shading.shade_incident(nrm, S, bc.double(), ro, view, env.base, torch.zeros(P, S, 3, device=DEV),
                               torch.zeros(P, S, device=DEV))


def test_full_size_chunk_properties():
    """One chunk of BASELINE's C3 shape (65 536 points x 256 samples = 2^24 rays): quantities with a closed form --
    visibility = 1 - mean(normalised alpha), light_indirect = mean(normalised colour), light = visibility-weighted sum --
    determinism, and linearity of the backward in the incoming gradient."""
    from irgs_b200 import shading
    P, S = 1 << 16, 256
    gen = torch.Generator(DEV).manual_seed(3)
    r = lambda *s: torch.rand(*s, device=DEV, generator=gen)                  # noqa: E731
    nrm = torch.nn.functional.normalize(torch.randn(P, 3, device=DEV, generator=gen), dim=-1)
    view = torch.nn.functional.normalize(nrm + 0.5 * torch.randn(P, 3, device=DEV, generator=gen), dim=-1)
    bc, ro, az = r(P, 3), 0.05 + 0.9 * r(P, 1), r(P) * 2 * math.pi
    color = (r(P, S, 3) * (r(P, S, 1) < 0.5)).requires_grad_(True)
    alpha = ((r(P, S) * 1.1).clamp(0, 0.999) * (r(P, S) < 0.6)).requires_grad_(True)
    env_base = (torch.randn(256, 512, 3, device=DEV, generator=gen) * 0.5).requires_grad_(True)
    run = lambda: shading.shade_incident(nrm, S, bc, ro, view, env_base, color, alpha, azimuth=az,   # noqa: E731
                                         transmittance_min=T_MIN)
    out = run()
    with torch.no_grad():
        cn, an = osh.normalise_trace(color, alpha, T_MIN)
        assert float((out["visibility"][:, 0] - (1 - an.mean(1))).abs().max()) <= 2e-5
        assert float((out["light_indirect"] - cn.mean(1)).abs().max()) <= 2e-5
        again = run()
        for k in out:
            assert torch.equal(out[k], again[k]), k
    w = {k: torch.randn(out[k].shape, device=DEV, generator=gen) for k in ("diffuse", "specular", "light_direct")}
    grads = []
    for scale in (1.0, -2.5):
        for t in (color, alpha, env_base):
            t.grad = None
        sum((run()[k] * w[k]).sum() for k in w).mul(scale).backward()
        grads.append([color.grad.clone(), alpha.grad.clone(), env_base.grad.clone()])
    for a, b, tol in zip(grads[0], grads[1], (1e-6, 1e-6, 2e-4)):     # texel sums are atomics: order noise
        assert float((a * -2.5 - b).abs().max()) <= tol * max(1.0, float(b.abs().max()))


def test_rendering_equation_from_surfel_parameters(small_scene):
    """SurfelScene.rendering_equation: the whole path from (means, scales, rotations, opacities, SH) and the shading-point
    inputs to diffuse / specular / light_direct; gradients reach every parameter and agree with the composition of the
    separately verified pieces (surfel_frames -> trace on materialised rays -> the oracle's torch rendering equation)."""
    from irgs_b200 import shading
    from irgs_b200.surfels import SurfelScene, surfel_frames
    sc, inp = small_scene
    gen = torch.Generator().manual_seed(23)
    P, S = 64, 32
    idx = torch.randint(0, inp["means3D"].shape[0], (P,), generator=gen)
    nrm = inp["normals"][idx].contiguous().to(DEV)
    pos = (inp["means3D"][idx] + 0.01 * inp["normals"][idx]).contiguous().to(DEV)
    view = torch.nn.functional.normalize(torch.tensor(synth.CAMERA_CENTER, dtype=torch.float32, device=DEV)[None] - pos, dim=-1)
    azim = (torch.rand(P, generator=gen) * 2 * math.pi).to(DEV)
    bc, ro = torch.rand(P, 3, generator=gen).to(DEV), (0.1 + 0.8 * torch.rand(P, 1, generator=gen)).to(DEV)
    w = [torch.randn(P, 3, generator=gen).to(DEV) for _ in range(3)]
    keys = ("means", "scales", "rotations", "opacity", "shs")
    res = []
    for fused in (True, False):
        leaf = {k: sc[k].to(DEV).clone().requires_grad_(True) for k in keys}
        env = shading.EnvLight(resolution=(16, 32), device=DEV)
        scene = SurfelScene(transmittance_min=synth.T_MIN, alpha_min=synth.ALPHA_MIN)
        scene.build(leaf["means"], leaf["scales"], leaf["rotations"], leaf["opacity"], synth.CAMERA_CENTER)
        if fused:
            out = scene.rendering_equation(bc, ro, nrm, pos, view, leaf["means"], leaf["scales"], leaf["rotations"],
                                           leaf["opacity"], leaf["shs"], env, S, training=True, azimuth=azim,
                                           camera_center=synth.CAMERA_CENTER)
            assert set(out) == {"diffuse", "specular", "light_direct"}
        else:
            ru, rv, n = surfel_frames(leaf["means"], leaf["scales"], leaf["rotations"], synth.CAMERA_CENTER)
            d = _torch_dirs(nrm, S, azim)
            color, _, _, _, alpha = scene.tracer.trace(pos[:, None] + d * 0.05, d, leaf["means"], leaf["opacity"], ru, rv, n,
                                                       None, leaf["shs"], synth.ALPHA_MIN)
            out = osh.rendering_equation(bc, ro, nrm, view, d, color, alpha, env.base, "exp", None, synth.T_MIN)
        sum((out[k] * wi).sum() for k, wi in zip(("diffuse", "specular", "light_direct"), w)).backward()
        res.append(({k: out[k].detach() for k in ("diffuse", "specular", "light_direct")},
                    {**{k: leaf[k].grad for k in keys}, "env_base": env.base.grad}))
    for k in res[0][0]:
        assert float((res[0][0][k] - res[1][0][k]).abs().max()) <= 1e-4 * max(1.0, float(res[1][0][k].abs().max())), k
    for k in res[0][1]:
        a = res[0][1][k].double().flatten()
        b = torch.nan_to_num(res[1][1][k].double().flatten(), 0.0, 0.0, 0.0)
        assert torch.isfinite(a).all() and float(b.norm()) > 0, k
        assert float(a @ b / (a.norm() * b.norm() + 1e-300)) >= 0.9999, k


def test_light_sampling_matches_the_oracle_and_the_unfused_composition(small_scene):
    """light_sample_num > 0 (gaussian_renderer/__init__.py:340-357): EnvLight.update_pdf / light_pdf / sample_light_directions
    against the oracle's restatement of scene/light.py:174-223, and rendering_equation with light samples against the
    un-fused composition (tracer on materialised rays + the oracle's torch code) on the very same light directions."""
    from irgs_b200 import shading
    from irgs_b200.raytracer import GaussianTracer
    sc, inp = small_scene
    gen = torch.Generator().manual_seed(31)
    env = shading.EnvLight(resolution=(16, 32), device=DEV)
    env.base.data += (0.8 * torch.randn(16, 32, 3, generator=gen)).to(DEV)
    env.update_pdf()
    pdf_ref = osh.update_pdf(env.base.detach().cpu(), "exp")
    assert float((env._pdf.cpu() - pdf_ref).abs().max()) <= 1e-6 * float(pdf_ref.max())
    for training in (False, True):
        d, p = env.sample_light_directions(4096, 8, training)
        assert d.shape == (4096, 8, 3) and p.shape == (4096, 8, 1)
        assert float((d.norm(dim=-1) - 1).abs().max()) <= 1e-5
        pr = osh.light_pdf(pdf_ref, d.cpu())
        assert float(((p.cpu() - pr).abs() / pr.clamp_min(1e-6)).median()) <= 1e-5
        # texel frequencies follow the probabilities (32 768 draws over 512 texels)
        l = d.reshape(-1, 3).cpu()
        u = torch.atan2(l[:, 0], -l[:, 2]) / (2 * math.pi) + 0.5
        v = torch.acos(l[:, 1].clamp(-1, 1)) / math.pi
        idx = (u * 32).clamp(0, 31).long() + (v * 16).clamp(0, 15).long() * 32
        freq = torch.bincount(idx, minlength=512).float() / idx.numel()
        assert float((freq - pdf_ref.reshape(-1)).abs().max()) <= 6 * math.sqrt(float(pdf_ref.max()) / idx.numel())
    # end to end with fixed light directions
    g = {k: v.to(DEV) for k, v in inp.items()}
    tr = GaussianTracer(transmittance_min=synth.T_MIN)
    tr.build_from_surfels(g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], synth.ALPHA_MIN)
    P, S, Sl = 64, 24, 16
    idx = torch.randint(0, inp["means3D"].shape[0], (P,), generator=gen)
    nrm = inp["normals"][idx].contiguous().to(DEV)
    pos = (inp["means3D"][idx] + 0.01 * inp["normals"][idx]).contiguous().to(DEV)
    view = torch.nn.functional.normalize(torch.tensor(synth.CAMERA_CENTER, dtype=torch.float32, device=DEV)[None] - pos, dim=-1)
    azim = (torch.rand(P, generator=gen) * 2 * math.pi).to(DEV)
    light_dirs, _ = env.sample_light_directions(P, Sl, True)
    env.sample_light_directions = lambda B, n, training=False: (light_dirs, None)
    keys_s = ("means3D", "opacity", "ru", "rv", "normals", "features", "shs")
    w = [torch.randn(P, 3, generator=gen).to(DEV) for _ in range(3)]
    res = []
    for fused in (True, False):
        leaf = {k: g[k].clone().requires_grad_(True) for k in keys_s}
        pt = {"base_color": torch.rand(P, 3, generator=torch.Generator().manual_seed(5)).to(DEV).requires_grad_(True),
              "roughness": (0.1 + 0.8 * torch.rand(P, 1, generator=torch.Generator().manual_seed(6))).to(DEV).requires_grad_(True),
              "normals_pt": nrm.clone().requires_grad_(True), "position": pos.clone().requires_grad_(True)}
        env.base.grad = None
        surf = tuple(leaf[k] for k in keys_s)
        if fused:
            out = shading.rendering_equation(pt["base_color"], pt["roughness"], pt["normals_pt"], pt["position"], view, tr, surf,
                                             env, S, training=True, azimuth=azim, light_sample_num=Sl, light_t_min=0.05,
                                             alpha_min=synth.ALPHA_MIN)
        else:
            d = torch.cat([_torch_dirs(pt["normals_pt"], S, azim), light_dirs], dim=1)
            color, _, _, _, alpha = tr.trace(pt["position"][:, None] + d * 0.05, d, *surf, synth.ALPHA_MIN)
            areas = osh.mis_areas(d, env._pdf, S, Sl)
            out = osh.rendering_equation(pt["base_color"], pt["roughness"], pt["normals_pt"], view, d, color, alpha, env.base,
                                         "exp", None, synth.T_MIN, incident_areas=areas)
        sum((out[k] * wi).sum() for k, wi in zip(("diffuse", "specular", "light_direct"), w)).backward()
        res.append(({k: out[k].detach() for k in ("diffuse", "specular", "light_direct")},
                    {**{k: v.grad for k, v in pt.items()}, **{k: leaf[k].grad for k in keys_s if k != "features"},
                     "env_base": env.base.grad.clone()}))
    for k in res[0][0]:
        assert float((res[0][0][k] - res[1][0][k]).abs().max()) <= 1e-4 * max(1.0, float(res[1][0][k].abs().max())), k
    for k in res[0][1]:
        a = res[0][1][k].double().flatten()
        b = torch.nan_to_num(res[1][1][k].double().flatten(), 0.0, 0.0, 0.0)
        assert torch.isfinite(a).all() and float(b.norm()) > 0, k
        assert float(a @ b / (a.norm() * b.norm() + 1e-300)) >= 0.9999, k
    with pytest.raises(RuntimeError):
        shading.rendering_equation(pt["base_color"], pt["roughness"], nrm, pos, view, tr, surf,
                                   shading.EnvLight(resolution=(8, 16), device=DEV), S, light_sample_num=4)


# ---------------------------------------------------------------------------------------------- relight branch
RELIGHT_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_relight.npz")
RELIGHT_CASES = ("eval32", "eval24_light12_xf", "eval16_sigmoid_wo")


def _relight_case(name):
    z = np.load(RELIGHT_GOLDEN)
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def _relight_env(case, device):
    t = lambda k: torch.from_numpy(case[k])                                                     # noqa: E731
    tr = t("in_transform") if "in_transform" in case else None
    env = osh.RelightEnvStandIn(t("in_base"), t("in_base_diffuse"), t("in_base_spec0"), t("in_base_spec1"), str(case["activation"]), tr)
    return env.to(device)


@pytest.mark.parametrize("name", RELIGHT_CASES)
def test_relight_kernels_match_reference_golden(name):
    """gaussian_renderer/__init__.py:362-381 recorded from the unmodified reference source (oracle/gen_golden_relight.py): the
    relight kernels (irgs_relight_hit / irgs_relight_combine) + the shading epilogue, fed with the recorded raw tracer outputs
    and the same environment stand-in the reference ran with, reproduce all six outputs."""
    from irgs_b200 import incident, shading
    case = _relight_case(name)
    t = lambda k: torch.from_numpy(case[k]).to(DEV)                                             # noqa: E731
    env = _relight_env(case, DEV)
    S, Sl = int(case["S"]), int(case["n_light"])
    f0, wo = float(case["f0"]), bool(case["wo_indirect_relight"])
    nrm, bc, ro, view = t("in_normals"), t("in_base_color"), t("in_roughness"), t("in_viewdirs")
    P = nrm.shape[0]
    dirs = incident.incident_dirs(nrm, S)
    assert np.abs(dirs.cpu().numpy() - case["rays_d"][:, :S]).max() <= 4e-7      # the directions the reference traced
    raw = (t("in_normal_raw"), t("in_feature_raw"), t("in_alpha_raw"))
    common = dict(activation=str(case["activation"]), transform=env.transform, transmittance_min=None)
    local, alpha_n = shading.relight_local_lights(dirs, raw[0][:, :S], raw[1][:, :S], raw[2][:, :S], env, t("in_fg_lut"), f0, T_MIN, wo)
    if Sl == 0:
        out = shading.shade_incident(nrm, S, bc, ro, view, env.base, local, alpha_n, **common)
    else:
        mix = dict(pdf=t("in_pdf"), p_diffuse=S / (S + Sl), p_light=Sl / (S + Sl), total_samples=S + Sl)
        ld = t("in_light_dirs")
        assert np.abs(ld.cpu().numpy() - case["rays_d"][:, S:]).max() == 0
        local_l, alpha_l = shading.relight_local_lights(ld, raw[0][:, S:], raw[1][:, S:], raw[2][:, S:], env, t("in_fg_lut"), f0, T_MIN, wo)
        od = shading.shade_incident(nrm, S, bc, ro, view, env.base, local, alpha_n, **common, **mix)
        ol = shading.shade_incident(nrm, Sl, bc, ro, view, env.base, local_l, alpha_l, dirs=ld, **common, **mix)
        out = {k: od[k] + ol[k] for k in od}
    for k in shading.OUT_SLICES:
        if wo and k == "light_indirect":     # pipe.wo_indirect_relight: exact zeros on both sides
            assert float(out[k].abs().max()) == 0.0 and not case["out_" + k].any()
            continue
        _close(out[k], case["out_" + k].reshape(out[k].shape), f"{name}/{k}", tol=3e-4, cos_min=0.99999)


def test_relight_rendering_equation_end_to_end(small_scene):
    """shading.rendering_equation(relight=True) on a real scene -- S = 4 feature trace (base colour + roughness), hit-point
    shading, epilogue -- against the CPU composition: oracle trace of the same incident rays -> oracle.shading.relight_local ->
    oracle.shading.rendering_equation."""
    import oracle
    from irgs_b200 import incident, shading
    from irgs_b200.raytracer import GaussianTracer
    sc, inp = small_scene
    g = {k: v.to(DEV) for k, v in inp.items()}
    tr = GaussianTracer(transmittance_min=synth.T_MIN)
    tr.build_from_surfels(g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], synth.ALPHA_MIN)
    gen = torch.Generator().manual_seed(44)
    N = inp["means3D"].shape[0]
    feats = torch.cat([torch.rand(N, 3, generator=gen), 0.05 + 0.9 * torch.rand(N, 1, generator=gen)], 1)
    surf = (g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], feats.to(DEV), g["shs"])
    P, S = 48, 32
    idx = torch.randint(0, N, (P,), generator=gen)
    pos = (inp["means3D"][idx] + 0.01 * inp["normals"][idx]).contiguous()
    nrm = inp["normals"][idx].contiguous()
    view = torch.nn.functional.normalize(torch.tensor(synth.CAMERA_CENTER)[None] - pos, dim=-1)
    bc, ro = torch.rand(P, 3, generator=gen), 0.1 + 0.8 * torch.rand(P, 1, generator=gen)
    maps = [0.6 * torch.randn(16, 32, 3, generator=gen) for _ in range(4)]
    env = osh.RelightEnvStandIn(*maps, "exp", None)
    lut = 0.05 + 0.9 * torch.rand(1, 32, 32, 2, generator=gen)
    with torch.no_grad():
        out = shading.rendering_equation(bc.to(DEV), ro.to(DEV), nrm.to(DEV), pos.to(DEV), view.to(DEV), tr, surf,
                                         osh.RelightEnvStandIn(*maps, "exp", None).to(DEV), S, relight=True, fg_lut=lut.to(DEV),
                                         f0=0.02, light_t_min=synth.LIGHT_T_MIN, alpha_min=synth.ALPHA_MIN)
        io, idr = incident.incident_rays(pos.to(DEV), nrm.to(DEV), S, None, synth.LIGHT_T_MIN)
    Sc = oracle.Scene(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], inp["shs"], feats)
    rt = oracle.trace_forward(Sc, io.reshape(-1, 3).cpu(), idr.reshape(-1, 3).cpu())
    ok = torch.from_numpy(((rt["margin"][:, 0] > 2e-4) & (rt["margin"][:, 1] > 2e-4)).reshape(P, S).all(-1))
    assert int(ok.sum()) >= P // 2 and rt["hit_count"].sum() > 0
    f = lambda k, c: torch.from_numpy(rt[k]).view(P, S, c) if c else torch.from_numpy(rt[k]).view(P, S)    # noqa: E731
    local, alpha_n = osh.relight_local(idr.cpu(), f("normal", 3), f("feature", 4), f("alpha", 0), env, lut, 0.02, synth.T_MIN)
    want = osh.rendering_equation(bc, ro, nrm, view, idr.cpu(), local, alpha_n, maps[0], "exp", None, None)
    for k in shading.OUT_SLICES:
        err = float((out[k].cpu() - want[k])[ok].abs().max())
        assert err <= 2e-4 * max(1.0, float(want[k].abs().max())), (k, err)
