"""CPU tests of the shading epilogue (SURVEY.md 8f rank 1 + 3), no GPU needed:

1. the torch restatement oracle/shading.py against golden vectors recorded from the UNMODIFIED reference functions
   (tests/golden/ref_shading.npz, generator oracle/gen_golden_shading.py): outputs and autograd gradients;
2. the per-sample arithmetic the CUDA kernels run (irgs_b200/csrc/shade_math.cuh, compiled for the host by g++ through
   tests/shade_host.cpp) against the same golden vectors and against float64 autograd of the oracle.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import shading as osh

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ref_shading.npz")
CASES = ("eval24", "train64", "train40_sigmoid_xf", "eval33_none", "eval24_light12", "train32_light16_xf")
OUT_KEYS = ("diffuse", "specular", "light_direct", "visibility", "light", "light_indirect")
OUT_SLICES = {"diffuse": slice(0, 3), "specular": slice(3, 6), "light_direct": slice(6, 9), "visibility": slice(9, 10),
              "light": slice(10, 13), "light_indirect": slice(13, 16)}
ACT = {"none": 0, "exp": 1, "sigmoid": 2}
T_MIN = 0.03


def load_case(name):
    z = np.load(GOLDEN)
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def _oracle(case, dtype, dirs_leaf=False):
    t = lambda k: torch.from_numpy(case[k]).to(dtype)          # noqa: E731
    leaves = {k: t("in_" + k).requires_grad_(True) for k in ("base_color", "roughness", "normals", "viewdirs", "color_raw",
                                                            "alpha_raw", "env_base")}
    transform = t("in_transform") if "in_transform" in case else None
    n_light = int(case["n_light"]) if "n_light" in case else 0
    if dirs_leaf:
        dirs = t("rays_d").requires_grad_(True)
        leaves["dirs"] = dirs
    else:
        az = torch.from_numpy(case["in_azimuth"]).to(dtype) if "in_azimuth" in case else None
        dirs = osh.fibonacci_dirs(leaves["normals"], int(case["S"]), az)
        if n_light > 0:
            dirs = torch.cat([dirs, t("in_light_dirs")], dim=1)
    areas = None
    if n_light > 0:   # the light_sample_num > 0 branch: mixed-sampling weights from the (gradient-free) texel probabilities
        pdf = osh.update_pdf(leaves["env_base"].detach().float(), str(case["activation"]))
        assert np.abs(pdf.numpy() - case["in_pdf"]).max() <= 1e-6 * case["in_pdf"].max()
        areas = osh.mis_areas(dirs, t("in_pdf"), int(case["S"]), n_light, transform)
    out = osh.rendering_equation(leaves["base_color"], leaves["roughness"], leaves["normals"], leaves["viewdirs"], dirs,
                                 leaves["color_raw"], leaves["alpha_raw"], leaves["env_base"], str(case["activation"]),
                                 transform, T_MIN, incident_areas=areas)
    return leaves, out, dirs


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    case = load_case(name)
    leaves, out, dirs = _oracle(case, torch.float32)
    assert np.abs(dirs.detach().numpy() - case["rays_d"]).max() <= 4e-7
    keys = [k for k in OUT_KEYS if f"out_{k}" in case]
    assert keys[:3] == ["diffuse", "specular", "light_direct"] and (len(keys) == 6) == (not bool(case["training"]))
    for k in keys:
        ref = case[f"out_{k}"]
        assert np.abs(out[k].detach().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), k
    sum((out[k] * torch.from_numpy(case[f"w_{k}"])).sum() for k in keys).backward()
    for k in ("base_color", "roughness", "normals", "viewdirs", "color_raw", "alpha_raw", "env_base"):
        a, b = leaves[k].grad.numpy().ravel().astype(np.float64), case[f"grad_{k}"].ravel().astype(np.float64)
        assert np.array_equal(np.isfinite(a), np.isfinite(b))          # the restatement reproduces the reference's NaNs too
        a, b = a[np.isfinite(b)], b[np.isfinite(b)]
        cos = a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300)
        assert cos > 0.999999 and np.abs(a - b).max() <= 2e-4 * np.abs(b).max(), (k, cos, np.abs(a - b).max(), np.abs(b).max())


@pytest.fixture(scope="module")
def host():
    so = os.path.join(HERE, "_shade_host.so")
    src = os.path.join(HERE, "shade_host.cpp")
    hdr = os.path.join(HERE, "..", "irgs_b200", "csrc", "shade_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _host_args(case):
    f = lambda k: np.ascontiguousarray(case[k], np.float32)    # noqa: E731
    n_light = int(case["n_light"]) if "n_light" in case else 0
    P, S = case["in_normals"].shape[0], int(case["S"]) + n_light         # the harness takes all samples as explicit directions
    env = f("in_env_base")
    tr = f("in_transform") if "in_transform" in case else None
    pdf = f("in_pdf") if n_light > 0 else None
    arrs = dict(normals=f("in_normals"), view=f("in_viewdirs"), rough=f("in_roughness").reshape(-1), base=f("in_base_color"),
                dirs=f("rays_d"), c=f("in_color_raw"), a=f("in_alpha_raw"), env=env, tr=tr, pdf=pdf)
    head = [ctypes.c_int64(P), ctypes.c_int(S)] + [_p(arrs[k]) for k in ("normals", "view", "rough", "base", "dirs", "c", "a")] + \
           [ctypes.c_float(1 - T_MIN), _p(env), ctypes.c_int(env.shape[0]), ctypes.c_int(env.shape[1]),
            ctypes.c_int(ACT[str(case["activation"])]), _p(tr), _p(pdf), ctypes.c_float(int(case["S"]) / S),
            ctypes.c_float(n_light / S), ctypes.c_int(S)]
    return P, S, arrs, head


@pytest.mark.parametrize("name", CASES)
def test_kernel_arithmetic_matches_reference_golden_on_the_host(host, name):
    case = load_case(name)
    P, S, arrs, head = _host_args(case)
    out = np.zeros((P, 16), np.float32)
    host.shade_host_forward(*head, _p(out))
    keys = [k for k in OUT_KEYS if f"out_{k}" in case]
    for k in keys:
        ref = case[f"out_{k}"]
        assert np.abs(out[:, OUT_SLICES[k]] - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), k
    # backward: the golden functional's weights as incoming gradients
    g_out = np.zeros((P, 16), np.float32)
    for k in keys:
        g_out[:, OUT_SLICES[k]] = case[f"w_{k}"]
    g_c, g_a, g_d = np.zeros((P * S, 3), np.float32), np.zeros(P * S, np.float32), np.zeros((P * S, 3), np.float32)
    g_pt, g_env = np.zeros((P, 16), np.float32), np.zeros_like(arrs["env"])
    host.shade_host_backward(*head, _p(g_out), _p(g_c), _p(g_a), _p(g_d), _p(g_pt), _p(g_env))

    def close(a, b, what, tol=2e-4):
        # The reference's torch.where(alpha < 1 - T_min, color, color / alpha) (gaussian_model.py:748-752) back-propagates
        # 0 / 0 = NaN to colour and alpha of every ray with alpha == 0 (rays that hit nothing; the tracer's backward never
        # reads those entries).  The kernel arithmetic returns the finite gradient of the taken branch there.
        a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
        assert np.isfinite(a).all(), what
        a, b = a[np.isfinite(b)], b[np.isfinite(b)]
        assert b.size > 0
        cos = a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300)
        assert cos > 0.999999 and np.abs(a - b).max() <= tol * np.abs(b).max(), (what, cos, np.abs(a - b).max(), np.abs(b).max())

    # quantities that do not involve the direction -> normal chain: straight against the reference's gradients
    close(g_pt[:, 0:3], case["grad_base_color"], "base_color")
    close(g_pt[:, 3], case["grad_roughness"], "roughness")
    close(g_pt[:, 7:10], case["grad_viewdirs"], "viewdirs")
    close(g_c, case["grad_color_raw"], "color_raw")
    close(g_a, case["grad_alpha_raw"], "alpha_raw")
    close(g_env, case["grad_env_base"], "env_base")
    # dL/d direction and the direct normal term: float64 autograd of the oracle with the directions as a leaf
    leaves, out64, _ = _oracle(case, torch.float64, dirs_leaf=True)
    sum((out64[k] * torch.from_numpy(case[f"w_{k}"]).double()).sum() for k in keys).backward()
    # (only the Fibonacci samples: the light samples are constants of the pipeline, and in evaluation mode they sit exactly on
    # texel centres, where the bilinear lookup has a kink and float32 / float64 pick different one-sided derivatives)
    Sd_ = int(case["S"])
    close(g_d.reshape(P, S, 3)[:, :Sd_], leaves["dirs"].grad.numpy().reshape(P, S, 3)[:, :Sd_], "dirs")
    close(g_pt[:, 4:7], leaves["normals"].grad.numpy(), "normal (direct)", tol=5e-4)   # float32 sums against float64
    # and the full normal gradient = direct + chain through the oracle's differentiable sampling == the reference's
    n = torch.from_numpy(case["in_normals"]).double().requires_grad_(True)
    az = torch.from_numpy(case["in_azimuth"]).double() if "in_azimuth" in case else None
    Sd = int(case["S"])                                   # only the Fibonacci samples depend on the normal
    d = osh.fibonacci_dirs(n, Sd, az)
    (d * torch.from_numpy(g_d.reshape(P, S, 3)[:, :Sd]).double()).sum().backward()
    close(g_pt[:, 4:7] + n.grad.numpy(), case["grad_normals"], "normal (total)", tol=5e-4)


def test_env_lookup_on_the_host_matches_the_oracle_at_the_seams(host):
    """Directions on the wrap seam (u = 0 / 1), at the poles, and axis-aligned: the texel addressing of the kernel's
    arithmetic against the torch restatement."""
    g = torch.Generator().manual_seed(5)
    d = torch.randn(4096, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    special = torch.tensor([[0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [1e-9, 0, 1], [-1e-9, 0, 1],
                            [0, 0.9999999, 1e-4], [0, 0, 0]], dtype=torch.float32)
    d = torch.cat([d, special]).contiguous()
    for act in ("exp", "sigmoid", "none"):
        for res in ((8, 16), (5, 7), (1, 1), (256, 512)):
            base = torch.randn(res[0], res[1], 3, generator=g)
            ref = osh.env_pure(base, d, act).numpy()
            out = np.zeros((d.shape[0], 3), np.float32)
            b = base.numpy()
            host.env_host_forward(ctypes.c_int64(d.shape[0]), _p(d.numpy()), _p(b), ctypes.c_int(res[0]), ctypes.c_int(res[1]),
                                  ctypes.c_int(ACT[act]), None, _p(out))
            # atan2 / acos differ by an ulp between libm and torch: compare with a tolerance scaled by the texel contrast
            err = np.abs(out - ref)
            assert np.median(err) <= 1e-6 and (err > 1e-3 * max(1.0, np.abs(ref).max())).mean() < 2e-3, (act, res, err.max())


@pytest.mark.parametrize("name", ["eval24_light12", "train32_light16_xf"])
def test_oracle_light_sampling_matches_the_reference(name):
    """scene/light.py:174-223 restated (oracle/shading.py) against what the unmodified EnvLight produced while the golden
    vectors were recorded: texel probabilities, the densities of the drawn directions, and -- in evaluation mode, where the
    draws sit on texel centres -- the directions themselves from the texel indices."""
    case = load_case(name)
    base = torch.from_numpy(case["in_env_base"])
    tr = torch.from_numpy(case["in_transform"]) if "in_transform" in case else None
    pdf = osh.update_pdf(base, str(case["activation"]))
    assert np.abs(pdf.numpy() - case["in_pdf"]).max() <= 1e-6 * case["in_pdf"].max()
    assert abs(float(pdf.sum()) - 1.0) <= 1e-5
    dirs = torch.from_numpy(case["in_light_dirs"])
    dens = osh.light_pdf(torch.from_numpy(case["in_pdf"]), dirs, tr).numpy()
    assert np.abs(dens - case["light_pdfs"]).max() <= 1e-5 * np.abs(case["light_pdfs"]).max()
    # the texel every direction came from, and back
    H, W = pdf.shape
    l = dirs.reshape(-1, 3) if tr is None else dirs.reshape(-1, 3) @ tr.T
    u = torch.atan2(l[:, 0], -l[:, 2]) / (2 * np.pi) + 0.5
    v = torch.acos(l[:, 1].clamp(-1, 1)) / np.pi
    idx = (u * W).clamp(0, W - 1).long() + (v * H).clamp(0, H - 1).long() * W
    assert float(torch.from_numpy(case["in_pdf"]).reshape(-1)[idx].min()) > 0        # only texels with probability are drawn
    back = osh.light_dirs_from_texels(idx, H, W, transform=tr).reshape(dirs.shape)
    if not bool(case["training"]):
        assert float((back - dirs).abs().max()) <= 2e-6                              # texel centres
    else:                                                                            # jittered inside the texel
        assert float((back - dirs).norm(dim=-1).max()) <= 1.5 * np.pi / min(H, W)
    # mixed-sampling weights of all samples: finite, positive, bounded by the clamp
    S, Sl = int(case["S"]), int(case["n_light"])
    areas = osh.mis_areas(torch.from_numpy(case["rays_d"]), torch.from_numpy(case["in_pdf"]), S, Sl, tr)
    assert areas.shape == (dirs.shape[0], S + Sl, 1) and bool(torch.isfinite(areas).all())
    assert float(areas.min()) > 0 and float(areas.max()) <= 2 * np.pi * (S + Sl) / S + 1e-3


RELIGHT_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_relight.npz")


@pytest.mark.parametrize("name", ["eval32", "eval24_light12_xf", "eval16_sigmoid_wo"])
def test_oracle_relight_branch_matches_reference_golden(name):
    """oracle.shading.relight_local + rendering_equation against the outputs of the UNMODIFIED reference rendering_equation
    (relight=True, gaussian_renderer/__init__.py:362-381) recorded by oracle/gen_golden_relight.py."""
    z = np.load(RELIGHT_GOLDEN)
    case = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}
    t = lambda k: torch.from_numpy(case[k])                                                     # noqa: E731
    tr = t("in_transform") if "in_transform" in case else None
    act = str(case["activation"])
    env = osh.RelightEnvStandIn(t("in_base"), t("in_base_diffuse"), t("in_base_spec0"), t("in_base_spec1"), act, tr)
    S, Sl = int(case["S"]), int(case["n_light"])
    dirs = t("rays_d")
    local, alpha_n = osh.relight_local(dirs, t("in_normal_raw"), t("in_feature_raw"), t("in_alpha_raw"), env, t("in_fg_lut"),
                                       float(case["f0"]), 0.03, bool(case["wo_indirect_relight"]))
    areas = osh.mis_areas(dirs, t("in_pdf"), S, Sl, tr) if Sl > 0 else None
    out = osh.rendering_equation(t("in_base_color"), t("in_roughness"), t("in_normals"), t("in_viewdirs"), dirs, local, alpha_n,
                                 t("in_base"), act, tr, None, areas)
    for k, v in out.items():
        ref = case["out_" + k].reshape(v.shape)
        assert np.abs(v.numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), (k, np.abs(v.numpy() - ref).max())
