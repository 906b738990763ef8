"""CPU tests of the oracle (oracle/surfel_oracle.c): internal consistency (brute force vs canonical LBVH), values and
gradients against the float64 torch-autograd twin, and -- when present -- the golden vectors produced by the
unmodified reference tracer on a B200 (tests/golden/ref_optix_*.npz, made by oracle/gen_golden_ref.py)."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from irgs_b200 import synth
from tests import torch_twin

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _scene(inp):
    return oracle.Scene(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], inp["shs"], inp["features"])


def _rays_secondary(inp, n_pix=24, S=32, seed=7):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, inp["means3D"].shape[0], (n_pix,), generator=g)
    pts = inp["means3D"][idx] + 0.01 * inp["normals"][idx]
    o, d = synth.secondary_rays(pts, inp["normals"][idx], S, seed=seed)
    return o.reshape(-1, 3), d.reshape(-1, 3)


def test_brute_force_equals_lbvh(small_scene):
    _, inp = small_scene
    S = _scene(inp)
    o, d = synth.primary_rays(48, 48)
    so, sd = _rays_secondary(inp)
    o, d = torch.cat([o, so]), torch.cat([d, sd])
    a = oracle.trace_forward(S, o, d, use_bvh=False)
    b = oracle.trace_forward(S, o, d, use_bvh=True)
    for k in ("color", "normal", "feature", "depth", "alpha", "hit_count", "hits"):
        assert np.array_equal(a[k], b[k]), k
    assert a["hit_count"].max() > 16, "the test must exercise more than one 16-hit pass"
    assert b["counters"][0] > 0 and b["counters"][1] < a["counters"][1]


def test_back_culling_and_degrees(small_scene):
    _, inp = small_scene
    S = _scene(inp)
    o, d = synth.primary_rays(24, 24)
    full = oracle.trace_forward(S, o, d)
    cull = oracle.trace_forward(S, o, d, back_culling=True)
    assert (cull["hit_count"] <= full["hit_count"]).all()
    # normals are flipped towards the camera, so primary rays see front faces only
    assert np.array_equal(cull["hit_count"], full["hit_count"])
    so, sd = _rays_secondary(inp)
    f2, c2 = oracle.trace_forward(S, so, sd), oracle.trace_forward(S, so, sd, back_culling=True)
    assert c2["hit_count"].sum() < f2["hit_count"].sum()
    d0 = oracle.trace_forward(S, o, d, deg=0)
    assert np.array_equal(d0["alpha"], full["alpha"]) and not np.array_equal(d0["color"], full["color"])


def test_outputs_are_bounded_and_zero_on_miss(small_scene):
    _, inp = small_scene
    S = _scene(inp)
    o, d = synth.primary_rays(32, 32)
    r = oracle.trace_forward(S, o, d)
    miss = r["hit_count"] == 0
    assert miss.any() and (~miss).any()
    for k in ("color", "normal", "depth", "alpha"):
        assert not r[k][miss].any()
    assert (r["alpha"] >= 0).all() and (r["alpha"] <= 1 + 1e-6).all()
    # early termination: alpha >= 1 - T_min only on the terminating hit
    assert (r["alpha"][~miss] > 0).all()


def _twin_inputs(inp, dtype=torch.float64):
    return {k: inp[k].detach().to(dtype).clone().requires_grad_(True)
            for k in ("means3D", "opacity", "ru", "rv", "normals", "features", "shs")}


@pytest.mark.parametrize("deg", [0, 3])
def test_forward_and_gradients_match_autograd_twin(small_scene, deg):
    _, inp = small_scene
    S = _scene(inp)
    o, d = synth.primary_rays(20, 20)
    so, sd = _rays_secondary(inp, n_pix=10, S=24)
    o, d = torch.cat([o, so]), torch.cat([d, sd])
    fwd = oracle.trace_forward(S, o, d, deg=deg, hit_cap=96)
    assert fwd["hit_count"].max() <= 96
    g = torch.Generator().manual_seed(synth.GRAD_SEED)
    R = o.shape[0]
    gout = dict(color=torch.randn(R, 3, generator=g), normal=torch.randn(R, 3, generator=g),
                feature=torch.randn(R, S.S, generator=g), depth=torch.randn(R, generator=g),
                alpha=torch.randn(R, generator=g))
    bwd = oracle.trace_backward(S, o, d, fwd, {k: v.numpy() for k, v in gout.items()}, deg=deg)

    tw = _twin_inputs(inp)
    o64, d64 = o.double().requires_grad_(True), d.double().requires_grad_(True)
    outs = torch_twin.composite(o64, d64, tw["means3D"], tw["opacity"], tw["ru"], tw["rv"], tw["normals"],
                                tw["features"], tw["shs"], torch.from_numpy(fwd["hits"]).long(),
                                torch.from_numpy(fwd["hit_count"]).long(), deg=deg)
    names = ("color", "normal", "feature", "depth", "alpha")
    for name, t in zip(names, outs):
        assert np.abs(t.detach().numpy() - fwd[name]).max() < 2e-5, name
    loss = sum((t * gout[n].double()).sum() for n, t in zip(names, outs))
    loss.backward()
    ref = dict(rays_o=o64.grad, rays_d=d64.grad, means=tw["means3D"].grad, opacity=tw["opacity"].grad.reshape(-1),
               ru=tw["ru"].grad, rv=tw["rv"].grad, normals=tw["normals"].grad, features=tw["features"].grad,
               shs=tw["shs"].grad)
    for k, r in ref.items():
        r = r.numpy().reshape(bwd[k].shape)
        scale = np.abs(r).max() + 1e-12
        err = np.abs(bwd[k] - r).max() / scale
        cos = (bwd[k].ravel() @ r.ravel()) / (np.linalg.norm(bwd[k]) * np.linalg.norm(r) + 1e-30)
        assert err < 2e-4 and cos > 0.99999, (k, err, cos)


def test_backward_lbvh_equals_brute_force(small_scene):
    _, inp = small_scene
    S = _scene(inp)
    o, d = synth.primary_rays(16, 16)
    fwd = oracle.trace_forward(S, o, d)
    R = o.shape[0]
    g = torch.Generator().manual_seed(1)
    gout = dict(color=torch.randn(R, 3, generator=g).numpy(), normal=torch.randn(R, 3, generator=g).numpy(),
                feature=torch.randn(R, S.S, generator=g).numpy(), depth=torch.randn(R, generator=g).numpy(),
                alpha=torch.randn(R, generator=g).numpy())
    a = oracle.trace_backward(S, o, d, fwd, gout, use_bvh=False)
    b = oracle.trace_backward(S, o, d, fwd, gout, use_bvh=True)
    for k in a:
        assert np.allclose(a[k], b[k], rtol=1e-6, atol=1e-9), k


def test_surfel_boxes_contain_all_hits(small_scene):
    """The analytic elliptical bounds must contain every accepted hit point (what makes BVH culling exact)."""
    _, inp = small_scene
    S = _scene(inp)
    boxes = S.boxes(synth.ALPHA_MIN)
    o, d = synth.primary_rays(32, 32)
    r = oracle.trace_forward(S, o, d, hit_cap=64)
    o, d = o.numpy(), d.numpy()
    for ray in np.nonzero(r["hit_count"])[0][:200]:
        for g in r["hits"][ray, :min(r["hit_count"][ray], 64)]:
            n, mu = S.normals[g], S.means[g]
            t = -np.dot(n, o[ray] - mu) / np.dot(n, d[ray])
            p = o[ray] + t * d[ray]
            assert (p >= boxes[g, :3] - 1e-6).all() and (p <= boxes[g, 3:] + 1e-6).all()


@pytest.mark.skipif(not glob.glob(os.path.join(GOLDEN, "ref_optix_*.npz")), reason="no reference golden vectors yet")
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_optix_*.npz"))))
def test_oracle_matches_reference_golden(path):
    """Pin the oracle against outputs of the UNMODIFIED reference tracer (OptiX) recorded on a B200."""
    from tests.golden_util import check_against_golden, oracle_runner
    check_against_golden(path, oracle_runner)


def test_incident_sampling_oracle_matches_reference_golden():
    """oracle/incident.py (numpy restatement of utils/graphics_utils.py:19-47,133-165) against vectors recorded from the
    unmodified reference functions (oracle/gen_golden_incident.py)."""
    import os
    from oracle import incident
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_incident.npz"))
    n = g["normals"]
    assert np.array_equal(incident.rotation_between_z(n), g["rotation"])
    for S in (24, 256):
        assert np.abs(incident.incident_dirs(n, S) - g[f"dirs_eval_{S}"]).max() <= 3e-7
        assert np.abs(incident.incident_dirs(n, S, g[f"azimuth_{S}"]) - g[f"dirs_train_{S}"]).max() <= 3e-7
        assert float(g[f"area_{S}"]) == np.float32(2 * np.pi)
    o, d = incident.incident_rays(np.ones((n.shape[0], 3), np.float32), n, 24, None, 0.05)
    assert np.abs(o - (1 + 0.05 * d)).max() <= 1e-7 and o.shape == (n.shape[0], 24, 3)


def test_torch_rotation_between_z_matches_reference_golden():
    import os
    import torch
    from irgs_b200.incident import rotation_between_z
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_incident.npz"))
    R = rotation_between_z(torch.from_numpy(g["normals"])).numpy()
    assert np.abs(R - g["rotation"]).max() <= 1e-7


@pytest.mark.skipif(not glob.glob(os.path.join(GOLDEN, "ref_optix_*.npz")), reason="no reference golden vectors yet")
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_optix_*.npz"))))
def test_chunk_restart_quirk_explains_the_rays_the_strict_comparison_leaves_out(path):
    """The strict golden comparison covers 28 ... 85 % of a case's rays; the rest went through several 16-hit chunks in the
    reference, whose restart may composite the chunk's last surfel twice.  tests/chunk_model.py restates
    exactly that -- proxy-triangle crossings, 16 per chunk, the surfel at every chunk boundary seen again or not -- and here
    (a) ONE of those possibilities reproduces the reference's recorded outputs to 1e-4 on >= 97 % of the left-out rays (measured
        98.5 ... 100 %, median error 2e-7), i.e. the reference differs from the restated math by that rounding quirk and nothing else;
    (b) the possibility "never seen again" IS the oracle (and therefore the CUDA tracer, tests/test_gpu_parity.py): <= 1e-4."""
    from tests import chunk_model as cm
    from tests.golden_util import load
    from oracle.gen_golden_ref import make_case
    z, meta = load(path)
    sc, inp, o, d, _ = make_case(meta)
    S64 = cm.Scene64(sc, inp, meta["alpha_min"])
    nf = meta["n_features"]
    want = np.concatenate([z["out_color"], z["out_normal"], z["out_depth"][:, None], z["out_alpha"][:, None]] +
                          ([z["out_feature"]] if nf else []), 1).astype(np.float64)
    S = oracle.Scene(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], inp["shs"], inp["features"])
    fwd = oracle.trace_forward(S, o, d, alpha_min=meta["alpha_min"], T_min=meta["T_min"], deg=meta["deg"],
                               back_culling=meta["back_culling"])
    mine = np.concatenate([fwd["color"], fwd["normal"], fwd["depth"][:, None], fwd["alpha"][:, None]] +
                          ([fwd["feature"]] if nf else []), 1).astype(np.float64)
    left_out = np.nonzero(~z["strict"])[0]
    rng = np.random.default_rng(7)
    pick = left_out if len(left_out) <= 300 else rng.choice(left_out, 300, replace=False)
    oo, dd = o.double().numpy(), d.double().numpy()
    explained, same_as_oracle = [], []
    for i in pick:
        ok, _, _ = cm.explain_ray(S64, oo[i], dd[i], want[i], meta["deg"], meta["back_culling"], meta["T_min"], nf)
        explained.append(ok)
        _, gs = cm.proxy_hits(S64, oo[i], dd[i])
        plain, _ = cm.composite(S64, oo[i], dd[i], list(gs), meta["deg"], meta["back_culling"], meta["T_min"], nf)
        scale = np.ones_like(plain)
        scale[6] = max(1.0, abs(plain[6]))
        same_as_oracle.append(np.max(np.abs(plain - mine[i]) / scale) <= 1e-4)
    assert np.mean(explained) >= 0.97, np.mean(explained)
    assert np.mean(same_as_oracle) >= 0.97, np.mean(same_as_oracle)      # (threshold-marginal rays excepted)
