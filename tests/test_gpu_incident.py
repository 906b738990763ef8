"""GPU parity tests of the fused incident-ray generation: the rays generated inside the tracing
kernels against the CPU restatement of the reference's sampling (oracle/incident.py) and against the golden vectors
recorded from the unmodified reference functions; the fused trace against the unfused one; the gradients that reach the
shading points against torch autograd through the unfused path."""
import math
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import incident as oinc
from irgs_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_incident.npz")


def _points(inp, n=48, seed=3):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, inp["means3D"].shape[0], (n,), generator=g)
    nrm = inp["normals"][idx]
    pos = inp["means3D"][idx] + 0.01 * nrm
    azim = torch.rand(n, generator=g) * 2 * math.pi
    return pos.contiguous(), nrm.contiguous(), azim.contiguous()


def _tracer(inp):
    from irgs_b200.raytracer import GaussianTracer
    g = {k: v.to(DEV) for k, v in inp.items()}
    tr = GaussianTracer(transmittance_min=synth.T_MIN)
    tr.build_from_surfels(g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], synth.ALPHA_MIN)
    return tr, g


@pytest.mark.parametrize("S", [24, 256])
def test_generated_rays_match_reference_golden_and_oracle(S):
    from irgs_b200 import incident
    gold = np.load(GOLDEN)
    n = torch.from_numpy(gold["normals"]).to(DEV)
    pos = torch.randn(n.shape[0], 3, generator=torch.Generator().manual_seed(1)).to(DEV)
    for mode in ("eval", "train"):
        az = torch.from_numpy(gold[f"azimuth_{S}"]).to(DEV) if mode == "train" else None
        o, d = incident.incident_rays(pos, n, S, az, 0.05)
        d, o = d.cpu().numpy(), o.cpu().numpy()
        assert np.abs(d - gold[f"dirs_{mode}_{S}"]).max() <= 4e-7           # the unmodified reference, run on the CPU
        ro, rd = oinc.incident_rays(pos.cpu().numpy(), gold["normals"], S, None if az is None else gold[f"azimuth_{S}"], 0.05)
        assert np.abs(d - rd).max() <= 4e-7 and np.abs(o - ro).max() <= 1e-6
        assert np.abs(np.linalg.norm(d, axis=-1) - 1).max() <= 2e-6
    assert np.abs(incident.incident_dirs(n, S).cpu().numpy() - gold[f"dirs_eval_{S}"]).max() <= 4e-7


@pytest.mark.parametrize("train", [False, True])
def test_fused_trace_equals_trace_of_the_generated_rays_and_the_oracle(small_scene, train):
    from irgs_b200 import incident
    sc, inp = small_scene
    tr, g = _tracer(inp)
    pos, nrm, azim = _points(inp)
    az = azim.to(DEV) if train else None
    S = 64
    with torch.no_grad():
        fused = tr.trace_incident(pos.to(DEV), nrm.to(DEV), S, g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"],
                                  g["features"], g["shs"], synth.ALPHA_MIN, azimuth=az, t_min=0.05)
        hc_f = tr.last_hit_count.clone()
        o, d = incident.incident_rays(pos.to(DEV), nrm.to(DEV), S, az, 0.05)
        plain = tr.trace(o, d, g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], g["features"], g["shs"],
                         synth.ALPHA_MIN)
    assert fused[0].shape == (pos.shape[0], S, 3) and fused[4].shape == (pos.shape[0], S)
    for a, b in zip(fused, plain):
        assert torch.equal(a, b)            # the very same rays: bit-identical results
    assert torch.equal(hc_f, tr.last_hit_count)
    # against the oracle on the rays of the reference's sampling (restated on the CPU): directions differ by ulps, so
    # only rays without a threshold-marginal hit are compared
    ro, rd = oinc.incident_rays(pos.numpy(), nrm.numpy(), S, azim.numpy() if train else None, 0.05)
    Sc = oracle.Scene(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], inp["shs"], inp["features"])
    ref = oracle.trace_forward(Sc, ro.reshape(-1, 3), rd.reshape(-1, 3))
    safe = (ref["margin"][:, 0] > 2e-4) & (ref["margin"][:, 1] > 2e-4)
    assert safe.mean() > 0.9 and (ref["hit_count"] > 0).mean() > 0.2
    same = hc_f.cpu().numpy().reshape(-1) == ref["hit_count"]
    assert same[safe].mean() > 0.999
    ok = safe & same
    for name, t in zip(("color", "normal", "feature", "depth", "alpha"), fused):
        err = np.abs(t.cpu().numpy().reshape(ref[name].shape) - ref[name])[ok]
        assert err.max() <= 1e-4, (name, err.max())


def _torch_dirs(normals, S, azim):
    """Differentiable torch restatement of graphics_utils.py:19-47 (test infrastructure)."""
    from irgs_b200.incident import rotation_between_z
    idx = torch.arange(S, dtype=torch.float32, device=normals.device)[None]
    z = (1 - 2 * idx / (2 * S - 1)).clamp_min(math.sin(10 / 180 * math.pi))
    rad = torch.sqrt(1 - z ** 2)
    theta = math.pi * (3.0 - math.sqrt(5.0)) * idx
    if azim is not None:
        theta = azim[:, None] + theta
    zs = torch.stack([(torch.sin(theta) * rad).expand(normals.shape[0], S), (torch.cos(theta) * rad).expand(normals.shape[0], S),
                      z.expand(normals.shape[0], S)], -2)
    v = rotation_between_z(normals) @ zs
    return torch.nn.functional.normalize(v, dim=-2).transpose(-1, -2)


@pytest.mark.parametrize("train", [False, True])
def test_gradients_reach_position_normals_and_surfels(small_scene, train):
    sc, inp = small_scene
    tr, g = _tracer(inp)
    pos, nrm, azim = _points(inp, n=40, seed=9)
    az = azim.to(DEV) if train else None
    S, t_min = 32, 0.05
    keys = ("means3D", "opacity", "ru", "rv", "normals", "features", "shs")
    gen = torch.Generator().manual_seed(21)
    w = [torch.randn(pos.shape[0], S, c, generator=gen).to(DEV) for c in (3, 3, inp["features"].shape[1])] + \
        [torch.randn(pos.shape[0], S, generator=gen).to(DEV) for _ in range(2)]
    res = []
    for fused in (True, False):
        leaf = {k: g[k].clone().requires_grad_(True) for k in keys}
        p = pos.to(DEV).clone().requires_grad_(True)
        n = nrm.to(DEV).clone().requires_grad_(True)
        if fused:
            outs = tr.trace_incident(p, n, S, *[leaf[k] for k in keys], synth.ALPHA_MIN, azimuth=az, t_min=t_min)
        else:
            d = _torch_dirs(n, S, az)
            outs = tr.trace(p[:, None] + d * t_min, d, *[leaf[k] for k in keys], synth.ALPHA_MIN)
        sum((o * wi).sum() for o, wi in zip(outs, w)).backward()
        res.append({"position": p.grad, "normals_pt": n.grad, **{k: leaf[k].grad for k in keys}})
    for k in res[0]:
        a, b = res[0][k].double().flatten(), res[1][k].double().flatten()
        cos = float(a @ b / (a.norm() * b.norm() + 1e-300))
        rel = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        # ulp-level direction differences move individual threshold decisions: cosine is the criterion here
        assert cos >= 0.9999, (k, cos, rel)
    # the flip branch (normal = -z) and empty input
    z = torch.tensor([[0.0, 0.0, -1.0]], device=DEV, requires_grad=True)
    p = torch.zeros(1, 3, device=DEV, requires_grad=True)
    outs = tr.trace_incident(p, z, 8, *[g[k] for k in keys], synth.ALPHA_MIN)
    outs[4].sum().backward()
    assert torch.isfinite(z.grad).all() and torch.isfinite(p.grad).all()
    e = tr.trace_incident(p[:0], z[:0], 8, *[g[k] for k in keys], synth.ALPHA_MIN)
    assert e[0].shape == (0, 8, 3)


def test_incident_host_entry_matches_the_device_path(small_scene):
    """irgs_trace_fwd_bwd_incident_host (per-point inputs in pinned host memory, chunks on two internal streams) against
    trace_incident + autograd on device tensors: identical alpha, per-point and per-surfel gradients up to the order noise of
    the float reductions."""
    import ctypes
    from irgs_b200 import _lib
    from irgs_b200.incident import IncidentDesc
    from irgs_b200.raytracer import _ptr
    sc, inp = small_scene
    tr, g = _tracer(inp)
    pos, nrm, azim = _points(inp, n=77, seed=5)
    S, t_min = 32, 0.05
    P, N, F = pos.shape[0], g["means3D"].shape[0], g["features"].shape[1]
    keys = ("means3D", "opacity", "ru", "rv", "normals", "features", "shs")
    period = 13 * S                                    # periodic incoming gradients, a whole number of points
    gen = torch.Generator().manual_seed(8)
    gout = [torch.randn(period, c, generator=gen).to(DEV) for c in (3, 3, F)] + [torch.randn(period, generator=gen).to(DEV) for _ in range(2)]
    # device path
    leaf = {k: g[k].clone().requires_grad_(True) for k in keys}
    p, n = pos.to(DEV).clone().requires_grad_(True), nrm.to(DEV).clone().requires_grad_(True)
    outs = tr.trace_incident(p, n, S, *[leaf[k] for k in keys], synth.ALPHA_MIN, azimuth=azim.to(DEV), t_min=t_min)
    rep = (P * S + period - 1) // period
    full = [t.repeat(*([rep] + [1] * (t.dim() - 1)))[:P * S] for t in gout]
    sum((o.reshape(P * S, *o.shape[2:]) * w).sum() for o, w in zip(outs, full)).backward()
    # host path
    ph, nh, ah = pos.contiguous().pin_memory(), nrm.contiguous().pin_memory(), azim.contiguous().pin_memory()
    al_h, gp_h, gn_h = torch.empty(P * S).pin_memory(), torch.empty(P, 3).pin_memory(), torch.empty(P, 3).pin_memory()
    fused, gfeat = torch.zeros(N, 64, device=DEV), torch.zeros(N, F, device=DEV)
    desc = IncidentDesc(ph.data_ptr(), nh.data_ptr(), ah.data_ptr(), P, S, t_min)
    lib = _lib.load()
    _lib.check(lib.irgs_trace_fwd_bwd_incident_host(
        tr.impl.h, ctypes.byref(desc), F, 16, 3, *[_ptr(g[k]) for k in keys], *[_ptr(t) for t in gout], period, _ptr(al_h),
        _ptr(gp_h), _ptr(gn_h), _ptr(fused), _ptr(gfeat), synth.ALPHA_MIN, synth.T_MIN, 0, 10,
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(al_h.view(P, S), outs[4].detach().cpu())
    for got, want in ((gp_h, p.grad), (gn_h, n.grad)):
        assert (got - want.cpu()).abs().max() <= 2e-4 * (want.abs().max().item() + 1e-30)
    got = tr._unpack(fused, gfeat, tuple(g["opacity"].shape), 16)
    for k, t in zip(keys, got):
        assert (t - leaf[k].grad).abs().max() <= 2e-4 * (leaf[k].grad.abs().max() + 1e-30), k


def test_generation_in_kernel_and_through_the_scratch_block_agree(small_scene):
    """The two ways a forward call can produce generated rays -- a small kernel writing them to the tracer's scratch block (the
    default) or the forward kernel's own ray queue (`gen_in_kernel`) -- run the same device function: bit-identical results, for
    incident and for camera rays."""
    from irgs_b200.primary import Camera, trace_camera
    sc, inp = small_scene
    tr, g = _tracer(inp)
    pos, nrm, azim = _points(inp, n=90, seed=2)
    args = (g["means3D"], g["opacity"], g["ru"], g["rv"], g["normals"], g["features"], g["shs"], synth.ALPHA_MIN)
    cam = Camera.look_at(synth.CAMERA_CENTER, (0.0, 0.0, 0.0), (0.0, 0.0, -1.0), 0.7, 61, 47)
    res = []
    for mode in (0, 1):
        tr.set_option("gen_in_kernel", mode)
        with torch.no_grad():
            a = tr.trace_incident(pos.to(DEV), nrm.to(DEV), 48, *args, azimuth=azim.to(DEV), t_min=0.05)
            b = trace_camera(tr, cam, *args)
        res.append(list(a) + list(b) + [tr.last_hit_count.clone()])
    tr.set_option("gen_in_kernel", 0)
    for x, y in zip(*res):
        assert torch.equal(x, y)
    assert bool(res[0][4].any()) and bool(res[0][9].any())
