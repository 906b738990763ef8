"""The drop-in boundary as IRGS reaches it: the top-level `surfel_tracer` package of this repository, and
the reference's OWN `surfel_tracer/raytracer.py` -- unmodified, read from baseline/_ref where oracle/build_ref.sh installs
the reference -- running on top of the ctypes stand-in `surfel_tracer/_C.py` for its pybind extension.

Bars as in test_gpu_parity.py: outputs within 1e-4 of the oracle, gradients cosine >= 0.9999 and <= 2e-3 relative.
"""
import importlib.util
import inspect
import os
import sys

import numpy as np
import pytest
import torch

import oracle
from irgs_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RAYTRACER = os.path.join(ROOT, "baseline", "_ref", "surfel_tracer", "raytracer.py")
DEV = "cuda:0"
KEYS = ("means3D", "opacity", "ru", "rv", "normals", "features", "shs")
NAMES = ("color", "normal", "feature", "depth", "alpha")


def _root_package():
    """`import surfel_tracer` the way IRGS does it with this repository first on PYTHONPATH."""
    for m in [m for m in sys.modules if m == "surfel_tracer" or m.startswith("surfel_tracer.")]:
        if "baseline" in (getattr(sys.modules[m], "__file__", "") or ""):
            del sys.modules[m]
    import surfel_tracer
    assert os.path.dirname(os.path.abspath(surfel_tracer.__file__)) == os.path.join(ROOT, "surfel_tracer")
    return surfel_tracer


def test_package_surface_cpu():
    """No GPU needed: the package resolves to the native tracer and the `_C` stand-in has the methods and the argument
    counts of the reference's pybind class (src/bindings.cu:30-94, 105-116)."""
    st = _root_package()
    from irgs_b200.raytracer import GaussianTracer
    assert st.GaussianTracer is GaussianTracer
    from surfel_tracer import _C
    assert callable(_C.create_gaussiantracer)
    arity = dict(build_bvh=1, update_bvh=1, intersection_test=9, trace_forward=19, trace_backward=33)
    for name, n in arity.items():
        params = list(inspect.signature(getattr(_C.GaussianTracer, name)).parameters)
        assert len(params) - 1 == n, (name, len(params) - 1)
    sig = inspect.signature(GaussianTracer.trace)
    assert list(sig.parameters)[1:] == ["rays_o", "rays_d", "means3D", "opacity", "ru", "rv", "normals", "features", "shs",
                                        "alpha_min", "deg", "back_culling"]
    assert sig.parameters["deg"].default == 3 and sig.parameters["back_culling"].default is False


def _case(n=6000, n_features=3):
    sc = synth.make_scene(n, n_features=n_features, scale_mult=4.0)
    inp = synth.derive_tracer_inputs(sc, synth.CAMERA_CENTER)
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, n, (24,), generator=g)
    so, sd = synth.secondary_rays(inp["means3D"][idx] + 0.01 * inp["normals"][idx], inp["normals"][idx], 64, seed=3)
    po, pd = synth.primary_rays(40, 40)
    o = torch.cat([po, so.reshape(-1, 3)]).contiguous()
    d = torch.cat([pd, sd.reshape(-1, 3)]).contiguous()
    return sc, inp, o, d


def _check(tracer, sc, inp, o, d, deg=3):
    S = oracle.Scene(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], inp["shs"], inp["features"])
    ref = oracle.trace_forward(S, o, d, deg=deg, hit_cap=4)
    m = ref["margin"]
    safe = (m[:, 0] > 2e-5) & (m[:, 1] > 2e-5)
    assert safe.mean() > 0.95 and (ref["hit_count"] > 0).mean() > 0.3
    gen = torch.Generator().manual_seed(synth.GRAD_SEED)
    R = o.shape[0]
    gout = dict(color=torch.randn(R, 3, generator=gen), normal=torch.randn(R, 3, generator=gen),
                feature=torch.randn(R, S.S, generator=gen), depth=torch.randn(R, generator=gen),
                alpha=torch.randn(R, generator=gen))
    ts = torch.from_numpy(safe)
    gout = {k: v * (ts[:, None] if v.dim() == 2 else ts) for k, v in gout.items()}
    rb = oracle.trace_backward(S, o, d, ref, {k: v.numpy() for k, v in gout.items()}, deg=deg)
    leaf = {k: inp[k].to(DEV).clone().requires_grad_(True) for k in KEYS}
    ro, rd = o.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    outs = tracer.trace(ro, rd, leaf["means3D"], leaf["opacity"], leaf["ru"], leaf["rv"], leaf["normals"], leaf["features"],
                        leaf["shs"], synth.ALPHA_MIN, deg=deg)
    assert len(outs) == 5
    for name, t in zip(NAMES, outs):
        assert t.shape[0] == R and t.dtype == torch.float32 and t.is_cuda
        err = np.abs(t.detach().cpu().numpy() - ref[name])[safe].max()
        assert err <= 1e-4, (name, err)
    sum((t * gout[nm].to(DEV)).sum() for nm, t in zip(NAMES, outs)).backward()
    got = dict(rays_o=ro.grad, rays_d=rd.grad, means=leaf["means3D"].grad, opacity=leaf["opacity"].grad, ru=leaf["ru"].grad,
               rv=leaf["rv"].grad, normals=leaf["normals"].grad, features=leaf["features"].grad, shs=leaf["shs"].grad)
    for k, v in got.items():
        a, b = v.cpu().numpy().reshape(rb[k].shape).astype(np.float64), rb[k].astype(np.float64)
        cos = float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
        rel = np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
        assert cos >= 0.9999 and rel <= 2e-3, (k, cos, rel)


@pytest.mark.gpu
def test_top_level_package_traces_like_the_oracle():
    """`from surfel_tracer import GaussianTracer` (scene/gaussian_model.py:16) + the caller's build_bvh(vertices_b, faces_b,
    gs_idxs) / update_bvh / trace sequence of scene/gaussian_model.py:725-749."""
    st = _root_package()
    sc, inp, o, d = _case()
    tracer = st.GaussianTracer(transmittance_min=synth.T_MIN)
    assert tracer.transmittance_min == synth.T_MIN     # read by gaussian_model.py:752
    vb, fb, gid = synth.proxy_mesh({k: v.to(DEV) for k, v in sc.items()}, synth.ALPHA_MIN)
    tracer.build_bvh(vb, fb, gid)
    tracer.update_bvh(vb, fb, gid)
    _check(tracer, sc, inp, o, d)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF_RAYTRACER), reason="baseline/_ref not installed (oracle/build_ref.sh)")
def test_reference_raytracer_runs_unmodified_over_the_C_stub():
    """The reference's raytracer.py, byte for byte, with `from surfel_tracer import _C` resolving to surfel_tracer/_C.py of this
    repository: mask / compaction pre-pass, its own autograd Function, re-trace backward."""
    _root_package()
    spec = importlib.util.spec_from_file_location("reference_raytracer_unmodified", REF_RAYTRACER)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from surfel_tracer import _C
    assert mod._C is _C and "irgs_b200" not in open(REF_RAYTRACER).read()
    sc, inp, o, d = _case()
    tracer = mod.GaussianTracer(transmittance_min=synth.T_MIN)
    assert isinstance(tracer.impl, _C.GaussianTracer)
    vb, fb, gid = synth.proxy_mesh({k: v.to(DEV) for k, v in sc.items()}, synth.ALPHA_MIN)
    tracer.build_bvh(vb, fb, gid)
    tracer.update_bvh(vb, fb, gid)
    torch.cuda.set_device(0)
    _check(tracer, sc, inp, o, d)
