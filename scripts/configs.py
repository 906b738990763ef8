"""The other configurations of BASELINE.json (C1, C2, C4, C5 shapes; C3 is bench.py) on ONE GPU, as a table:
rays/s and, where the configuration has it, build / refit time.  Parity for these shapes is covered by tests/."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from irgs_b200 import synth, incident
from irgs_b200.raytracer import GaussianTracer
dev = torch.device("cuda:0")
KEYS = ("means3D", "opacity", "ru", "rv", "normals", "shs")


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def scene(n, feats=0):
    sc = synth.make_scene(n, n_features=feats, device=dev)
    inp = synth.derive_tracer_inputs(sc, synth.CAMERA_CENTER)
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    t_build = timed(lambda: tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN), 3)
    t_refit = timed(lambda: tr.update_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN), 5)
    return sc, inp, tr, t_build, t_refit


def shading_points(tr, inp, img):
    o, d = synth.primary_rays(img, img, device=dev)
    with torch.no_grad():
        outs = tr.trace(o, d, *[inp[k] for k in KEYS[:5]], None, inp["shs"], synth.ALPHA_MIN)
    return synth.shading_points_from_primary(o, d, outs[3], outs[4], outs[1])


def fwd_bwd(tr, leaf, o, d, feats=None, chunk=1 << 22):
    tr.accumulate_grads = True
    g = torch.Generator(dev).manual_seed(1)
    n = o.shape[0]
    gc = torch.randn(min(chunk, n), 3, device=dev, generator=g)

    def body(b, e):
        outs = tr.trace(o[b:e], d[b:e], *[leaf[k] for k in KEYS[:5]], feats, leaf["shs"], synth.ALPHA_MIN)
        torch.autograd.backward([outs[0], outs[4]], [gc[:e - b], gc[:e - b, 0]])
    tr.run_chunks(n, chunk, body)
    return tr.flush_grads(K=16, opacity_shape=tuple(leaf["opacity"].shape))


rows = []
# C1: 10k surfels, 64 x 64 primary rays, forward
sc, inp, tr, tb, tf = scene(10000)
o, d = synth.primary_rays(64, 64, device=dev)
ms = timed(lambda: tr.trace_with_hits(o, d, *[inp[k] for k in KEYS[:5]], None, inp["shs"], synth.ALPHA_MIN, hit_cap=0))
rows.append(("C1 10k surfels, 64x64 primary, fwd", o.shape[0], ms, tb, tf))
# C2: 300k surfels, 800 x 800 primary rays, forward + backward
sc, inp, tr, tb, tf = scene(300000)
leaf = {k: inp[k].clone().requires_grad_(True) for k in KEYS}
o, d = synth.primary_rays(800, 800, device=dev)
ms = timed(lambda: fwd_bwd(tr, leaf, o, d))
rows.append(("C2 300k surfels, 800x800 primary, fwd+bwd", o.shape[0], ms, tb, tf))
# C4: 300k surfels, S = 4 features, (512 Fibonacci + 256 uniform-sphere light) rays per pixel, forward only; 200 x 200 pixels here
sc4, inp4, tr4, tb, tf = scene(300000, feats=4)
pts, nrm = shading_points(tr4, inp4, 200)
od, dd = incident.incident_rays(pts, nrm, 512, None, synth.LIGHT_T_MIN)
gl = torch.Generator(dev).manual_seed(7)
ld = torch.nn.functional.normalize(torch.randn(pts.shape[0], 256, 3, device=dev, generator=gl), dim=-1)
o4 = torch.cat([od, pts[:, None] + ld * synth.LIGHT_T_MIN], 1).reshape(-1, 3).contiguous()
d4 = torch.cat([dd, ld], 1).reshape(-1, 3).contiguous()
def c4():
    with torch.no_grad():
        tr4.run_chunks(o4.shape[0], 1 << 22, lambda b, e: tr4.trace(o4[b:e], d4[b:e], *[inp4[k] for k in KEYS[:5]], inp4["features"], inp4["shs"], synth.ALPHA_MIN))
ms = timed(c4, 3)
rows.append(("C4 300k surfels, S=4, 200x200x(512+256) rays, fwd only", o4.shape[0], ms, tb, tf))
del o4, d4, od, dd, ld
# C5: 1M surfels, parameters perturbed + refit every iteration, 128 x 128 x 256 secondary rays, fwd+bwd
sc5, inp5, tr5, tb, tf = scene(1000000)
pts, nrm = shading_points(tr5, inp5, 128)
gen = torch.Generator(dev).manual_seed(3)
az = torch.rand(pts.shape[0], device=dev, generator=gen) * 2 * np.pi
o5, d5 = incident.incident_rays(pts, nrm, 256, az, synth.LIGHT_T_MIN)
o5, d5 = o5.reshape(-1, 3), d5.reshape(-1, 3)
leaf5 = {k: inp5[k].clone().requires_grad_(True) for k in KEYS}
def c5():
    with torch.no_grad():
        leaf5["means3D"].add_(1e-3 * torch.randn_like(leaf5["means3D"]))
        tr5.update_from_surfels(*[leaf5[k] for k in KEYS[:5]], synth.ALPHA_MIN)
    fwd_bwd(tr5, leaf5, o5, d5)
ms = timed(c5, 3)
rows.append(("C5 1M surfels, perturb + refit per iteration, 128x128x256 rays, fwd+bwd", o5.shape[0], ms, tb, tf))
print(f"{'configuration':78s} {'rays':>10s} {'ms':>9s} {'M rays/s':>9s} {'build ms':>9s} {'refit ms':>9s}")
for name, n, ms, tb, tf in rows:
    print(f"{name:78s} {n:10d} {ms:9.3f} {n / ms / 1e3:9.1f} {tb:9.3f} {tf:9.3f}")
