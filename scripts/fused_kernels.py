"""Per-kernel device time of one forward+backward call: rays in memory (trace) vs generated in the kernels (trace_incident)
vs the whole rendering equation, on 2^22 C3-like rays (torch profiler, CUDA activities)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from irgs_b200 import synth, incident, shading
from irgs_b200.raytracer import GaussianTracer
import bench
dev = torch.device("cuda:0")
class A: surfels=300000; img=128; spp=256
def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr
sc, inp, tr, ro, rd = bench.build_workload(A, dev, 0, 1, factory)
pts, nrm, azim = bench.build_workload.points
o, d = incident.incident_rays(pts, nrm, 256, azim, synth.LIGHT_T_MIN)
o, d = o.reshape(-1, 3), d.reshape(-1, 3)
n = o.shape[0]
tr.accumulate_grads = True
leaf = {k: inp[k].clone().requires_grad_(True) for k in ("means3D", "opacity", "ru", "rv", "normals", "shs")}
largs = (leaf["means3D"], leaf["opacity"], leaf["ru"], leaf["rv"], leaf["normals"], None, leaf["shs"], synth.ALPHA_MIN)
gout = bench.make_gout(n, dev)
gf = [g.view(n // 256, 256, *g.shape[1:]) if g.numel() else g for g in gout]
pl, nl = pts.clone().requires_grad_(True), nrm.clone().requires_grad_(True)
env = shading.EnvLight(resolution=(256, 512), device=dev)
bc = torch.rand(pts.shape[0], 3, device=dev).requires_grad_(True)
rg = (0.1 + 0.8 * torch.rand(pts.shape[0], 1, device=dev)).requires_grad_(True)
view = torch.nn.functional.normalize(torch.tensor(synth.CAMERA_CENTER, device=dev)[None] - pts, dim=-1)
gpix = torch.randn(pts.shape[0], 9, device=dev)
def plain():
    outs = tr.trace(o, d, *largs)
    torch.autograd.backward([outs[0], outs[1], outs[3], outs[4]], [gout[0], gout[1], gout[3], gout[4]])
def fused():
    outs = tr.trace_incident(pl, nl, 256, *largs, azimuth=azim, t_min=synth.LIGHT_T_MIN)
    torch.autograd.backward([outs[0], outs[1], outs[3], outs[4]], [gf[0], gf[1], gf[3], gf[4]])
def shaded():
    out = shading.rendering_equation(bc, rg, nl, pl, view, tr, largs[:7], env, 256, training=True, azimuth=azim,
                                     light_t_min=synth.LIGHT_T_MIN, alpha_min=synth.ALPHA_MIN)
    torch.autograd.backward([out["diffuse"], out["specular"], out["light_direct"]], [gpix[:, 0:3], gpix[:, 3:6], gpix[:, 6:9]])
for name, fn in (("rays in memory", plain), ("generated in the kernels", fused), ("rendering equation", shaded)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    rows = sorted(((e.device_time_total, e.count, e.key) for e in prof.key_averages() if e.device_time_total > 0), reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"== {name}: {n} rays, {e0.elapsed_time(e1):.3f} ms by events, {tot / 1e3:.3f} ms summed over kernels")
    for t, c, k in rows[:12]:
        print(f"   {t / 1e3:8.3f} ms x{c:<3d} {k[:100]}")
