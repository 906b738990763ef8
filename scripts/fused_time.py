"""Forward time of trace (rays in memory) vs trace_incident (rays generated in the kernel) on 2^22 C3-like rays."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irgs_b200 import synth, incident
from irgs_b200.raytracer import GaussianTracer
import bench
dev = torch.device("cuda:0")
class A: surfels=300000; img=int(os.environ.get("IMG", 128)); spp=256
def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    if os.environ.get("WIDE_FOLD"): tr.set_option("wide_fold", int(os.environ["WIDE_FOLD"]))
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr
sc, inp, tr, ro, rd = bench.build_workload(A, dev, 0, 1, factory)
pts, nrm, azim = bench.build_workload.points
args = (inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], None, inp["shs"], synth.ALPHA_MIN)
o, d = incident.incident_rays(pts, nrm, 256, azim, synth.LIGHT_T_MIN)
o, d = o.reshape(-1, 3), d.reshape(-1, 3)
def t(fn):
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad(): fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
n = o.shape[0]
a = t(lambda: tr.trace(o, d, *args)); b = t(lambda: tr.trace_incident(pts, nrm, 256, *args, azimuth=azim, t_min=synth.LIGHT_T_MIN))
print(f"n={n} rays in memory {a:.3f} ms {n/a/1e3:.1f} Mrays/s | generated in kernel {b:.3f} ms {n/b/1e3:.1f} Mrays/s")
# forward + backward
tr.accumulate_grads = True
leaf = {k: inp[k].clone().requires_grad_(True) for k in ("means3D", "opacity", "ru", "rv", "normals", "shs")}
largs = (leaf["means3D"], leaf["opacity"], leaf["ru"], leaf["rv"], leaf["normals"], None, leaf["shs"], synth.ALPHA_MIN)
gout = bench.make_gout(n, dev)
gf = [g.view(n // 256, 256, *g.shape[1:]) if g.numel() else g for g in gout]
pl, nl = pts.clone().requires_grad_(True), nrm.clone().requires_grad_(True)
def fb_plain():
    outs = tr.trace(o, d, *largs)
    torch.autograd.backward([outs[0], outs[1], outs[3], outs[4]], [gout[0], gout[1], gout[3], gout[4]])
def fb_fused():
    outs = tr.trace_incident(pl, nl, 256, *largs, azimuth=azim, t_min=synth.LIGHT_T_MIN)
    torch.autograd.backward([outs[0], outs[1], outs[3], outs[4]], [gf[0], gf[1], gf[3], gf[4]])
def t2(fn):
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
a = t2(fb_plain); b = t2(fb_fused)
print(f"fwd+bwd: rays in memory {a:.3f} ms | generated in kernel {b:.3f} ms")
