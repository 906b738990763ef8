"""Does the ORDER in which the rays of a launch are started matter (long, grazing rays first => short drain)?
Same rays, three orders: bundle-major (as the caller lays them out), sample-major descending (grazing samples of all
pixels first), sample-major ascending."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from irgs_b200 import synth, incident
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
args = (inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], None, inp["shs"], synth.ALPHA_MIN)
def t(o, d):
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad(): tr.trace(o, d, *args)
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for P in (64, 1024, 4096, 16384):
    o, d = incident.incident_rays(pts[:P], nrm[:P], 256, azim[:P], synth.LIGHT_T_MIN)     # [P, S, 3]
    res = []
    for name, perm in (("bundle-major", None), ("sample-major desc", "desc"), ("sample-major asc", "asc")):
        if perm is None:
            oo, dd = o.reshape(-1, 3), d.reshape(-1, 3)
        else:
            idx = torch.arange(255, -1, -1, device=dev) if perm == "desc" else torch.arange(256, device=dev)
            oo = o[:, idx].transpose(0, 1).reshape(-1, 3).contiguous(); dd = d[:, idx].transpose(0, 1).reshape(-1, 3).contiguous()
        res.append(f"{name} {t(oo, dd):.3f} ms")
    print(f"rays={P * 256:8d}: " + " | ".join(res))
