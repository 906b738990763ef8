"""One forward launch of trace_incident (rays generated in the kernel) on 2^22 C3 rays between cudaProfilerStart/Stop, for
ncu --profile-from-start off --set full --import-source on (per-line cost of the generation inside the forward kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irgs_b200 import synth
from irgs_b200.raytracer import GaussianTracer
import bench
dev = torch.device("cuda:0")
class A: surfels = 300000; img = 128; spp = 256
def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr
sc, inp, tr, ro, rd = bench.build_workload(A, dev, 0, 1, factory)
pts, nrm, azim = bench.build_workload.points
args = (inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], None, inp["shs"], synth.ALPHA_MIN)
with torch.no_grad():
    for _ in range(2):
        tr.trace_incident(pts, nrm, 256, *args, azimuth=azim, t_min=synth.LIGHT_T_MIN)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr.trace_incident(pts, nrm, 256, *args, azimuth=azim, t_min=synth.LIGHT_T_MIN)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
