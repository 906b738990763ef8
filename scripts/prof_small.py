"""One small forward call (2^16 C3 rays) for ncu: where does a latency-bound launch spend its time?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irgs_b200 import synth
from irgs_b200.raytracer import GaussianTracer
import bench
dev = torch.device("cuda:0")
class A: surfels=300000; img=64; spp=256
def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr
sc, inp, tr, ro, rd = bench.build_workload(A, dev, 0, 1, factory)
args = (inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], None, inp["shs"], synth.ALPHA_MIN)
n = int(os.environ.get("N", 1 << 16))
tr.set_stats(bool(os.environ.get("STATS")))
for _ in range(4):
    with torch.no_grad(): tr.trace(ro[:n], rd[:n], *args)
torch.cuda.synchronize()
if os.environ.get("STATS"): print([x / n for x in tr.get_stats()])
