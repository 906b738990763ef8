"""Does running consecutive chunks on two streams (slots 0 / 1) hide the drain of the persistent kernels?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irgs_b200 import synth
from irgs_b200.raytracer import GaussianTracer
import bench
dev = torch.device("cuda:0")
class A: surfels=300000; img=int(os.environ.get("IMG", 400)); spp=256
def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr
sc, inp, tr, ro, rd = bench.build_workload(A, dev, 0, 1, factory)
chunk = 1 << 22
n = (ro.shape[0] // chunk) * chunk
leaf = {k: inp[k].clone().requires_grad_(True) for k in ("means3D", "opacity", "ru", "rv", "normals", "shs")}
args = (leaf["means3D"], leaf["opacity"], leaf["ru"], leaf["rv"], leaf["normals"], None, leaf["shs"], synth.ALPHA_MIN)
gout = bench.make_gout(chunk, dev)
tr.accumulate_grads = True
side = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
def run(two, bwd):
    cur = torch.cuda.current_stream()
    for s in side: s.wait_stream(cur)
    for i, b in enumerate(range(0, n, chunk)):
        k = (i & 1) if two else 0
        tr.set_option("slot", k)
        with torch.cuda.stream(side[k]):
            if bwd:
                outs = tr.trace(ro[b:b + chunk], rd[b:b + chunk], *args)
                torch.autograd.backward([outs[0], outs[1], outs[3], outs[4]], [gout[0], gout[1], gout[3], gout[4]])
            else:
                with torch.no_grad(): tr.trace(ro[b:b + chunk], rd[b:b + chunk], *args)
    tr.set_option("slot", 0)
    for s in side: cur.wait_stream(s)
for bwd in (False, True):
    for two in (False, True):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(two, bwd); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        g = tr.flush_grads(K=16, opacity_shape=tuple(inp["opacity"].shape)) if bwd else None
        print(f"bwd={bwd} two_streams={two}: {n} rays {best:.2f} ms {n/best/1e3:.1f} Mrays/s", (float(g["shs"].abs().sum()) if g else ""))
