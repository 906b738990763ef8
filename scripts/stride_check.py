"""Small-call latency with and without the stride start order (option "stride_rays_max"), forward and forward+backward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irgs_b200 import synth
from irgs_b200.raytracer import GaussianTracer
import bench
dev = torch.device("cuda:0")
class A: surfels=300000; img=128; spp=256
def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr
sc, inp, tr, ro, rd = bench.build_workload(A, dev, 0, 1, factory)
leaf = {k: inp[k].clone().requires_grad_(True) for k in ("means3D", "opacity", "ru", "rv", "normals", "shs")}
args = (leaf["means3D"], leaf["opacity"], leaf["ru"], leaf["rv"], leaf["normals"], None, leaf["shs"], synth.ALPHA_MIN)
for n in (1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 19, 1 << 20):
    gout = bench.make_gout(n, dev)
    row = []
    for limit in (0, 1 << 30):
        tr.set_option("stride_rays_max", limit)
        bf = bb = 1e9
        for _ in range(6):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            outs = tr.trace(ro[:n], rd[:n], *args)
            e1.record()
            torch.autograd.backward([outs[0], outs[1], outs[3], outs[4]], [gout[0], gout[1], gout[3], gout[4]])
            e2.record(); torch.cuda.synchronize()
            bf = min(bf, e0.elapsed_time(e1)); bb = min(bb, e1.elapsed_time(e2))
            for v in leaf.values(): v.grad = None
        row.append(f"{'stride' if limit else 'caller'} order: fwd {bf:6.3f} ms bwd {bb:6.3f} ms")
    print(f"n={n:8d}  " + "  |  ".join(row), flush=True)
