"""Forward (with hit lists) and backward time on 2^22 C3 rays, for each backward mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irgs_b200 import synth
from irgs_b200.raytracer import GaussianTracer
import bench
dev = torch.device("cuda:0")
class A: surfels=300000; img=int(os.environ.get("IMG", 200)); spp=256
def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    if os.environ.get("WIDE_FOLD"): tr.set_option("wide_fold", int(os.environ["WIDE_FOLD"]))
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr
sc, inp, tr, ro, rd = bench.build_workload(A, dev, 0, 1, factory)
n = min(1 << 22, ro.shape[0])
gout = bench.make_gout(n, dev)
leaf = {k: inp[k].clone().requires_grad_(True) for k in ("means3D", "opacity", "ru", "rv", "normals", "shs")}
tr.accumulate_grads = True
if os.environ.get("CARVE"): tr.set_option("bwd_carveout_pct", int(os.environ["CARVE"]))
for mode in [int(m) for m in os.environ.get("MODES", "0,2,1").split(",")]:
    tr.set_option("bwd_mode", mode)
    best_f = best_b = 1e9
    for _ in range(int(os.environ.get("REPS", 5))):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        outs = tr.trace(ro[:n], rd[:n], leaf["means3D"], leaf["opacity"], leaf["ru"], leaf["rv"], leaf["normals"], None, leaf["shs"], synth.ALPHA_MIN)
        e1.record()
        torch.autograd.backward([outs[0], outs[1], outs[3], outs[4]], [gout[0], gout[1], gout[3], gout[4]])
        e2.record(); torch.cuda.synchronize()
        best_f = min(best_f, e0.elapsed_time(e1)); best_b = min(best_b, e1.elapsed_time(e2))
    g = tr.flush_grads(K=16, opacity_shape=tuple(inp["opacity"].shape))
    print(f"bwd_mode={mode} n={n} fwd {best_f:.3f} ms  bwd {best_b:.3f} ms  checksum {float(g['shs'].abs().sum()):.1f}")
