"""GPU diagnostic: forward time vs ray count, traversal statistics, slowest-ray hunt."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
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
args = (inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], None, inp["shs"], synth.ALPHA_MIN)
print("rays", ro.shape[0])
def timeit(o, d, reps=3, cap=0):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = tr.trace_with_hits(o, d, *args, hit_cap=cap); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out
for n in (1<<16, 1<<18, 1<<20, 1<<21, 1<<22, 1<<23, 1<<24, ro.shape[0]):
    n = min(n, ro.shape[0])
    ms, out = timeit(ro[:n], rd[:n])
    print(f"n={n:9d} fwd {ms:8.3f} ms  {n/ms/1e3:8.1f} Mrays/s")
n = 1 << 22
tr.set_stats(True)
ms, out = timeit(ro[:n], rd[:n], reps=1)
st = tr.get_stats(); tr.set_stats(False)
print("stats per ray: nodes %.1f leaves %.1f hits %.2f passes %.3f" % tuple(x / n for x in st))
hc = out["hit_count"]
print("hit frac", (hc > 0).float().mean().item(), "max hits", hc.max().item(), "p99.99", torch.quantile(hc[:1<<20].float(), 0.9999).item())
# slowest-ray hunt: time blocks of 4096 rays
blk = 1 << 14
times = []
for b in range(0, n, blk):
    ms, _ = timeit(ro[b:b+blk], rd[b:b+blk], reps=1)
    times.append(ms)
times = np.array(times)
print("block times (16k rays): median %.3f max %.3f argmax %d" % (np.median(times), times.max(), times.argmax()))
b = int(times.argmax()) * blk
sub = []
for bb in range(b, b + blk, 256):
    ms, o2 = timeit(ro[bb:bb+256], rd[bb:bb+256], reps=1)
    sub.append((ms, bb, o2["hit_count"].max().item()))
sub.sort(reverse=True)
print("slowest 256-ray bundles:", sub[:5])
ms, bb, _ = sub[0]
for r in range(bb, bb + 256, 32):
    ms, o2 = timeit(ro[r:r+32], rd[r:r+32], reps=1)
    print(r, "%.3f ms" % ms, o2["hit_count"].tolist())
