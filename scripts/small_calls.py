"""Per-call latency at the reference's own call sizes (trace_num_rays = 2^18, arguments/__init__.py:154): forward and
forward+backward through GaussianTracer.trace + autograd, CUDA events, best of 5."""
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
if os.environ.get("BPS"): tr.set_option("fwd_blocks_per_sm", int(os.environ["BPS"]))
for n in (1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22):
    gout = bench.make_gout(n, dev)
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
    print(f"n={n:8d} fwd {bf:7.3f} ms ({n/bf/1e3:6.1f} Mrays/s)  bwd {bb:7.3f} ms  fwd+bwd {n/(bf+bb)/1e3:6.1f} Mrays/s")
if os.environ.get("BPS"): sys.exit(0)
# where does the small-call floor come from: host enqueue time vs device time
import time
n = 1 << 16
with torch.no_grad():
    for _ in range(3): tr.trace(ro[:n], rd[:n], *args)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): tr.trace(ro[:n], rd[:n], *args)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"n={n}: host enqueue {1e3*(t1-t0)/20:.3f} ms per forward call, with device drain {1e3*(t2-t0)/20:.3f} ms")
    tr.set_stats(True)
    tr.trace(ro[:n], rd[:n], *args); print("stats nodes/leaves/hits/passes per ray", [x / n for x in tr.get_stats()])
    tr.set_stats(False)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    with torch.no_grad():
        tr.trace(ro[:n], rd[:n], *args)
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
    print(f"  {e.key[:60]:60s} {e.device_time_total:9.1f} us x{e.count}")
