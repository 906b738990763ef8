"""Where does the latency of a small forward call come from?  Times a 2^14-ray call, then the same rays split by how many
hits they composite (the rays with the most hits alone, the rest alone), and prints the traversal statistics of each part.
    python scripts/drain_diag.py            (IRGS_LIB=<other build> to compare builds)"""
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
args = (inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], None, inp["shs"], synth.ALPHA_MIN)
def timed(o, d, label):
    best = 1e9
    with torch.no_grad():
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tr.trace(o, d, *args); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tr.set_stats(True); tr.trace(o, d, *args); torch.cuda.synchronize(); st = tr.get_stats(); tr.set_stats(False)
    n = o.shape[0]
    print(f"{label:34s} n={n:7d} fwd {best:7.3f} ms  per ray: nodes {st[0]/n:8.1f} leaves {st[1]/n:7.1f} hits {st[2]/n:6.2f} passes {st[3]/n:5.2f}", flush=True)
    return best
for n in (1 << 14, 1 << 16, 1 << 18, 1 << 20):
    o, d = ro[:n].contiguous(), rd[:n].contiguous()
    timed(o, d, "all rays")
    hc = tr.last_hit_count.view(-1).clone()
    for w in (32, 128, 1024):   # consecutive fetch indices map to rays n/w apart: heavy rays of one pixel bundle land in different warps
        perm = torch.arange(n, device=dev).view(w, n // w).t().reshape(-1)
        timed(o[perm].contiguous(), d[perm].contiguous(), f"  all rays, transposed order ({w})")
    perm = torch.randperm(n, device=dev)
    timed(o[perm].contiguous(), d[perm].contiguous(), "  all rays, random order")
    order = torch.argsort(hc, descending=True)
    print("   hit counts: max", int(hc.max()), "p99.9", int(torch.quantile(hc.float(), 0.999)), "mean", float(hc.float().mean()))
    for top in (32, 256, 2048):
        timed(o[order[:top]].contiguous(), d[order[:top]].contiguous(), f"  {top} rays with the most hits")
    timed(o[order[2048:]].contiguous(), d[order[2048:]].contiguous(), "  all but those 2048")
    z = order[hc[order] == 0]
    if z.numel() > 0:
        timed(o[z].contiguous(), d[z].contiguous(), "  rays that hit nothing")
