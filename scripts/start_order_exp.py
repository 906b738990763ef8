"""Small-call tail: does the order in which a 2^18-ray call STARTS its rays matter?  (heavy rays first = longest-processing-time-first)
Orders: the default (caller's bundle-major order + stride start order), and with the stride off: bundle-major, sample-major with the
grazing samples (high Fibonacci index) first / last, and rays sorted by their own hit count (a proxy for their cost; an upper bound on
what any a-priori heuristic could achieve)."""
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
def t(o, d):
    best = 1e9
    with torch.no_grad():
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tr.trace(o, d, *args); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for lg in (16, 18, 20):
    n = 1 << lg; P = n // 256
    o, d = ro[:n].reshape(P, 256, 3), rd[:n].reshape(P, 256, 3)
    res = {}
    tr.set_option("stride_rays_max", 1 << 19)
    res["default (stride)"] = t(o.reshape(-1, 3), d.reshape(-1, 3))
    tr.set_option("stride_rays_max", 0)
    res["bundle-major"] = t(o.reshape(-1, 3), d.reshape(-1, 3))
    os_, ds_ = o.transpose(0, 1).contiguous(), d.transpose(0, 1).contiguous()     # [S, P, 3]
    res["sample-major, grazing last"] = t(os_.reshape(-1, 3), ds_.reshape(-1, 3))
    res["sample-major, grazing first"] = t(os_.flip(0).contiguous().reshape(-1, 3), ds_.flip(0).contiguous().reshape(-1, 3))
    with torch.no_grad():
        tr.trace(o.reshape(-1, 3), d.reshape(-1, 3), *args)
    # cost proxy: hit count (needs hit lists: trace with grad) -- use alpha-weighted depth instead? simplest: re-trace with requires_grad
    leaf = inp["means3D"].clone().requires_grad_(True)
    outs = tr.trace(o.reshape(-1, 3), d.reshape(-1, 3), leaf, *args[1:])
    hc = tr.last_hit_count.to(torch.int64)
    order = torch.argsort(hc, descending=True)
    res["by hit count, heavy first"] = t(o.reshape(-1, 3)[order].contiguous(), d.reshape(-1, 3)[order].contiguous())
    perm = torch.randperm(n, device=dev)
    res["random"] = t(o.reshape(-1, 3)[perm].contiguous(), d.reshape(-1, 3)[perm].contiguous())
    print(f"n=2^{lg}: " + " | ".join(f"{k} {v:.3f}" for k, v in res.items()))
