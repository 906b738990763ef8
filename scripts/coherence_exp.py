"""How much would ray COHERENCE inside a warp buy the forward kernel?  The same 2^22 C3 rays (training mode: random azimuth per
pixel) traced in different physical orders; the order is applied outside the timed region, so the numbers are the kernel's."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irgs_b200 import synth, incident
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
def t(o, d):
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad(): tr.trace(o, d, *args)
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
def stats(o, d):
    tr.set_stats(True)
    with torch.no_grad(): tr.trace(o, d, *args)
    s = tr.get_stats(); tr.set_stats(False)
    return s[0] / o.shape[0], s[1] / o.shape[0]
P = 16384
for mode, az in (("train (random azimuth)", azim[:P]), ("eval (no azimuth)", None)):
    o, d = incident.incident_rays(pts[:P], nrm[:P], 256, az, synth.LIGHT_T_MIN)     # [P, S, 3]
    o2, d2 = o.reshape(-1, 3), d.reshape(-1, 3)
    n = o2.shape[0]
    print(f"{mode}: {n} rays")
    print(f"  bundle-major (caller's order)            {t(o2, d2):7.3f} ms   nodes/leaves per ray {stats(o2, d2)}")
    os_, ds_ = o.transpose(0, 1).reshape(-1, 3).contiguous(), d.transpose(0, 1).reshape(-1, 3).contiguous()
    print(f"  sample-major (32 pixels x same sample)   {t(os_, ds_):7.3f} ms")
    # direction bin (octahedral 16 x 16) major, pixel minor
    s = d2.abs().sum(-1, keepdim=True)
    pxy = d2[:, :2] / s
    neg = d2[:, 2] < 0
    ax = (1 - pxy[:, 1].abs()) * torch.where(pxy[:, 0] >= 0, 1.0, -1.0)
    ay = (1 - pxy[:, 0].abs()) * torch.where(pxy[:, 1] >= 0, 1.0, -1.0)
    px = torch.where(neg, ax, pxy[:, 0]); py = torch.where(neg, ay, pxy[:, 1])
    for bins in (8, 16, 64):
        du = ((px * 0.5 + 0.5) * bins).long().clamp(0, bins - 1); dv = ((py * 0.5 + 0.5) * bins).long().clamp(0, bins - 1)
        key = (du * bins + dv) * n + torch.arange(n, device=dev)            # direction bin major, original (pixel, sample) order minor
        perm = torch.argsort(key)
        print(f"  direction-bin major ({bins:2d} x {bins:2d} bins)        {t(o2[perm].contiguous(), d2[perm].contiguous()):7.3f} ms")
    # per 32-pixel block: direction-bin major inside the block (what a cheap local binning could produce)
    blk = (torch.arange(n, device=dev) // (256 * 32))
    du = ((px * 0.5 + 0.5) * 16).long().clamp(0, 15); dv = ((py * 0.5 + 0.5) * 16).long().clamp(0, 15)
    key = (blk * 256 + du * 16 + dv) * n + torch.arange(n, device=dev)
    perm = torch.argsort(key)
    print(f"  32-pixel blocks, direction-bin major     {t(o2[perm].contiguous(), d2[perm].contiguous()):7.3f} ms")
    perm = torch.randperm(n, device=dev)
    print(f"  random order                             {t(o2[perm].contiguous(), d2[perm].contiguous()):7.3f} ms")
