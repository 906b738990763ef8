"""Load balance of the ray sharding at 8 ranks: one GPU plays every rank in turn (forward + backward over the rank's rays) for
different interleave block sizes; prints max / mean of the per-rank times."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from irgs_b200 import synth
from irgs_b200.raytracer import GaussianTracer
dev = torch.device("cuda:0")
world = 8
tr_cache = {}
def factory(sc, inp):
    if "tr" not in tr_cache:
        tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
        tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
        tr_cache["tr"] = tr
    return tr_cache["tr"]
for block in (32, 8, 1):
    times = []
    for rank in range(world):
        args = argparse.Namespace(surfels=300000, img=800, spp=256, shard_block=block)
        sc, inp, tr, ro, rd = bench.build_workload(args, dev, rank, world, factory)
        n = ro.shape[0]
        leaf = {k: inp[k].clone().requires_grad_(True) for k in ("means3D", "opacity", "ru", "rv", "normals", "shs")}
        tr.accumulate_grads = True
        gout = bench.make_gout(n // 4, dev)
        best = 1e9
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for b in range(0, n, n // 4):
                e = min(b + n // 4, n)
                outs = tr.trace(ro[b:e], rd[b:e], leaf["means3D"], leaf["opacity"], leaf["ru"], leaf["rv"], leaf["normals"], None,
                                leaf["shs"], synth.ALPHA_MIN)
                torch.autograd.backward([outs[0], outs[1], outs[3], outs[4]], [gout[0][:e - b], gout[1][:e - b], gout[3][:e - b], gout[4][:e - b]])
            e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        times.append(best)
        del ro, rd
        torch.cuda.empty_cache()
    t = np.array(times)
    print(f"block {block:3d}: per-rank ms {np.round(t, 2).tolist()}  max/mean {t.max() / t.mean():.4f}", flush=True)
