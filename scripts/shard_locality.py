"""Does the way pixel bundles are dealt to the ranks change the per-ray cost of the forward kernel?  One GPU plays rank 0 of a
2-rank job with different interleave block sizes (and the whole image for comparison); forward only, one stream, CUDA events."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from irgs_b200 import synth
from irgs_b200.raytracer import GaussianTracer
dev = torch.device("cuda:0")


def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr


for world, block in ((1, 32), (2, 32), (2, 256), (2, 800), (2, 3200), (8, 32), (8, 800)):
    args = argparse.Namespace(surfels=300000, img=800, spp=256, shard_block=block)
    sc, inp, tr, ro, rd = bench.build_workload(args, dev, 0, world, factory)
    n = ro.shape[0]
    chunk = min(n, 16384000)
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            for b in range(0, n, chunk):
                tr.trace(ro[b:b + chunk], rd[b:b + chunk], inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], None,
                         inp["shs"], synth.ALPHA_MIN)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    hits = float(tr.last_hit_count.float().mean())
    print(f"world {world} block {block}: {n} rays, forward {best:.2f} ms, {best * 1e6 / n:.4f} ns/ray, hits/ray (last chunk) {hits:.2f}", flush=True)
    del tr, ro, rd
    torch.cuda.empty_cache()
