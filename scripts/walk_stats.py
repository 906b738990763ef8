"""Traversal statistics of the forward kernel on C3 rays: node visits, leaf tests, hits, passes per ray (statistics build)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from irgs_b200 import synth
from irgs_b200.raytracer import GaussianTracer
dev = torch.device("cuda", 0)
args = argparse.Namespace(surfels=300000, img=int(os.environ.get("IMG", 256)), spp=256)
def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    if os.environ.get("WIDE_FOLD"): tr.set_option("wide_fold", int(os.environ["WIDE_FOLD"]))
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr
sc, inp, tr, ro, rd = bench.build_workload(args, dev, 0, 1, factory)
n = min(ro.shape[0], 1 << int(os.environ.get("LOG2N", 24)))
tr.set_stats(True)
with torch.no_grad():
    tr.trace(ro[:n], rd[:n], inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], None, inp["shs"], synth.ALPHA_MIN)
nodes, leaves, hits, passes = tr.get_stats()
cs, full = tr.get_info("comp_stats"), tr.get_info("full_rows")
rounds = max(full >> 32, 1)
print(json.dumps(dict(comp_rounds_per_ray=rounds / n, longest_segment_per_round=(cs >> 32) / rounds,
                      candidates_per_round=(cs & 0xffffffff) / rounds, full_row_sorts_per_ray=(full & 0xffffffff) / n)))
print(json.dumps(dict(rays=n, nodes_per_ray=nodes / n, leaf_tests_per_ray=leaves / n, hits_per_ray=hits / n, passes_per_ray=passes / n)))
