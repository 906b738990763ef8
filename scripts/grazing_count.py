"""How often does the hit test drop a GRAZING ray / surfel pair (|n.d| < 1e-3) that the reference would have evaluated with its
clamped depth formula (gaussiantrace_forward.cu:61-81), on the C3 workload?  Counts from the statistics build of the forward
kernel (first pass of every ray): pairs whose geometric plane crossing lies inside the surfel's support and the depth range, and
how many of them the reference's clamped evaluation would have composited (alpha >= alpha_min at the clamped position).

    python scripts/grazing_count.py [n_pixels]      -> one JSON line
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from irgs_b200 import synth  # noqa: E402
from irgs_b200.raytracer import GaussianTracer  # noqa: E402


def main():
    n_pix = int(sys.argv[1]) if len(sys.argv) > 1 else 800 * 800
    args = argparse.Namespace(surfels=300000, img=800, spp=256)
    dev = torch.device("cuda", 0)

    def factory(sc, inp):
        tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
        tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
        return tr
    sc, inp, tr, rays_o, rays_d = bench.build_workload(args, dev, 0, 1, factory)
    n = min(rays_o.shape[0], n_pix * 256)
    tr.set_stats(True)
    tot = dict(rays=0, hits=0, leaf_tests=0, grazing_pairs=0, grazing_pairs_compositing=0)
    for b in range(0, n, 1 << 22):
        e = min(b + (1 << 22), n)
        with torch.no_grad():
            tr.trace(rays_o[b:e], rays_d[b:e], inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], None,
                     inp["shs"], synth.ALPHA_MIN)
        nodes, leaves, hits, passes = tr.get_stats()
        tot["rays"] += e - b; tot["hits"] += hits; tot["leaf_tests"] += leaves
        tot["grazing_pairs"] += tr.get_info("grazing_pairs")
        tot["grazing_pairs_compositing"] += tr.get_info("grazing_pairs_compositing")
    tot["grazing_per_hit"] = tot["grazing_pairs"] / max(tot["hits"], 1)
    tot["compositing_per_hit"] = tot["grazing_pairs_compositing"] / max(tot["hits"], 1)
    print(json.dumps(tot))


if __name__ == "__main__":
    main()
