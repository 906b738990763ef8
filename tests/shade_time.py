#!/usr/bin/env python
"""Shading epilogue on its own (SURVEY 8f rank 1 + 3): irgs_shade_forward / irgs_shade_backward against the same math as
element-wise torch kernels (the way the reference's rendering_equation evaluates it), on one chunk of the C3 shape.

    python tests/shade_time.py [--points 16384] [--spp 256]        # prints one JSON line

Bytes per ray (algorithmic): forward 16 (traced colour + alpha), backward 16 read + 16 written; per point 56 B in /
64 B out, amortised over S.  The torch arm runs the restatement of oracle/shading.py on the GPU (this script is test infrastructure: it lives under tests/ because only tests may import oracle/; the restatement stands for
"how the reference evaluates it" and is never used by the product).
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 14)
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--profile", action="store_true",
                    help="one forward + backward between cudaProfilerStart/Stop (for ncu --profile-from-start off), no timing")
    a = ap.parse_args()
    from irgs_b200 import shading
    from irgs_b200.incident import incident_dirs
    from oracle import shading as osh
    dev = torch.device("cuda", 0)
    P, S = a.points, a.spp
    g = torch.Generator(dev).manual_seed(1)
    r = lambda *s: torch.rand(*s, device=dev, generator=g)                    # noqa: E731
    nrm = torch.nn.functional.normalize(torch.randn(P, 3, device=dev, generator=g), dim=-1).requires_grad_(True)
    view = torch.nn.functional.normalize(nrm.detach() + 0.5 * torch.randn(P, 3, device=dev, generator=g), dim=-1)
    bc, ro, az = r(P, 3).requires_grad_(True), (0.05 + 0.9 * r(P, 1)).requires_grad_(True), r(P) * 2 * math.pi
    color = (r(P, S, 3) * (r(P, S, 1) < 0.5)).requires_grad_(True)
    alpha = ((r(P, S) * 1.1).clamp(0, 0.999) * (r(P, S) < 0.6)).requires_grad_(True)
    env = (torch.randn(256, 512, 3, device=dev, generator=g) * 0.5).requires_grad_(True)
    w = [torch.randn(P, 3, device=dev, generator=g) for _ in range(3)]
    keys = ("diffuse", "specular", "light_direct")

    def ours(backward):
        out = shading.shade_incident(nrm, S, bc, ro, view, env, color, alpha, azimuth=az, transmittance_min=0.03)
        if backward:
            torch.autograd.backward([out[k] for k in keys], w)

    def eager(backward):
        d = incident_dirs(nrm.detach(), S, az)      # directions from the device function; the torch arm does not pay for them
        out = osh.rendering_equation(bc, ro, nrm, view, d, color, alpha, env, "exp", None, 0.03)
        if backward:
            torch.autograd.backward([out[k] for k in keys], w)

    def timeit(fn, backward):
        for _ in range(3):
            fn(backward)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn(backward)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    if a.profile:
        ours(True)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        ours(True)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    res = {"points": P, "spp": S, "rays": P * S}
    for name, fn in (("kernels", ours), ("torch_eager", eager)):
        f = timeit(fn, False)
        fb = timeit(fn, True)
        res[name] = {"forward_ms": f, "forward_backward_ms": fb, "forward_GBps_algorithmic": 16.0 * P * S / (f * 1e-3) / 1e9,
                     "fwd_bwd_rays_per_s": P * S / (fb * 1e-3)}
    res["speedup_forward"] = res["torch_eager"]["forward_ms"] / res["kernels"]["forward_ms"]
    res["speedup_forward_backward"] = res["torch_eager"]["forward_backward_ms"] / res["kernels"]["forward_backward_ms"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
