"""Generate golden vectors from the UNMODIFIED reference tracer (OptiX) -- run on the GPU box:

    python oracle/gen_golden_ref.py gpurun_out/golden        # then copy *.npz into tests/golden/

Needs baseline/_ref (oracle/build_ref.sh), a GPU, and libnvoptix.so.1 from the driver.  The scene is regenerated
from seeds by the tests; the file stores the rays, the reference's outputs and gradients, and checksums of the
scene arrays.  Test infrastructure only.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))  # the reference's own `surfel_tracer` package
sys.path.insert(1, ROOT)

CASES = {
    # name: scene kwargs, ray spec, features
    "c1_primary": dict(n=4000, scale_mult=2.5, n_features=0, rays="primary", hw=64, deg=3, back_culling=False),
    "secondary_feat": dict(n=4000, scale_mult=3.0, n_features=4, rays="secondary", pix=24, S=96, deg=3, back_culling=False),
    "deg1_cull": dict(n=4000, scale_mult=2.5, n_features=2, rays="secondary", pix=16, S=64, deg=1, back_culling=True),
    # the hot-path population: 100k surfels at the hit density of the 300k scene (x sqrt(3) in size), secondary rays that
    # run through >= 16 proxy crossings, i.e. several reference chunks
    "dense_100k": dict(n=100000, scale_mult=1.7320508, n_features=0, rays="secondary", pix=12, S=128, deg=3, back_culling=False),
}


def checksum(inp):
    h = hashlib.sha256()
    for k in ("means3D", "opacity", "ru", "rv", "normals", "features", "shs"):
        h.update(np.ascontiguousarray(inp[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def make_case(cfg):
    from irgs_b200 import synth
    sc = synth.make_scene(cfg["n"], n_features=cfg["n_features"], scale_mult=cfg["scale_mult"])
    inp = synth.derive_tracer_inputs(sc, synth.CAMERA_CENTER)
    if cfg["rays"] == "primary":
        o, d = synth.primary_rays(cfg["hw"], cfg["hw"])
    else:
        g = torch.Generator().manual_seed(synth.RAY_SEED)
        idx = torch.randint(0, cfg["n"], (cfg["pix"],), generator=g)
        pts = inp["means3D"][idx] + 0.01 * inp["normals"][idx]
        o, d = synth.secondary_rays(pts, inp["normals"][idx], cfg["S"])
        o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    R = o.shape[0]
    g = torch.Generator().manual_seed(synth.GRAD_SEED)
    gout = dict(color=torch.randn(R, 3, generator=g), normal=torch.randn(R, 3, generator=g),
                feature=torch.randn(R, cfg["n_features"], generator=g), depth=torch.randn(R, generator=g),
                alpha=torch.randn(R, generator=g))
    return sc, inp, o.contiguous(), d.contiguous(), gout


def proxy_crossings(inp, o, d, alpha_min):
    """Upper bound of the number of proxy polygons each ray crosses in (0, 100): plane hits inside the ellipse
    scaled to the proxy's circumradius (1.2584 x the alpha_min radius).  A ray with fewer than 16 of them is
    guaranteed to have been handled by the reference in a single 16-hit chunk."""
    mu, n = inp["means3D"].double().numpy(), inp["normals"].double().numpy()
    ru, rv, op = inp["ru"].double().numpy(), inp["rv"].double().numpy(), inp["opacity"].double().numpy().reshape(-1)
    r2 = 2 * np.log(np.maximum(op / alpha_min, 1e-30)) * 1.2584 ** 2 * 1.0001
    out = np.zeros(o.shape[0], np.int64)
    oo, dd = o.double().numpy(), d.double().numpy()
    if mu.shape[0] > 20000:
        # large scenes: only surfels whose bounding sphere (proxy circumradius) comes within reach of the ray can be crossed
        rad = np.sqrt(r2 / np.minimum((ru ** 2).sum(-1), (rv ** 2).sum(-1)))
    for i in range(o.shape[0]):
        if mu.shape[0] > 20000:
            rel = mu - oo[i]
            tc = rel @ dd[i]
            near = ((rel - tc[:, None] * dd[i]) ** 2).sum(-1) <= (rad * 1.01) ** 2
            sub = np.nonzero(near)[0]
            rel = oo[i] - mu[sub]
            og, dg = (n[sub] * rel).sum(-1), n[sub] @ dd[i]
            with np.errstate(divide="ignore", invalid="ignore"):
                t = -og / dg
            pos = rel + t[:, None] * dd[i]
            q = (ru[sub] * pos).sum(-1) ** 2 + (rv[sub] * pos).sum(-1) ** 2
            out[i] = np.count_nonzero((t > 0) & (t < 100) & (q <= r2[sub]) & (op[sub] > alpha_min))
            continue
        rel = oo[i] - mu
        og, dg = (n * rel).sum(-1), n @ dd[i]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = -og / dg
        pos = rel + t[:, None] * dd[i]
        q = (ru * pos).sum(-1) ** 2 + (rv * pos).sum(-1) ** 2
        out[i] = np.count_nonzero((t > 0) & (t < 100) & (q <= r2) & (op > alpha_min))
    return out


def strict_mask(inp, o, d, cfg):
    """Rays on which the reference and the restated math must agree to 1e-4: a single reference chunk, no near tie
    between consecutive hits, no alpha / transmittance within 2e-4 (relative) of its threshold."""
    import oracle
    from irgs_b200 import synth
    S = oracle.Scene(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], inp["shs"], inp["features"])
    fwd = oracle.trace_forward(S, o, d, alpha_min=synth.ALPHA_MIN, T_min=synth.T_MIN, deg=cfg["deg"],
                               back_culling=cfg["back_culling"])
    m = fwd["margin"]
    return (proxy_crossings(inp, o, d, synth.ALPHA_MIN) < 16) & (m[:, 0] > 2e-4) & (m[:, 1] > 2e-4) & (m[:, 2] > 2e-5)


def main(outdir, only=None):
    from surfel_tracer import GaussianTracer  # reference
    import surfel_tracer
    assert "baseline/_ref" in surfel_tracer.__file__, surfel_tracer.__file__
    from irgs_b200 import synth
    os.makedirs(outdir, exist_ok=True)
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        sc, inp, o, d, gout = make_case(cfg)
        dev = "cuda"
        tracer = GaussianTracer(transmittance_min=synth.T_MIN)
        vb, fb, gid = synth.proxy_mesh({k: v.to(dev) for k, v in sc.items()}, synth.ALPHA_MIN)
        tracer.build_bvh(vb, fb, gid)
        leaf = {k: inp[k].to(dev).clone().requires_grad_(True) for k in ("means3D", "opacity", "ru", "rv", "normals", "features", "shs")}
        ro, rd = o.to(dev).requires_grad_(True), d.to(dev).requires_grad_(True)
        feats = leaf["features"] if cfg["n_features"] > 0 else None
        outs = tracer.trace(ro, rd, leaf["means3D"], leaf["opacity"], leaf["ru"], leaf["rv"], leaf["normals"], feats,
                            leaf["shs"], alpha_min=synth.ALPHA_MIN, deg=cfg["deg"], back_culling=cfg["back_culling"])
        names = ("color", "normal", "feature", "depth", "alpha")
        strict = strict_mask(inp, o, d, cfg)
        res = {"out_" + n: t.detach().cpu().numpy() for n, t in zip(names, outs)}
        res["strict"] = strict
        # two backward passes: all rays (informational) and the strictly comparable rays only (asserted)
        for tag, mask in (("grad_", None), ("gstrict_", torch.from_numpy(strict).to(dev))):
            for t in list(leaf.values()) + [ro, rd]:
                t.grad = None
            loss = 0
            for n, t in zip(names, outs):
                if t.numel() == 0:
                    continue
                gw = gout[n].to(dev)
                if mask is not None:
                    gw = gw * (mask[:, None] if gw.dim() == 2 else mask)
                loss = loss + (t * gw).sum()
            loss.backward(retain_graph=True)
            torch.cuda.synchronize()
            res[tag + "rays_o"], res[tag + "rays_d"] = ro.grad.cpu().numpy(), rd.grad.cpu().numpy()
            for k in ("means3D", "opacity", "ru", "rv", "normals", "shs"):
                res[tag + k] = leaf[k].grad.cpu().numpy()
            res[tag + "features"] = (leaf["features"].grad.cpu().numpy() if leaf["features"].grad is not None
                                     else np.zeros((cfg["n"], cfg["n_features"]), np.float32))
        res.update(rays_o=o.numpy(), rays_d=d.numpy())
        meta = dict(cfg, checksum=checksum(inp), alpha_min=synth.ALPHA_MIN, T_min=synth.T_MIN,
                    torch=torch.__version__, gpu=torch.cuda.get_device_name(0),
                    source="reference surfel_tracer (OptiX) via baseline/_ref, unmodified")
        res["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(outdir, f"ref_optix_{name}.npz")
        np.savez_compressed(path, **res)
        print(name, "rays", o.shape[0], "alpha mean", float(res["out_alpha"].mean()), "->", path, os.path.getsize(path))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"), only=sys.argv[2:] or None)
