"""Dump surfel boxes + a sample of C3 secondary rays (with the depth at which each ray's compositing ends) for
scripts/exp/exp_bvh.c.  CPU only (uses the oracle for the primary pass)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import oracle
from irgs_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
n_pix = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sc = synth.make_scene(N)
inp = synth.derive_tracer_inputs(sc, synth.CAMERA_CENTER)
S = oracle.Scene(*(inp[k] for k in ("means3D", "opacity", "ru", "rv", "normals", "shs")))
o, d = synth.primary_rays(800, 800)
g = torch.Generator().manual_seed(1)
sel = torch.randperm(640000, generator=g)[:n_pix * 4]
prim = oracle.trace_forward(S, o[sel], d[sel], use_bvh=True, hit_cap=4)
ok = torch.from_numpy(prim["alpha"] > 0.5)
pts, nrm = synth.shading_points_from_primary(o[sel], d[sel], torch.from_numpy(prim["depth"]), torch.from_numpy(prim["alpha"]), torch.from_numpy(prim["normal"]))
pts, nrm = pts[ok][:n_pix], nrm[ok][:n_pix]
ro, rd = synth.secondary_rays(pts, nrm, 256)
ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
f = oracle.trace_forward(S, ro, rd, use_bvh=True, hit_cap=64)
hc = f["hit_count"]; hits = f["hits"]
tclip = np.full(ro.shape[0], 100.0, np.float32)
term = f["alpha"] >= 1 - synth.T_MIN
idx = np.nonzero(term & (hc > 0) & (hc <= 64))[0]
last = hits[idx, hc[idx] - 1]
n = inp["normals"].numpy()[last]; mu = inp["means3D"].numpy()[last]
og = ((ro.numpy()[idx] - mu) * n).sum(1); dg = (rd.numpy()[idx] * n).sum(1)
tclip[idx] = -og * dg / np.maximum(1e-6, dg * dg) * 1.00001
print("rays", ro.shape[0], "hit frac", (hc > 0).mean(), "terminated frac", term.mean(), "counters/ray", f["counters"] / ro.shape[0])
S.boxes(synth.ALPHA_MIN).astype(np.float32).tofile("/tmp/exp_boxes.bin")
np.concatenate([ro.numpy(), rd.numpy(), tclip[:, None]], 1).astype(np.float32).tofile("/tmp/exp_rays.bin")
print(N, ro.shape[0])
