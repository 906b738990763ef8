import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, oracle
from irgs_b200 import synth
from irgs_b200.raytracer import GaussianTracer
DEV = "cuda:0"
n = 300
g = torch.Generator().manual_seed(5)
means = torch.zeros(n, 3) + 1e-4 * torch.randn(n, 3, generator=g)
scale = 0.01 * 1.03 ** torch.arange(n, dtype=torch.float32)
nrm = torch.tensor([0.0, 0.0, 1.0]).expand(n, 3).contiguous()
ru = (torch.tensor([1.0, 0.0, 0.0]).expand(n, 3) / scale[:, None]).contiguous()
rv = (torch.tensor([0.0, 1.0, 0.0]).expand(n, 3) / scale[:, None]).contiguous()
inp = dict(means3D=means, opacity=torch.full((n, 1), 0.02), ru=ru, rv=rv, normals=nrm, features=torch.zeros(n, 0),
           shs=torch.randn(n, 16, 3, generator=g) * 0.2)
o = torch.tensor([[0.3, 0.1, 2.0], [5.0, 0.0, 3.0], [-0.02, 0.01, 1.0]]).repeat(20, 1) + 0.05 * torch.randn(60, 3, generator=g)
d = -o / o.norm(dim=1, keepdim=True)
S = oracle.Scene(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], inp["shs"], inp["features"])
ref = oracle.trace_forward(S, o, d, hit_cap=256)
gi = {k: v.to(DEV) for k, v in inp.items()}
for builder in (0, 1):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, hit_cap=256)
    tr.set_option("builder", builder)
    tr.build_from_surfels(gi["means3D"], gi["opacity"], gi["ru"], gi["rv"], gi["normals"], synth.ALPHA_MIN)
    res = tr.trace_with_hits(o.to(DEV), d.to(DEV), gi["means3D"], gi["opacity"], gi["ru"], gi["rv"], gi["normals"], gi["features"], gi["shs"], synth.ALPHA_MIN, hit_cap=256)
    err = np.abs(res["color"].cpu().numpy() - ref["color"]).max(1)
    bad = np.nonzero(err > 1e-4)[0]
    print("builder", builder, "bad rays", bad, err[bad])
    hits = res["hits"].cpu().numpy(); hc = res["hit_count"].cpu().numpy()
    for r in bad[:3]:
        a, b = hits[r, :hc[r]], ref["hits"][r, :ref["hit_count"][r]]
        print(" ray", r, "hc", hc[r], ref["hit_count"][r], "lists equal", np.array_equal(a, b), "first diff", (np.nonzero(a != b)[0][:5] if len(a) == len(b) else None), "alpha", res["alpha"][r].item(), ref["alpha"][r], "margin", ref["margin"][r])
        if len(a) == len(b) and not np.array_equal(a, b):
            i = np.nonzero(a != b)[0][0]; print("   around", a[max(0,i-2):i+4], b[max(0,i-2):i+4])
    for r in bad[:2]:
        a, b = list(hits[r, :hc[r]]), list(ref["hits"][r, :ref["hit_count"][r]])
        import collections
        dup = [k for k, v in collections.Counter(a).items() if v > 1]
        miss = [x for x in b if x not in a]
        extra = [x for x in a if x not in b]
        print("   dup", dup, "missing", miss, "extra", extra)
        for x in miss[:3]:
            print("    missing", x, "oracle pos", b.index(x))
        for x in dup[:3]:
            print("    dup", x, "gpu pos", [i for i, y in enumerate(a) if y == x], "oracle pos", b.index(x) if x in b else None)
        # depths of the oracle list around the first diff
        i = int(np.nonzero(np.array(a) != np.array(b))[0][0])
        mu = inp["means3D"].numpy(); oo = o[r].numpy(); dd = d[r].numpy()
        ts = [float(-(oo[2]-mu[x][2])/dd[2]) for x in b[i-3:i+4]]
        print("    oracle t around", ts)
