"""C4 (relighting evaluation shape) broken into its parts on a pixel subset: diffuse trace_incident, light-direction sampling,
light-ray trace."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from irgs_b200 import shading, synth
from irgs_b200.raytracer import GaussianTracer
dev = torch.device("cuda", 0)
args = argparse.Namespace(surfels=300000, img=256, spp=256)
def factory(sc, inp):
    tr = GaussianTracer(transmittance_min=synth.T_MIN, device=dev)
    tr.build_from_surfels(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], synth.ALPHA_MIN)
    return tr
sc, inp, tr, ro, rd = bench.build_workload(args, dev, 0, 1, factory)
pts, nrm, _ = bench.build_workload.points
P = pts.shape[0]
gen = torch.Generator(dev).manual_seed(31)
feats = torch.rand(inp["means3D"].shape[0], 4, device=dev, generator=gen)
env = shading.EnvLight(resolution=(256, 512), activation="exp", device=dev)
env.base.data += 0.5 * torch.randn(env.base.shape, device=dev, generator=gen)
env.update_pdf()
surf = (inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], feats, inp["shs"])
surf0 = surf[:5] + (None,) + surf[6:]
def t(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad(): out = fn()
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best, out
pc = 1 << 15
a, _ = t(lambda: tr.trace_incident(pts[:pc], nrm[:pc], 512, *surf, synth.ALPHA_MIN, t_min=synth.LIGHT_T_MIN))
a0, _ = t(lambda: tr.trace_incident(pts[:pc], nrm[:pc], 512, *surf0, synth.ALPHA_MIN, t_min=synth.LIGHT_T_MIN))
print(f"diffuse 512/pixel, {pc} pixels = {pc*512} rays: S=4 {a:.2f} ms ({pc*512/a/1e3:.0f} Mrays/s) | S=0 {a0:.2f} ms ({pc*512/a0/1e3:.0f} Mrays/s)")
b, dirs = t(lambda: env.sample_light_directions(pc, 256, False)[0])
print(f"sampling {pc*256} light directions: {b:.2f} ms")
c, org = t(lambda: pts[:pc, None] + dirs * synth.LIGHT_T_MIN)
print(f"origins: {c:.2f} ms")
d, _ = t(lambda: tr.trace(org, dirs, *surf, synth.ALPHA_MIN))
print(f"light rays {pc*256}: trace {d:.2f} ms ({pc*256/d/1e3:.0f} Mrays/s)  hits/ray {float(tr.last_hit_count.float().mean()):.2f}")
tr.trace_incident(pts[:pc], nrm[:pc], 512, *surf, synth.ALPHA_MIN, t_min=synth.LIGHT_T_MIN)
print(f"diffuse hits/ray {float(tr.last_hit_count.float().mean()):.2f}")
