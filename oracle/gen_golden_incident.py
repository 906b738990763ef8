"""Records golden vectors of the reference's incident-ray sampling by running the UNMODIFIED functions of
/root/reference/utils/graphics_utils.py on the CPU of the build container (their hard-coded device='cuda' arguments
are dropped by a shim around torch.arange / rand / zeros / eye; nothing in the reference file is edited).

    python oracle/gen_golden_incident.py      ->  tests/golden/ref_incident.npz
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/utils/graphics_utils.py"

_rand_log = []


def _strip_device(fn, log=None):
    def wrapped(*a, **k):
        k.pop("device", None)
        out = fn(*a, **k)
        if log is not None:
            log.append(out.clone())
        return out
    return wrapped


def load_reference():
    for name in ("arange", "zeros", "eye"):
        setattr(torch, name, _strip_device(getattr(torch, name)))
    torch.rand = _strip_device(torch.rand, _rand_log)
    spec = importlib.util.spec_from_file_location("ref_graphics_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    g = torch.Generator().manual_seed(4321)
    n = torch.randn(61, 3, generator=g)
    n = n / n.norm(dim=-1, keepdim=True)
    special = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.0, -1.0, 0.0],
                            [6e-4, -8e-4, -0.9999995]])
    special = special / special.norm(dim=-1, keepdim=True)
    normals = torch.cat([n, special]).contiguous()
    out = {"normals": normals.numpy()}
    for S in (24, 256):
        d_eval, area = ref.fibonacci_sphere_sampling(normals, S, random_rotate=False)
        torch.manual_seed(99 + S)
        _rand_log.clear()
        d_train, _ = ref.fibonacci_sphere_sampling(normals, S, random_rotate=True)
        azim = (_rand_log[-1] * 2 * np.pi).reshape(-1)          # the reference's `rand * 2 * np.pi` term
        out[f"dirs_eval_{S}"] = d_eval.numpy()
        out[f"dirs_train_{S}"] = d_train.numpy()
        out[f"azimuth_{S}"] = azim.numpy()
        assert float(area.min()) == float(area.max())
        out[f"area_{S}"] = np.float32(area.flatten()[0].item())
    out["rotation"] = ref.rotation_between_z(normals).numpy()
    path = os.path.join(ROOT, "tests", "golden", "ref_incident.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
