"""CPU oracle of the IRGS surfel tracer math -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
irgs_b200/ (the product) never does.  The arithmetic lives in surfel_oracle.c (gcc, OpenMP), a restatement of
/root/reference/submodules/surfel_tracer/src/optix/gaussiantrace_{forward,backward}.cu and auxiliary.h; this file
is only the numpy/ctypes shim around it.
"""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_DIR, "liboracle.so")
    src = os.path.join(_DIR, "surfel_oracle.c")
    def stale():
        return not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so))
    if force or stale():
        import fcntl
        with open(so + ".lock", "w") as lock:   # several processes may import the checker at once: one of them builds
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or stale():
                    subprocess.check_call(["make", "-s", "-C", _DIR, "liboracle.so"])
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_lbvh_build.restype = ctypes.c_void_p
        _LIB.oracle_lbvh_free.argtypes = [ctypes.c_void_p]
    return _LIB


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def set_variant(k=16, opacity_cull=False, t_min=0.03):
    """Traversal-design experiments on the CPU (never changes results; the canonical counters use the default)."""
    lib().oracle_set_variant(ctypes.c_int(k), ctypes.c_int(int(opacity_cull)), ctypes.c_float(t_min))


def num_threads():
    return int(lib().oracle_num_threads())


class Scene:
    """Contiguous float32 copies of the per-surfel arrays in the tracer's layouts (SURVEY.md 8a')."""

    def __init__(self, means, opacity, ru, rv, normals, shs, features=None):
        self.means = _f32(means).reshape(-1, 3)
        self.n = self.means.shape[0]
        self.opacity = _f32(opacity).reshape(self.n)
        self.ru = _f32(ru).reshape(self.n, 3)
        self.rv = _f32(rv).reshape(self.n, 3)
        self.normals = _f32(normals).reshape(self.n, 3)
        self.shs = _f32(shs).reshape(self.n, -1, 3)
        self.K = self.shs.shape[1]
        self.features = _f32(features).reshape(self.n, -1) if features is not None else np.zeros((self.n, 0), np.float32)
        self.S = self.features.shape[1]
        self._bvh = None
        self._bvh_alpha_min = None

    def bvh(self, alpha_min):
        if self._bvh is None or self._bvh_alpha_min != alpha_min:
            self.free()
            self._bvh = lib().oracle_lbvh_build(ctypes.c_int(self.n), _p(self.means), _p(self.opacity), _p(self.ru),
                                                _p(self.rv), _p(self.normals), ctypes.c_float(alpha_min))
            self._bvh_alpha_min = alpha_min
        return self._bvh

    def free(self):
        if self._bvh is not None:
            lib().oracle_lbvh_free(ctypes.c_void_p(self._bvh))
            self._bvh = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def boxes(self, alpha_min):
        out = np.empty((self.n, 6), np.float32)
        lib().oracle_surfel_boxes(ctypes.c_int(self.n), _p(self.means), _p(self.opacity), _p(self.ru), _p(self.rv),
                                  _p(self.normals), ctypes.c_float(alpha_min), _p(out))
        return out


def trace_forward(scene, rays_o, rays_d, alpha_min=1.0 / 255.0, T_min=0.03, deg=3, back_culling=False,
                  use_bvh=False, hit_cap=64):
    """Returns dict(color, normal, feature, depth, alpha, hit_count, hits, margin, counters)."""
    o = _f32(rays_o).reshape(-1, 3)
    d = _f32(rays_d).reshape(-1, 3)
    R = o.shape[0]
    out = dict(color=np.zeros((R, 3), np.float32), normal=np.zeros((R, 3), np.float32),
               feature=np.zeros((R, scene.S), np.float32), depth=np.zeros(R, np.float32),
               alpha=np.zeros(R, np.float32), hit_count=np.zeros(R, np.int32),
               hits=np.full((R, hit_cap), -1, np.int32), margin=np.zeros((R, 3), np.float32),
               counters=np.zeros(3, np.int64))
    bvh = scene.bvh(alpha_min) if use_bvh else None
    rc = lib().oracle_trace_forward(
        ctypes.c_int64(R), ctypes.c_int(scene.n), ctypes.c_int(scene.S), ctypes.c_int(scene.K), ctypes.c_int(deg),
        ctypes.c_int(int(back_culling)), ctypes.c_float(alpha_min), ctypes.c_float(T_min), _p(o), _p(d),
        _p(scene.means), _p(scene.opacity), _p(scene.ru), _p(scene.rv), _p(scene.normals), _p(scene.features),
        _p(scene.shs), _p(out["color"]), _p(out["normal"]), _p(out["feature"]), _p(out["depth"]), _p(out["alpha"]),
        _p(out["hit_count"]), _p(out["hits"]), ctypes.c_int(hit_cap), _p(out["margin"]), _p(out["counters"]),
        ctypes.c_void_p(bvh))
    if rc != 0:
        raise ValueError("oracle_trace_forward: bad arguments (S>12, deg not in 0..3 or K<(deg+1)^2)")
    return out


def trace_backward(scene, rays_o, rays_d, fwd, gout, alpha_min=1.0 / 255.0, T_min=0.03, deg=3, back_culling=False,
                   use_bvh=False):
    """fwd: dict from trace_forward; gout: dict(color, normal, feature, depth, alpha) of incoming grads.
    Returns dict(rays_o, rays_d, means, opacity, ru, rv, normals, features, shs)."""
    o = _f32(rays_o).reshape(-1, 3)
    d = _f32(rays_d).reshape(-1, 3)
    R = o.shape[0]
    g = {k: _f32(gout[k]) for k in ("color", "normal", "feature", "depth", "alpha")}
    g["feature"] = g["feature"].reshape(R, scene.S)
    res = dict(rays_o=np.zeros((R, 3), np.float32), rays_d=np.zeros((R, 3), np.float32),
               means=np.zeros((scene.n, 3), np.float32), opacity=np.zeros(scene.n, np.float32),
               ru=np.zeros((scene.n, 3), np.float32), rv=np.zeros((scene.n, 3), np.float32),
               normals=np.zeros((scene.n, 3), np.float32), features=np.zeros((scene.n, scene.S), np.float32),
               shs=np.zeros((scene.n, scene.K, 3), np.float32))
    bvh = scene.bvh(alpha_min) if use_bvh else None
    rc = lib().oracle_trace_backward(
        ctypes.c_int64(R), ctypes.c_int(scene.n), ctypes.c_int(scene.S), ctypes.c_int(scene.K), ctypes.c_int(deg),
        ctypes.c_int(int(back_culling)), ctypes.c_float(alpha_min), ctypes.c_float(T_min), _p(o), _p(d),
        _p(scene.means), _p(scene.opacity), _p(scene.ru), _p(scene.rv), _p(scene.normals), _p(scene.features),
        _p(scene.shs), _p(_f32(fwd["color"])), _p(_f32(fwd["normal"])), _p(_f32(fwd["feature"])),
        _p(_f32(fwd["depth"])), _p(_f32(fwd["alpha"])), _p(g["color"]), _p(g["normal"]), _p(g["feature"]),
        _p(g["depth"]), _p(g["alpha"]), _p(res["rays_o"]), _p(res["rays_d"]), _p(res["means"]), _p(res["opacity"]),
        _p(res["ru"]), _p(res["rv"]), _p(res["normals"]), _p(res["features"]), _p(res["shs"]), ctypes.c_void_p(bvh))
    if rc != 0:
        raise ValueError("oracle_trace_backward: bad arguments")
    return res
