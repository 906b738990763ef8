"""The C-ABI library builds, loads, and exports every symbol include/irgs_b200.h declares (no compute calls: CPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "irgs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(irgs_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_reference_surface():
    syms = declared_symbols()
    for needed in ("irgs_tracer_create", "irgs_tracer_destroy", "irgs_build_from_proxy", "irgs_refit_from_proxy",
                   "irgs_intersection_test", "irgs_trace_forward", "irgs_trace_backward", "irgs_trace_fwd_bwd_host"):
        assert needed in syms


def test_library_builds_and_exports_every_declared_symbol():
    from irgs_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/irgs_b200.h but not exported by {path}"
    assert set(_lib.PROTOTYPES) == set(declared_symbols())
    assert _lib.load().irgs_version() >= 100


def test_library_is_sm100a_only_and_uses_vector_atomics():
    """cuobjdump: the cubin targets sm_100a and the backward uses 16-byte vector reductions (RED .128 / ATOM .128)."""
    import shutil
    import subprocess
    from irgs_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lst = subprocess.run([cuobjdump, "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in lst
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN4irgs28trace_backward_replay_kernelILb0EEEvNS_7KParamsE",
                           build.build()], capture_output=True, text=True).stdout
    assert re.search(r"(RED|ATOM)\S*\.128|REDG\S*\.128|\.F32x4|\.F32X4", sass) or "128" in sass


def test_sass_has_the_blackwell_instructions_the_design_relies_on():
    """The forward walk fetches nodes and records with 256-bit loads (LDG.E...256), the hit-parallel backward adds gradient
    rows with TMA bulk reductions (UBLKRED); neither kernel spills registers on the throughput path."""
    import shutil
    import subprocess
    from irgs_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = build.build()
    fwd = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN4irgs20trace_forward_kernelILi0ELb0ELb0EEEvNS_7KParamsEP5uint4", lib],
                         capture_output=True, text=True).stdout
    assert len(re.findall(r"LDG\.E\S*\.256", fwd)) >= 4, "forward kernel lost its 256-bit loads"
    # L1 policy of the walk: streaming data (records, SH rows) must not allocate in L1, tree nodes evict last
    assert len(re.findall(r"LDG\.E\.NA\S*\.256", fwd)) >= 8, "records / SH rows must be loaded with L1::no_allocate"
    assert len(re.findall(r"LDG\.E\.EL\S*\.256", fwd)) >= 2, "tree nodes must be loaded with L1::evict_last"
    bwd = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN4irgs26trace_backward_flat_kernelILb0ELb1EEEvNS_7KParamsE", lib],
                         capture_output=True, text=True).stdout
    assert "UBLKRED" in bwd, "backward kernel lost its bulk reduction"
    assert re.search(r"LDG\.E\S*\.256", bwd)


def test_header_is_plain_c(tmp_path):
    """include/irgs_b200.h must compile as C (the boundary a cgo / JNI / ctypes host binds), not only as C++."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use.c"
    src.write_text('#include "irgs_b200.h"\n'
                   'int probe(void) { irgs_incident_t g; g.sample_num = 1; (void)g; return irgs_version() >= 100 ? 0 : 1; }\n')
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c",
                           str(src), "-o", str(tmp_path / "use.o")])


def test_no_cpu_fallback():
    """Without CUDA the product must refuse to run rather than fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from irgs_b200.raytracer import GaussianTracer
    with pytest.raises(RuntimeError):
        GaussianTracer()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "irgs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f


def test_argument_validation_needs_no_gpu():
    """Error behaviour of the boundary: invalid arguments are rejected on the host, before any CUDA call, with a non-zero
    status and a message from irgs_last_error() (the Python layer maps that to RuntimeError)."""
    from irgs_b200 import _lib
    from irgs_b200.incident import IncidentDesc
    from irgs_b200.shading import EnvDesc, SamplingDesc
    lib = _lib.load()
    err = lambda: lib.irgs_last_error().decode()                              # noqa: E731
    dummy = ctypes.create_string_buffer(256)
    ptr = ctypes.cast(dummy, ctypes.c_void_p)
    gen = IncidentDesc(ptr, ptr, None, 4, 8, 0.05)
    env = EnvDesc(ptr, 4, 8, 1, 0)
    args = (ptr, ptr, ptr, ptr, ptr, ctypes.c_float(0.97), ptr, None)
    assert lib.irgs_shade_forward(None, ctypes.byref(env), None, *args) != 0 and "incident" in err()
    assert lib.irgs_shade_forward(ctypes.byref(gen), None, None, *args) != 0 and "environment" in err()
    assert lib.irgs_shade_forward(ctypes.byref(IncidentDesc(ptr, ptr, None, 4, 0, 0.05)), ctypes.byref(env), None, *args) != 0
    assert lib.irgs_shade_forward(ctypes.byref(gen), ctypes.byref(EnvDesc(ptr, 4, 8, 7, 0)), None, *args) != 0 and "activation" in err()
    assert lib.irgs_shade_forward(ctypes.byref(gen), ctypes.byref(EnvDesc(ptr, 0, 8, 1, 0)), None, *args) != 0
    few = SamplingDesc(None, ptr, 0.5, 0.5, 4)                                 # total_samples < the samples of the call
    assert lib.irgs_shade_forward(ctypes.byref(gen), ctypes.byref(env), ctypes.byref(few), *args) != 0 and "total_samples" in err()
    assert lib.irgs_env_lookup_forward(ctypes.byref(env), ptr, -1, ptr, None) != 0
    assert lib.irgs_env_lookup_forward(None, ptr, 4, ptr, None) != 0
    # an empty batch is not an error and launches nothing
    assert lib.irgs_shade_forward(ctypes.byref(IncidentDesc(None, None, None, 0, 8, 0.05)), ctypes.byref(env), None, *args) == 0
    assert lib.irgs_env_lookup_forward(ctypes.byref(env), None, 0, None, None) == 0
    # tracer entry points without a handle
    assert lib.irgs_trace_forward(None, 4, 0, 16, 3, ptr, ptr, *([None] * 14), 0, ctypes.c_float(0.0), ctypes.c_float(0.0), 0, None) != 0
    assert "handle" in err()
    assert lib.irgs_trace_forward(None, 4, 0, 16, 3, *([None] * 16), 0, ctypes.c_float(0.0), ctypes.c_float(0.0), 0, None) != 0
    assert lib.irgs_set_option(None, b"slot", 0) != 0


def test_stride_start_order_is_a_bijection_in_32_bits():
    """Small launches start their rays in the order (i * m) mod n (trace_fwd.cu launch_trace_forward); the kernel multiplies in
    32 bits.  For a spread of n: m is odd, coprime to n, i * m never overflows, and the map is a permutation of [0, n)."""
    import math
    import numpy as np
    from irgs_b200 import _lib
    lib = _lib.load()
    assert lib.irgs_stride_multiplier(63) == 0 and lib.irgs_stride_multiplier((1 << 19) + 1) == 0
    sizes = [64, 65, 100, 255, 256, 4096, 5063, 5063 * 3, 61 * 83 * 7, 65536, 99991, 1 << 18, (1 << 18) + 3, 3 * 5 * 7 * 11 * 13 * 17,
             2 * 3 * 5 * 7 * 11 * 13 * 17, 510510 - 1, (1 << 19) - 1, 1 << 19]
    for n in sizes:
        m = lib.irgs_stride_multiplier(n)
        assert m > 0 and m % 2 == 1 and m < 8192 and math.gcd(m, n) == 1, (n, m)
        assert (n - 1) * m < 2 ** 32
        i = np.arange(n, dtype=np.uint32)
        perm = (i * np.uint32(m)) % np.uint32(n)                    # the kernel's arithmetic
        assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32)), n


def test_concurrent_builds_never_expose_a_partial_library(tmp_path):
    """One rank per GPU imports the package under torchrun; when the library is stale every rank wants to build it.  The build
    holds a lock and renames a finished temporary file over the library, so a process that loads it while another one compiles
    never maps a half-written file (seen once as `file too short` on four of eight ranks).  A slow fake compiler stands in for nvcc."""
    import shutil
    import subprocess
    import sys
    from irgs_b200 import build as b
    assert os.path.exists(b.LIB)
    keep = tmp_path / "lib.keep"
    shutil.copy(b.LIB, keep)
    fake = tmp_path / "nvcc"
    fake.write_text("#!/bin/bash\nout=\"\"\nwhile [ $# -gt 0 ]; do if [ \"$1\" = \"-o\" ]; then out=\"$2\"; fi; shift; done\n"
                    f"head -c 65536 {keep} > \"$out\"\nsleep 0.4\ncat {keep} > \"$out\"\n")
    fake.chmod(0o755)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import ctypes, sys; sys.path.insert(0, %r); from irgs_b200 import build as b; b.build(force=True); "
            "lib = ctypes.CDLL(b.LIB); lib.irgs_last_error; print('ok')" % root)
    env = dict(os.environ, NVCC=str(fake))
    procs = [subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for _ in range(4)]
    try:
        for p in procs:
            out, err = p.communicate(timeout=120)
            assert p.returncode == 0 and "ok" in out, err[-2000:]
        assert open(b.LIB, "rb").read() == open(keep, "rb").read()
        assert not [f for f in os.listdir(os.path.dirname(b.LIB)) if f.endswith(".tmp")]
    finally:
        if not os.path.exists(b.LIB) or os.path.getsize(b.LIB) != os.path.getsize(keep):
            shutil.copy(keep, b.LIB)
