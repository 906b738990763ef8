"""Builds libirgs_b200.so (hand-written CUDA for sm_100a + the C ABI of include/irgs_b200.h) in-tree with nvcc.

No torch headers are involved: the library's boundary is plain C, and Python reaches it through ctypes
(irgs_b200/_lib.py).  nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
SOURCES = ["capi.cu", "lbvh.cu", "trace.cu", "trace_fwd.cu", "shade.cu", "surfel_params.cu"]
HEADERS = ["internal.cuh", "trace_common.cuh", "shade_math.cuh", os.path.join("..", "..", "include", "irgs_b200.h")]
LIB = os.path.join(CSRC, "libirgs_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(os.path.join(CSRC, f)) and os.path.getmtime(os.path.join(CSRC, f)) > t
               for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the library unless it is up to date.  Safe under several processes at once (one rank per GPU under torchrun, each
    importing the package): one process holds a lock file while it compiles into a temporary file that is renamed over the library
    atomically, the others wait for the lock and then find the library up to date -- nobody ever maps a half-written file."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():   # somebody else built it while this process waited
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            extra = os.environ.get("IRGS_NVCC_DEFS", "").split()  # tuning experiments only, e.g. "-DIRGS_MIN_ACTIVE=16"
            tmp = f"{LIB}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
            try:
                subprocess.check_call(cmd, cwd=CSRC)
                os.replace(tmp, LIB)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
