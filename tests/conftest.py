import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def small_scene():
    """10k-surfel synthetic scene, surfels enlarged x4 so rays see the 300k scene's hit density (multi-pass)."""
    from irgs_b200 import synth
    sc = synth.make_scene(10000, n_features=4, scale_mult=4.0)
    inp = synth.derive_tracer_inputs(sc, synth.CAMERA_CENTER)
    return sc, inp
