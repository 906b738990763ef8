"""Drop-in for IRGS's `surfel_tracer` package (reference: submodules/surfel_tracer/surfel_tracer/__init__.py, which
is the single line `from .raytracer import GaussianTracer`).  Put the repository root on PYTHONPATH and
`from surfel_tracer import GaussianTracer` (scene/gaussian_model.py:16) resolves to the B200-native tracer."""
from irgs_b200.raytracer import GaussianTracer  # noqa: F401
