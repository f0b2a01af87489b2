"""rayen_b200: the RAYEN feasibility layer (leggedrobotics/rayen) rebuilt for NVIDIA B200 (sm_100a).

Public surface = the reference's: ``constraints`` (value classes + ``ConvexConstraints``) and
``constraint_module.ConstraintModule``.  ``from rayen import constraints, constraint_module`` also
works through the ``rayen`` shim package at the repo root.
"""
from . import utils, constraints  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # constraint_module pulls in torch + the CUDA C-ABI binding; keep `import rayen_b200` light.
    if name in ("constraint_module", "plan", "synthetic", "sharding"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
