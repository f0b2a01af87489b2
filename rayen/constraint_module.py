"""Shim for ``rayen.constraint_module`` -> ``rayen_b200.constraint_module``."""
from rayen_b200.constraint_module import ConstraintModule  # noqa: F401
