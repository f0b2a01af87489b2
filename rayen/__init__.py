"""Import shim: ``from rayen import constraints, constraint_module, utils`` (reference readme.md:39)
resolves to the B200-native implementation in ``rayen_b200``."""
from rayen_b200 import constraints, utils  # noqa: F401


def __getattr__(name):
    if name == "constraint_module":
        import importlib
        return importlib.import_module("rayen_b200.constraint_module")
    raise AttributeError(name)
