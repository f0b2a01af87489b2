"""Shim for ``rayen.utils`` -> ``rayen_b200.utils``."""
from rayen_b200.utils import *  # noqa: F401,F403
