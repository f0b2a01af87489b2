"""Shim for ``rayen.constraints`` -> ``rayen_b200.constraints``."""
from rayen_b200.constraints import *  # noqa: F401,F403
from rayen_b200.constraints import (ConvexConstraints, ConvexQuadraticConstraint, LinearConstraint,  # noqa: F401
                                    LMIConstraint, SOCConstraint)
