"""Loader for the UNMODIFIED reference package (leggedrobotics/rayen) -- test infrastructure only.

This file is part of the oracle (test infrastructure).  It is used
  (a) in the build container, where the reference tree is mounted read-only at /root/reference, to validate the
      restatement in oracle/rayen_oracle.py against the real reference and to generate the committed golden vectors
      under tests/golden/ (tests/golden/make_golden.py);
  (b) by bench.py's CPU-baseline legs (`cpu_baseline`, `--impl reference`), which time the reference's own code on the
      host cores from the verbatim copy oracle/make_ref.py leaves in oracle/_ref/ (git-ignored; /root/reference does not
      exist on the GPU box).
It is never imported by the product package (rayen_b200/), by `-m gpu` tests or by __graft_entry__.smoke().

The reference imports four third-party modules that are absent here (cvxpy, cvxpylayers, cdd,
colorama; rayen/constraints.py:7, rayen/constraint_module.py:10-12, rayen/utils.py:5-8).  None is
touched by method='RAYEN' forward/backward when the caller passes an explicit strictly interior
y0 and do_preprocessing_linear=False (constraints.py:224, :256, :412 are all skipped), so they are
replaced by inert stubs injected into sys.modules.  The reference source itself is imported as is,
under the module name `rayen_reference` so that it cannot collide with this repo's own `rayen` shim.
"""
import importlib.util
import os
import sys
import types

def _find_root():
    """The mounted reference tree (build container), else the verbatim copy oracle/make_ref.py left in oracle/_ref/
    (git-ignored; it travels to the GPU box with the snapshot so that bench.py can time the reference itself there)."""
    here = os.path.dirname(os.path.abspath(__file__))
    for root in (os.environ.get("RAYEN_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "_ref")):
        if root and os.path.isfile(os.path.join(root, "rayen", "constraint_module.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rayen", "constraint_module.py"))


class _Sym:
    """Inert stand-in for any cvxpy expression / variable / constraint."""
    __array_ufunc__ = None  # make numpy defer to our reflected operators
    value = None

    def __init__(self, *a, **k):
        pass

    def _same(self, *a, **k):
        return self

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = _same
    __matmul__ = __rmatmul__ = __neg__ = __truediv__ = _same
    __le__ = __ge__ = __eq__ = __rshift__ = __lshift__ = __getitem__ = _same
    __hash__ = object.__hash__

    @property
    def T(self):
        return self


class _Problem:
    status = "stub"

    def __init__(self, *a, **k):
        pass

    def solve(self, *a, **k):
        raise RuntimeError("cvxpy is stubbed in this container: pass y0 and do_preprocessing_linear=False")

    def is_dpp(self):
        return True


def _install_stubs():
    if "cvxpy" not in sys.modules:
        cp = types.ModuleType("cvxpy")
        cp.installed_solvers = lambda: ["SCS"]
        for name in ("Variable", "Parameter", "Minimize", "Maximize", "sum_squares", "quad_form", "norm"):
            setattr(cp, name, _Sym)
        cp.Problem = _Problem
        sys.modules["cvxpy"] = cp
    if "cvxpylayers" not in sys.modules:
        pkg = types.ModuleType("cvxpylayers")
        sub = types.ModuleType("cvxpylayers.torch")
        sub.CvxpyLayer = object
        pkg.torch = sub
        sys.modules["cvxpylayers"] = pkg
        sys.modules["cvxpylayers.torch"] = sub
    if "cdd" not in sys.modules:
        sys.modules["cdd"] = types.ModuleType("cdd")
    if "colorama" not in sys.modules:
        col = types.ModuleType("colorama")

        class _Blank:
            def __getattr__(self, _):
                return ""

        col.Fore = col.Back = col.Style = _Blank()
        sys.modules["colorama"] = col


_cached = None


def load_reference():
    """Return the reference package (attributes .constraints, .constraint_module, .utils)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_stubs()
    pkg_dir = os.path.join(REFERENCE_ROOT, "rayen")
    init = os.path.join(pkg_dir, "__init__.py")
    name = "rayen_reference"
    if os.path.isfile(init):
        spec = importlib.util.spec_from_file_location(name, init, submodule_search_locations=[pkg_dir])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules[name] = pkg
        spec.loader.exec_module(pkg)
    else:  # namespace-style package (the reference ships no __init__.py)
        pkg = types.ModuleType(name)
        pkg.__path__ = [pkg_dir]
        sys.modules[name] = pkg
    for sub in ("utils", "constraints", "constraint_module"):
        full = f"{name}.{sub}"
        spec = importlib.util.spec_from_file_location(full, os.path.join(pkg_dir, sub + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, sub, mod)
    _cached = pkg
    return pkg
