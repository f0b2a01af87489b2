"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every declared symbol; the
Python drop-in keeps the reference's API surface; nothing computes without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from rayen_b200 import _cabi, synthetic
from rayen_b200.constraint_module import ConstraintModule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rayen_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rayen_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    names = declared_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(built_library, name), f"{name} declared in include/rayen_b200.h but not exported"
    assert sorted(_cabi.SYMBOLS) == names          # the ctypes table covers the whole header
    assert built_library.rayen_abi_version() == _cabi.ABI_VERSION


def test_library_is_built_for_sm_100a_only(built_library):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_plan_create_without_gpu_returns_error_code(built_library):
    from rayen_b200 import plan
    packed = plan.build_plan_from_constraints(synthetic.build_constraints(synthetic.example_spec(0)))
    handle = ctypes.c_void_p()
    desc = packed.desc()
    rc = built_library.rayen_plan_create(ctypes.byref(desc), 0, ctypes.byref(handle))
    assert rc == -4 and not handle.value       # RAYEN_ERR_NO_DEVICE, never a crash
    assert b"not available" in built_library.rayen_last_error()
    desc.abi_version = 1
    assert built_library.rayen_plan_create(ctypes.byref(desc), 0, ctypes.byref(handle)) == -3
    assert built_library.rayen_plan_create(None, 0, ctypes.byref(handle)) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="not built"):
        _cabi.lib()


def test_module_api_surface_and_buffers():
    """Buffer names / shapes of the README set as the reference registers them."""
    cs = synthetic.build_constraints(synthetic.example_spec("readme"))
    layer = ConstraintModule(cs, input_dim=64, create_map=True)
    shapes = {k: tuple(v.shape) for k, v in layer.named_buffers()}
    assert shapes == {
        "mHinv": (2, 2), "L": (2, 2), "D": (6, 2), "all_P": (1, 3, 3), "all_q": (1, 3, 1), "all_r": (1, 1, 1),
        "all_M": (1, 3, 3), "all_s": (1, 3, 1), "all_c": (1, 3, 1), "all_d": (1, 1, 1), "all_F": (4, 2, 2),
        "A_p": (6, 2), "b_p": (6, 1), "yp": (3, 1), "NA_E": (3, 2), "z0": (2, 1), "y0": (3, 1),
        "all_delta": (1, 3, 3), "all_phi": (1, 1, 3)}
    assert (layer.k, layer.n, layer.getDimAfterMap(), layer.method) == (3, 2, 2, "RAYEN")
    assert isinstance(layer.mapper, torch.nn.Linear) and layer.mapper.out_features == 2
    np.testing.assert_allclose(layer.gety0().numpy(), cs.y0, atol=1e-6)
    z = torch.randn(5, 2, 1)
    torch.testing.assert_close(layer.getzFromy(layer.getyFromz(z)), z, atol=1e-5, rtol=0)
    assert ConstraintModule(cs, create_map=False, method="RAYEN_old").getDimAfterMap() == 3
    # empty families are shape-(0,) buffers like the reference
    lin_only = ConstraintModule(synthetic.build_constraints(synthetic.example_spec(0)), create_map=False)
    assert lin_only.all_P.shape == (0,) and lin_only.all_F.shape == (0,) and not hasattr(lin_only, "L")


def test_baseline_methods_are_out_of_scope():
    cs = synthetic.build_constraints(synthetic.example_spec(0))
    for method in ("UU", "Bar", "PP", "UP", "DC3"):
        with pytest.raises(NotImplementedError):
            ConstraintModule(cs, create_map=False, method=method)
    with pytest.raises(RuntimeError, match="input_dim"):
        ConstraintModule(cs, create_map=True)


def test_no_cpu_fallback():
    cs = synthetic.build_constraints(synthetic.example_spec(0))
    layer = ConstraintModule(cs, create_map=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.randn(4, 2, 1))


def test_state_dict_round_trip_repacks_the_plan():
    cs_a = synthetic.build_constraints(synthetic.example_spec(13))
    spec_b = synthetic.example_spec(13)
    spec_b["y0"] = np.array([[0.6], [0.0], [0.8]])
    cs_b = synthetic.build_constraints(spec_b)
    a, b = ConstraintModule(cs_a, create_map=False), ConstraintModule(cs_b, create_map=False)
    assert not np.allclose(a._packed.blob, b._packed.blob)
    b.load_state_dict(a.state_dict())
    np.testing.assert_allclose(b._packed.blob, a._packed.blob, rtol=2e-5, atol=2e-6)  # rebuilt from fp32 buffers


def test_wide_module_state_dict_round_trip_and_wide_descriptor():
    """A set with n > 32: same buffers as the reference, a wide plan (descriptor flag + WIDE section), and a reload from
    the float32 buffers of another module rebuilds a plan that maps the same inputs to the same outputs."""
    from rayen_b200 import plan as plan_mod
    cs_a = synthetic.build_constraints(synthetic.wide_spec(40, 30, 1, 1, 12, 2, seed=1))
    cs_b = synthetic.build_constraints(synthetic.wide_spec(40, 30, 1, 1, 12, 2, seed=2))
    a, b = ConstraintModule(cs_a, create_map=False), ConstraintModule(cs_b, create_map=False)
    assert (a.k, a.n) == (40, 38) and a.D.shape == (cs_a.A_p.shape[0], 38) and a.all_delta.shape == (1, 40, 40)
    desc = a._packed.desc()
    assert desc.wide == 1 and desc.off_wide > 0 and desc.np == 40 and desc.tc_panels == 0 and desc.lmi_r == 0
    v = np.random.default_rng(0).uniform(-2, 2, size=(32, 38))
    ya, ka, ta = plan_mod.evaluate_wide_numpy(a._packed, v)
    yb, _, _ = plan_mod.evaluate_wide_numpy(b._packed, v)
    assert np.abs(ya - yb).max() > 1e-3
    b.load_state_dict(a.state_dict())
    yb, kb, tb = plan_mod.evaluate_wide_numpy(b._packed, v)
    np.testing.assert_allclose(yb, ya, rtol=0, atol=2e-5 * np.abs(ya).max())
    assert (tb == ta).mean() > 0.9
    with pytest.raises(RuntimeError, match="CUDA"):
        a(torch.randn(4, 38, 1))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rayen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
    for f in os.listdir(os.path.join(ROOT, "rayen")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "rayen", f)).read()
