"""The oracle (oracle/rayen_oracle.py) pinned against the golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py) and, when the reference tree is mounted, against the reference live."""
import numpy as np
import pytest
import torch

from helpers import golden_names, load_golden
from oracle.rayen_oracle import OracleSet, TorchOracle, closed_form_numpy, max_violation
from oracle.reference_loader import load_reference, reference_available
from rayen_b200 import synthetic


def _finite_rows(g):
    return np.isfinite(g["gv64"]).all(axis=1) & np.isfinite(g["gv32"]).all(axis=1)


@pytest.mark.parametrize("name", golden_names())
def test_torch_oracle_matches_reference_golden(name):
    g = load_golden(name)
    cs = synthetic.build_constraints(g["spec"])
    oset = OracleSet.from_constraints(cs)
    fin = _finite_rows(g)
    v, gy = torch.tensor(g["v"]), torch.tensor(g["gy"])
    method = "RAYEN_old" if name.startswith("old_") else "RAYEN"
    y64, gv64 = TorchOracle(oset, torch.float64).forward_backward(v.double(), gy.double(), method)
    scale_y, scale_g = np.abs(g["y64"]).max(), np.abs(g["gv64"][fin]).max()
    assert np.abs(y64.numpy() - g["y64"]).max() <= 1e-10 * scale_y
    assert np.abs(gv64.numpy() - g["gv64"])[fin].max() <= 1e-8 * scale_g
    y32, gv32 = TorchOracle(oset, torch.float32).forward_backward(v, gy, method)
    assert np.abs(y32.numpy() - g["y32"]).max() <= 5e-6 * scale_y
    assert np.abs(gv32.numpy() - g["gv32"])[fin].max() <= 5e-5 * scale_g


@pytest.mark.parametrize("name", [n for n in golden_names() if not n.startswith("old_")])
def test_closed_form_matches_reference_golden(name):
    """The analytic forward/backward equals the reference's autograd away from ties."""
    g = load_golden(name)
    cs = synthetic.build_constraints(g["spec"])
    cf = closed_form_numpy(OracleSet.from_constraints(cs), g["v"], g["gy"])
    assert np.abs(cf["y"] - g["y64"]).max() <= 1e-9 * np.abs(g["y64"]).max()
    ok = _finite_rows(g) & (cf["margin"] > 1e-6)
    assert ok.sum() >= 0.8 * len(ok)
    assert np.abs(cf["gv"] - g["gv64"])[ok].max() <= 1e-6 * np.abs(g["gv64"][ok]).max()


@pytest.mark.parametrize("name", golden_names())
def test_reference_outputs_are_feasible(name):
    g = load_golden(name)
    cs = synthetic.build_constraints(g["spec"])
    s = g["spec"]
    viol = max_violation(OracleSet.from_constraints(cs), g["y64"], s["A1"], s["b1"], s["A2"], s["b2"])
    assert viol <= 1e-9 * max(1.0, np.abs(g["y64"]).max())


def test_preprocessed_fields_match_reference():
    """A_p, b_p, NA_E, yp, z0, y0 of this package's ConvexConstraints equal the reference's (stored in the goldens)."""
    checked = 0
    for name in golden_names():
        g = load_golden(name)
        if not g["ref"]:
            continue
        cs = synthetic.build_constraints(g["spec"])
        for f, ref in g["ref"].items():
            np.testing.assert_allclose(getattr(cs, f), ref, rtol=0, atol=1e-12, err_msg=f"{name}.{f}")
        checked += 1
    assert checked >= 15


@pytest.mark.skipif(not reference_available(), reason="/root/reference is only mounted in the build container")
@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_matches_live_reference(seed):
    ref = load_reference()
    spec = synthetic.random_spec(k=6, m=10, eta=2, mu=2, r_M=5, r=5, seed=seed)
    spec["b1"] = spec["b1"] * 3
    v, gy = synthetic.sample_inputs(200, 6, 6, seed_v=seed + 11, dtype=torch.float64)
    torch.set_default_dtype(torch.float64)
    try:
        cs_ref = synthetic.build_constraints(spec, module=ref.constraints)
        layer = ref.constraint_module.ConstraintModule(cs_ref, method="RAYEN", create_map=False)
        x = v.reshape(200, 6, 1).clone().requires_grad_(True)
        y = layer(x)
        (y[:, :, 0] * gy).sum().backward()
    finally:
        torch.set_default_dtype(torch.float32)
    oset = OracleSet.from_constraints(synthetic.build_constraints(spec))
    y_o, g_o = TorchOracle(oset, torch.float64).forward_backward(v, gy)
    assert (y_o - y.detach()[:, :, 0]).abs().max() < 1e-12
    assert (g_o - x.grad[:, :, 0]).abs().max() < 1e-10
