"""Constraint value classes and the solver-free host preprocessing (drop-in for reference constraints.py)."""
import numpy as np
import pytest

from rayen import constraints  # the drop-in import path of the reference (readme.md:39)
from rayen_b200 import synthetic


def test_linear_constraint_validation():
    A, b = np.eye(3), np.ones((3, 1))
    lc = constraints.LinearConstraint(A, b, None, None)
    assert lc.hasIneqConstraints() and not lc.hasEqConstraints() and lc.dim() == 3
    with pytest.raises(RuntimeError):
        constraints.LinearConstraint(None, None, None, None)
    with pytest.raises(RuntimeError):
        constraints.LinearConstraint(A, np.ones((2, 1)), None, None)
    with pytest.raises(RuntimeError):
        constraints.LinearConstraint(A, b, np.ones((1, 2)), np.ones((1, 1)))


def test_quadratic_soc_lmi_validation():
    with pytest.raises(RuntimeError):
        constraints.ConvexQuadraticConstraint(np.zeros((2, 2)), np.zeros((2, 1)), -np.ones((1, 1)))
    with pytest.raises(RuntimeError):
        constraints.ConvexQuadraticConstraint(np.array([[1.0, 2.0], [0.0, 1.0]]), np.zeros((2, 1)), -np.ones((1, 1)))
    with pytest.raises(RuntimeError):
        constraints.ConvexQuadraticConstraint(np.diag([1.0, -1.0]), np.zeros((2, 1)), -np.ones((1, 1)))
    qc = constraints.ConvexQuadraticConstraint(np.diag([1.0, -1e-9]), np.zeros((2, 1)), -np.ones((1, 1)))
    assert np.linalg.eigvalsh(qc.P).min() >= 0  # round-off repaired like the reference
    with pytest.raises(RuntimeError):
        constraints.SOCConstraint(np.zeros((2, 2)), np.zeros((2, 1)), np.ones((2, 1)), np.ones((1, 1)))
    with pytest.raises(RuntimeError):
        constraints.LMIConstraint([np.array([[0.0, 1.0], [0.0, 0.0]]), np.eye(2), np.eye(2)])
    assert constraints.LMIConstraint([np.eye(2), np.eye(2), np.eye(2)]).dim() == 2


def test_no_constraints_is_an_error():
    with pytest.raises(RuntimeError, match="no constraints"):
        constraints.ConvexConstraints()


@pytest.mark.parametrize("ex", synthetic.EXAMPLE_IDS)
def test_interior_point_and_subspace_without_y0(ex):
    """y0=None + do_preprocessing_linear=True (the README default): every canned set of the reference gets a
    strictly interior point, an orthonormal null-space basis and the right subspace dimension."""
    spec = synthetic.example_spec(ex)
    cs = synthetic.build_constraints(spec, y0=None, do_preprocessing_linear=True)
    assert cs.NA_E.shape == (cs.k, cs.n)
    np.testing.assert_allclose(cs.NA_E.T @ cs.NA_E, np.eye(cs.n), atol=1e-10)
    expected_n = {"readme": 2, 0: 2, 1: 2, 6: 1, 7: 2, 9: 2}.get(ex, cs.k)
    assert cs.n == expected_n
    if spec["A2"] is not None:
        np.testing.assert_allclose(spec["A2"] @ cs.y0, spec["b2"], atol=1e-9)
    # strictly inside every inequality-type constraint
    assert np.all(cs.b_p - cs.A_p @ cs.z0 > 1e-8)
    for c in list(cs.qcs) + list(cs.socs):
        assert c.residual(cs.y0)[0] < -1e-8
    if cs.lmic is not None:
        assert cs.lmic.residual(cs.y0)[0] < -1e-8
    assert cs.getViolation(cs.y0) <= 1e-12      # only the equality residual's round-off


def test_redundant_rows_and_implicit_equalities():
    # x <= 1, x <= 2 (redundant), y <= 0, -y <= 0 (implicit equality y = 0), -x <= 1
    A1 = np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.0, -1.0], [-1.0, 0.0]])
    b1 = np.array([[1.0], [2.0], [0.0], [0.0], [1.0]])
    lc = constraints.LinearConstraint(A1, b1, None, None)
    cs = constraints.ConvexConstraints(lc=lc)
    assert cs.n == 1 and cs.k == 2
    assert cs.A_E.shape[0] == 2                       # the pair y <= 0, -y <= 0
    assert cs.A_I.shape[0] == 2                       # x <= 1 and -x <= 1 (x <= 2 was dropped)
    assert abs(cs.y0[1, 0]) < 1e-9 and -1 < cs.y0[0, 0] < 1


def test_empty_set_is_reported():
    A1 = np.array([[1.0], [-1.0]])
    b1 = np.array([[-1.0], [-1.0]])    # x <= -1 and x >= 1
    with pytest.raises(Exception):
        constraints.ConvexConstraints(lc=constraints.LinearConstraint(A1, b1, None, None))


def test_data_dict_and_violation():
    cs = synthetic.build_constraints(synthetic.example_spec("readme"))
    d = cs.getDataAsDict()
    assert set(d) == {"A1", "b1", "A2", "b2", "all_P", "all_q", "all_r", "all_M", "all_s", "all_c", "all_d", "all_F"}
    assert cs.getViolation(cs.y0) <= 1e-12
    assert cs.getViolation(np.array([[2.0], [2.0], [2.0]])) > 0.5
