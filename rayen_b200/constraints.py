"""Constraint value types and the one-time HOST preprocessing of the feasible set.

Drop-in for the reference's ``rayen/constraints.py`` (same class names, constructor
signatures, attribute names and error behaviour):

* ``LinearConstraint(A1, b1, A2, b2)``            reference constraints.py:17-61
* ``ConvexQuadraticConstraint(P, q, r, do_checks_P)``   reference constraints.py:63-106
* ``SOCConstraint(M, s, c, d)``                   reference constraints.py:108-130
* ``LMIConstraint(all_F)``                        reference constraints.py:132-155
* ``ConvexConstraints(lc, qcs, socs, lmic, y0, do_preprocessing_linear, print_debug_info)``
                                                  reference constraints.py:159-447

Everything here is numpy/scipy in float64 and runs once, on the host (BASELINE.json north_star:
"preprocessing (interior point z0, null-space N) stays one-time on the host").  The reference
drives cvxpy (ECOS/SCS/GUROBI) for its LPs and for the interior point; this build is solver-free:
the LPs go through ``scipy.optimize.linprog`` (HiGHS) and the strictly interior point of a
non-linear set through SLSQP on the same max-margin program (see ``_interior_point``).
``asCvxpy``/``project``/``getViolation`` of the reference need a conic solver and are replaced by
direct residual checks (``getViolation`` here returns the max constraint residual).
"""
import math

import numpy as np
import scipy.linalg
import scipy.optimize

from . import utils


# --------------------------------------------------------------------------- value types
class LinearConstraint:
    """A1 y <= b1 and A2 y = b2.  Either pair may be (None, None), not both."""

    def __init__(self, A1, b1, A2, b2):
        self.A1, self.b1, self.A2, self.b2 = A1, b1, A2, b2
        utils.verify(self.hasEqConstraints() or self.hasIneqConstraints())
        for A, b in ((A1, b1), (A2, b2)):
            if A is not None and b is not None:
                utils.verify(A.ndim == 2)
                utils.verify(b.ndim == 2)
                utils.verify(b.shape[1] == 1)
                utils.verify(A.shape[0] == b.shape[0])
        if self.hasIneqConstraints() and self.hasEqConstraints():
            utils.verify(A1.shape[1] == A2.shape[1])

    def hasEqConstraints(self):
        return self.A2 is not None and self.b2 is not None

    def hasIneqConstraints(self):
        return self.A1 is not None and self.b1 is not None

    def dim(self):
        return self.A1.shape[1] if self.hasIneqConstraints() else self.A2.shape[1]

    def residual(self, y):
        """Max violation of the rows at the column vector(s) y:[k,B] (<= 0 means satisfied)."""
        res = np.full(y.shape[1], -np.inf)
        if self.hasIneqConstraints():
            res = np.maximum(res, np.max(self.A1 @ y - self.b1, axis=0))
        if self.hasEqConstraints():
            res = np.maximum(res, np.max(np.abs(self.A2 @ y - self.b2), axis=0))
        return res


class ConvexQuadraticConstraint:
    """(1/2) y'P y + q'y + r <= 0 with P symmetric PSD."""

    def __init__(self, P, q, r, do_checks_P=True):
        self.P, self.q, self.r = P, q, r
        if do_checks_P:
            utils.checkMatrixisNotZero(self.P)
            utils.checkMatrixisSymmetric(self.P)
            smallest = float(np.amin(np.linalg.eigvalsh(self.P)))
            tol = 1e-7
            utils.verify(smallest > -tol, f"Matrix P is not PSD, smallest eigenvalue is {smallest}")
            if -tol <= smallest < 0:
                # repair round-off: shift the spectrum so that P is PSD (reference constraints.py:88-92)
                self.P = self.P + abs(smallest) * np.eye(self.P.shape[0])

    def dim(self):
        return self.P.shape[1]

    def residual(self, y):
        return 0.5 * np.einsum("ib,ij,jb->b", y, self.P, y) + (self.q.T @ y)[0] + float(np.asarray(self.r).reshape(-1)[0])


class SOCConstraint:
    """||M y + s|| <= c'y + d."""

    def __init__(self, M, s, c, d):
        utils.checkMatrixisNotZero(M)
        utils.checkMatrixisNotZero(c)
        utils.verify(M.shape[1] == c.shape[0])
        utils.verify(M.shape[0] == s.shape[0])
        utils.verify(s.shape[1] == 1)
        utils.verify(c.shape[1] == 1)
        utils.verify(d.shape[0] == 1)
        utils.verify(d.shape[1] == 1)
        self.M, self.s, self.c, self.d = M, s, c, d

    def dim(self):
        return self.M.shape[1]

    def residual(self, y):
        return np.linalg.norm(self.M @ y + self.s, axis=0) - (self.c.T @ y)[0] - float(self.d[0, 0])


class LMIConstraint:
    """y_0 F_0 + ... + y_{k-1} F_{k-1} + F_k >= 0 (positive semidefinite)."""

    def __init__(self, all_F):
        for F in all_F:
            utils.checkMatrixisSymmetric(F)
        for F in all_F:
            utils.verify(F.shape == all_F[0].shape)
        self.all_F = all_F

    def dim(self):
        return len(self.all_F) - 1

    def evaluate(self, y):
        """F(y) for y:[k,B] -> [B,r,r]."""
        F = np.asarray(self.all_F)
        return np.einsum("ab,aij->bij", y, F[:-1]) + F[-1]

    def residual(self, y):
        return -np.linalg.eigvalsh(self.evaluate(y))[:, 0]


# --------------------------------------------------------------------------- the feasible set
class ConvexConstraints:
    """Intersection of the given constraints plus everything the layer needs about it.

    ``y0`` (a point in the relative interior) may be supplied; it is then trusted, exactly as in
    the reference (constraints.py:160-162).  ``do_preprocessing_linear=False`` may only be used
    when the caller knows that aff{y: A1 y <= b1} = R^k (constraints.py:164).

    Public fields (same names as the reference): ``k, n, A_p, b_p, NA_E, yp, z0, y0, A_E, b_E,
    A_I, b_I, lc, qcs, socs, lmic, has_*``.
    """

    def __init__(self, lc=None, qcs=[], socs=[], lmic=None, y0=None,
                 do_preprocessing_linear=True, print_debug_info=False):
        self.lc, self.qcs, self.socs, self.lmic = lc, qcs, socs, lmic
        self.has_linear_eq_constraints = lc is not None and lc.hasEqConstraints()
        self.has_linear_ineq_constraints = lc is not None and lc.hasIneqConstraints()
        self.has_linear_constraints = self.has_linear_eq_constraints or self.has_linear_ineq_constraints
        self.has_quadratic_constraints = len(qcs) > 0
        self.has_soc_constraints = len(socs) > 0
        self.has_lmi_constraints = lmic is not None
        utils.verify(self.has_linear_constraints or self.has_quadratic_constraints
                     or self.has_soc_constraints or self.has_lmi_constraints, "There are no constraints!")

        dims = ([lc.dim()] if self.has_linear_constraints else []) + [c.dim() for c in qcs] \
            + [c.dim() for c in socs] + ([lmic.dim()] if self.has_lmi_constraints else [])
        utils.verify(utils.all_equal(dims))
        self.k = dims[0]
        self.solver = "scipy-highs+newton"  # the reference stores the cvxpy solver name here
        self._debug = print_debug_info

        if self.has_linear_constraints:
            A, b = self._stacked_inequalities()
            if do_preprocessing_linear:
                A, b = self._drop_redundant_rows(A, b)
                E = self._equality_set(A, b)
            else:
                # E := the rows that came from (A2, -A2)   (reference constraints.py:331-339)
                first = lc.A1.shape[0] if self.has_linear_ineq_constraints else 0
                E = list(range(first, A.shape[0]))
            I = [i for i in range(A.shape[0]) if i not in E]
            A_E, b_E = (A[E, :], b[E, :]) if E else (np.zeros((1, self.k)), np.zeros((1, 1)))
            A_I, b_I = (A[I, :], b[I, :]) if I else (np.zeros((1, self.k)), np.ones((1, 1)))
            NA_E = scipy.linalg.null_space(A_E)
            yp = np.linalg.pinv(A_E) @ b_E
            A_p = A_I @ NA_E
            b_p = b_I - A_I @ yp
            self.n = A_p.shape[1]
        else:
            self.n = self.k
            NA_E, yp = np.eye(self.k), np.zeros((self.k, 1))
            A_p, b_p = np.zeros((1, self.k)), np.ones((1, 1))
            A_E, b_E = np.zeros((1, self.k)), np.zeros((1, 1))
            A_I, b_I = np.zeros((1, self.k)), np.ones((1, 1))

        self.A_E, self.b_E, self.A_I, self.b_I = A_E, b_E, A_I, b_I
        self.A_p, self.b_p, self.yp, self.NA_E = A_p, b_p, yp, NA_E
        utils.verify(self.n == self.k - np.linalg.matrix_rank(self.A_E))
        utils.verify(np.allclose(NA_E.T @ NA_E, np.eye(NA_E.shape[1])))

        if y0 is None:
            self.z0 = self._interior_point()
            self.y0 = self.NA_E @ self.z0 + self.yp
        else:
            self.y0 = y0
            self.z0 = self.NA_E.T @ (self.y0 - self.yp)

    # ----------------------------------------------------------------- linear preprocessing
    def _stacked_inequalities(self):
        """[A1; A2; -A2] y <= [b1; b2; -b2]   (reference constraints.py:240-250)."""
        rows_A, rows_b = [], []
        if self.has_linear_ineq_constraints:
            rows_A.append(self.lc.A1)
            rows_b.append(self.lc.b1)
        if self.has_linear_eq_constraints:
            rows_A += [self.lc.A2, -self.lc.A2]
            rows_b += [self.lc.b2, -self.lc.b2]
        return np.concatenate(rows_A, axis=0).astype(float), np.concatenate(rows_b, axis=0).astype(float)

    @staticmethod
    def _lp(c, A_ub, b_ub):
        """min c'z s.t. A_ub z <= b_ub, z free.  Returns (status, value): status in {'optimal','unbounded'}."""
        res = scipy.optimize.linprog(c, A_ub=A_ub, b_ub=b_ub.reshape(-1), bounds=(None, None), method="highs")
        if res.status == 0:
            return "optimal", float(res.fun)
        if res.status == 3:
            return "unbounded", -math.inf
        if res.status == 2:
            raise Exception("The feasible set is empty")
        raise Exception(f"LP failed: {res.message}")

    def _drop_redundant_rows(self, A, b, tol=1e-7):
        """Row i is redundant iff max A_i z s.t. the other rows and A_i z <= b_i + 1 stays <= b_i.

        Same test and the same back-to-front sweep as the reference (constraints.py:260-285).
        """
        if A.shape[0] <= 1:
            return A, b
        for i in reversed(range(A.shape[0])):
            others = [j for j in range(A.shape[0]) if j != i]
            A_ub = np.concatenate((A[others, :], A[i:i + 1, :]), axis=0)
            b_ub = np.concatenate((b[others, :], b[i:i + 1, :] + 1.0), axis=0)
            status, val = self._lp(-A[i, :], A_ub, b_ub)
            if status != "optimal":
                raise Exception("Value is not optimal")
            if (-val) - b[i, 0] <= tol:
                A = np.delete(A, i, axis=0)
                b = np.delete(b, i, axis=0)
        return A, b

    def _equality_set(self, A, b, tol=1e-5):
        """Indices of rows that hold with equality on the whole set (reference constraints.py:295-329)."""
        E = []
        for i in range(A.shape[0]):
            status, val = self._lp(A[i, :], A, b)
            val = val - b[i, 0] if status == "optimal" else -math.inf
            utils.verify(val < tol, f"The objective should be negative. It's {val} right now")
            if val > -tol:
                E.append(i)
        return E

    # ----------------------------------------------------------------- strictly interior point
    def _slacks(self, z):
        """All constraint slacks at z (positive inside): returns (linear[m], quad[eta], soc[mu], lmi_min_eig or None)."""
        y = self.NA_E @ z + self.yp
        lin = (self.b_p - self.A_p @ z)[:, 0]
        quad = np.array([-float(qc.residual(y)[0]) for qc in self.qcs])
        soc = np.array([-float(sc.residual(y)[0]) for sc in self.socs])
        lmi = -float(self.lmic.residual(y)[0]) if self.has_lmi_constraints else None
        return lin, quad, soc, lmi

    def _min_slack(self, z):
        lin, quad, soc, lmi = self._slacks(z)
        vals = [np.min(lin)] + ([np.min(quad)] if quad.size else []) + ([np.min(soc)] if soc.size else []) \
            + ([lmi] if lmi is not None else [])
        return float(min(vals))

    def _interior_point(self):
        """z0 = argmax eps s.t. every constraint holds with margin eps, 0 <= eps <= 0.5.

        Restates the conic program of the reference (constraints.py:412-432).  A linear-only set is
        one LP.  Otherwise the same program is solved with SLSQP on smooth forms of the constraints
        (the SOC as (c'y+d)^2-||My+s||^2 plus c'y+d >= eps; the LMI through its smallest eigenvalue,
        whose gradient is q'F_i q), started from the LP / least-squares point and verified afterwards.
        """
        n = self.n
        nonlinear = self.has_quadratic_constraints or self.has_soc_constraints or self.has_lmi_constraints
        # LP on (z, eps): max eps s.t. A_p z + eps <= b_p, 0 <= eps <= 0.5
        c = np.zeros(n + 1)
        c[-1] = -1.0
        A_ub = np.concatenate((self.A_p, np.ones((self.A_p.shape[0], 1))), axis=1)
        bounds = [(None, None)] * n + [(0.0, 0.5)]
        if not nonlinear:
            res = scipy.optimize.linprog(c, A_ub=A_ub, b_ub=self.b_p[:, 0], bounds=bounds, method="highs")
            if res.status == 2:
                raise Exception("The feasible set is empty")
            if res.status == 3:  # unbounded directions: bound z and retry
                res = scipy.optimize.linprog(c, A_ub=A_ub, b_ub=self.b_p[:, 0],
                                             bounds=[(-1e6, 1e6)] * n + [(0.0, 0.5)], method="highs")
            if res.status != 0:
                raise Exception(f"Value is not optimal, prob_status={res.message}")
            utils.verify(res.x[-1] > 1e-8, "There are no strictly feasible points in the subspace")
            return res.x[:n].reshape(n, 1)

        N, yp = self.NA_E, self.yp
        F_all = np.asarray(self.lmic.all_F) if self.has_lmi_constraints else None

        def cons(x):
            z, eps = x[:n].reshape(n, 1), x[n]
            lin, quad, soc, lmi = self._slacks(z)
            parts = [lin - eps, quad - eps, soc - eps]
            if lmi is not None:
                parts.append(np.array([lmi - eps]))
            return np.concatenate(parts)

        x0 = np.zeros(n + 1)
        best = None
        rng = np.random.default_rng(0)
        for attempt in range(8):
            res = scipy.optimize.minimize(lambda x: -x[n], x0, method="SLSQP",
                                          constraints=[{"type": "ineq", "fun": cons}],
                                          bounds=[(None, None)] * n + [(0.0, 0.5)],
                                          options={"maxiter": 500, "ftol": 1e-10})
            z = res.x[:n].reshape(n, 1)
            margin = self._min_slack(z)
            if best is None or margin > best[0]:
                best = (margin, z)
            if margin > 1e-6:
                break
            x0 = np.concatenate((rng.normal(size=n), [0.0]))
        margin, z = best
        if margin <= 0:
            raise Exception("The feasible set is empty")
        utils.verify(margin > 1e-8, "There are no strictly feasible points in the subspace")
        return z

    # ----------------------------------------------------------------- data export / checks
    def getDataAsDict(self):
        """Same keys and 'no constraint' placeholders as the reference (constraints.py:450-497)."""
        k = self.k
        A2, b2 = (self.lc.A2, self.lc.b2) if self.has_linear_eq_constraints else (np.zeros((1, k)), np.array([[0]]))
        A1, b1 = (self.lc.A1, self.lc.b1) if self.has_linear_ineq_constraints else (np.zeros((1, k)), np.array([[1]]))
        if self.has_quadratic_constraints:
            all_P, all_q, all_r = utils.getAllPqrFromQcs(self.qcs)
        else:
            all_P, all_q, all_r = [np.zeros((k, k))], [np.zeros((k, 1))], [-np.ones((1, 1))]
        if self.has_soc_constraints:
            all_M, all_s, all_c, all_d = utils.getAllMscdFromSocs(self.socs)
        else:
            all_M, all_s, all_c, all_d = [np.zeros((k, k))], [np.zeros((k, 1))], [np.zeros((k, 1))], [np.ones((1, 1))]
        if self.has_lmi_constraints:
            all_F = self.lmic.all_F
        else:
            all_F = [np.zeros((k, k)) for _ in range(k)] + [np.eye(k)]
        return dict(A2=A2, b2=b2, A1=A1, b1=b1, all_P=all_P, all_q=all_q, all_r=all_r,
                    all_M=all_M, all_s=all_s, all_c=all_c, all_d=all_d, all_F=all_F)

    def residuals(self, y):
        """Max residual over every constraint for each column of y:[k,B] (float64; <= 0 is feasible).

        Replaces the reference's per-sample cvxpy projection distance (constraints.py:539-559), which
        needs a conic solver; this is the direct check SURVEY §8f-2 asks for.
        """
        y = np.asarray(y, dtype=np.float64)
        worst = np.full(y.shape[1], -np.inf)
        if self.has_linear_constraints:
            worst = np.maximum(worst, self.lc.residual(y))
        for c in list(self.qcs) + list(self.socs):
            worst = np.maximum(worst, c.residual(y))
        if self.has_lmi_constraints:
            worst = np.maximum(worst, self.lmic.residual(y))
        return worst

    def getViolation(self, y_to_be_projected):
        """Max constraint residual clipped at 0 for one point (0 means feasible)."""
        y = np.asarray(y_to_be_projected, dtype=np.float64)
        if y.ndim == 1:
            y = y[:, None]
        return float(max(0.0, self.residuals(y)[0]))
