"""Small host-side helpers shared by the constraint classes and the layer.

Mirrors the part of the reference's ``rayen/utils.py`` that the RAYEN path uses
(reference: rayen/utils.py:21-46 verify / getAllPqrFromQcs / getAllMscdFromSocs,
:113-131 matrix checks, :245-251 all_equal, :49-61 CudaTimer).  The baselines-only helpers
(rref, H_to_V, power iteration; utils.py:74-106, :138-207, :272-337) are out of scope.
"""
import numpy as np
import torch


def verify(condition, message="Condition not satisfied"):
    """Raise ``RuntimeError(message)`` unless ``condition`` holds (reference utils.py:21-23)."""
    if not bool(condition):
        raise RuntimeError(message)


def getAllPqrFromQcs(qcs):
    """Split a list of quadratic constraints into three parallel lists (reference utils.py:25-33)."""
    return [qc.P for qc in qcs], [qc.q for qc in qcs], [qc.r for qc in qcs]


def getAllMscdFromSocs(socs):
    """Split a list of SOC constraints into four parallel lists (reference utils.py:35-46)."""
    return ([soc.M for soc in socs], [soc.s for soc in socs],
            [soc.c for soc in socs], [soc.d for soc in socs])


def isZero(A):
    return not np.any(A)


def checkMatrixisNotZero(A):
    verify(not isZero(A), "Matrix is identically zero")


def checkMatrixisSymmetric(A):
    verify(A.ndim == 2 and A.shape[0] == A.shape[1], "Matrix is not square")
    verify(np.allclose(A, A.T), "Matrix is not symmetric")


def checkMatrixisPsd(A, tol=0.0):
    checkMatrixisSymmetric(A)
    lam = np.linalg.eigvalsh(A)
    verify(np.all(lam >= -tol), f"Matrix is not PSD, min eigenvalue is {np.amin(lam)}")


def checkMatrixisPd(A):
    checkMatrixisSymmetric(A)
    lam = np.linalg.eigvalsh(A)
    verify(np.all(lam > 0.0), f"Matrix is not PD, min eigenvalue is {np.amin(lam)}")


def all_equal(iterable):
    items = list(iterable)
    return all(item == items[0] for item in items[1:])


def quadExpression(y, P, q, r):
    """Batched (1/2) y'Py + q'y + r for y:[B,k,1] (reference utils.py:229-243)."""
    P, q, r = P.to(y.device), q.to(y.device), r.to(y.device)
    qT = q.T if q.ndim == 2 else torch.transpose(q, 1, 2)
    return 0.5 * torch.transpose(y, 1, 2) @ P @ y + qT @ y + r


class CudaTimer:
    """CUDA-event stopwatch on the current stream (reference utils.py:49-61)."""

    def start(self):
        self._t0 = torch.cuda.Event(enable_timing=True)
        self._t1 = torch.cuda.Event(enable_timing=True)
        self._t0.record()

    def endAndGetTimeSeconds(self):
        self._t1.record()
        torch.cuda.synchronize()
        return 1e-3 * self._t0.elapsed_time(self._t1)
