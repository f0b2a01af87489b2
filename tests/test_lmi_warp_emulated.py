"""The definiteness filter and the one-warp-per-matrix eigen-solver of rayen_b200/csrc/lmi_warp.cuh, compiled for the
HOST under the SIMT emulator (tests/emu) and checked against numpy / the float64 oracle.  Test infrastructure only:
the arithmetic, the shuffle / barrier structure and the scratch indexing are exercised on the CPU, where there is no
GPU; the product never loads this library (the GPU parity tests are in test_gpu_parity.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.rayen_oracle import OracleSet, closed_form_numpy
from rayen_b200 import _cabi, plan, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int32)
_LL = ctypes.c_longlong


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("lmi_warp_emu")
    lib_path = str(out / "liblmi_warp_emu.so")
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-w", f"-I{os.path.join(HERE, 'emu')}", "-o", lib_path,
           os.path.join(HERE, "emu", "lmi_warp_emu.cpp")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    lib = ctypes.CDLL(lib_path)
    lib.emu_lmi_warp.restype = ctypes.c_int
    lib.emu_lmi_warp.argtypes = [_F, ctypes.c_int, ctypes.c_int, _F, _F, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F, _LL, _F,
                                 _F, _I, _F, _LL, _I, _LL, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _I, _I, ctypes.c_int]
    return lib


def _ptr(a, t=_F):
    return a.ctypes.data_as(t) if a is not None else None


def run_warp_path(lib, p, v, kprior, tprior, mode=0, use_filter=True, with_grad=True, work_list=None, warps=3,
                  solves_per_warp=1 << 20, return_fails=False, mt=4):
    """kappa_io / active_io start as the prior of the other families; returns y, kappa, active, dkappa (and the list of
    samples handed over to the other kernel when a warp's solve budget is limited)."""
    f = p.fields
    n, k = f["n"], f["k"]
    blob = p.blob
    FW = np.ascontiguousarray(blob[f["off_lmiw"]:f["off_lmiw"] + n * 32 * 36])
    y0 = np.ascontiguousarray(blob[f["off_y0"]:f["off_y0"] + k])
    nmat = np.ascontiguousarray(blob[f["off_nmat"]:f["off_nmat"] + k * (f["np"] + 4)]) if not f["n_is_identity"] else None
    v = np.ascontiguousarray(v, dtype=np.float32)
    B, cols = v.shape
    y = np.full((B, k), np.nan, dtype=np.float32)
    kap = np.ascontiguousarray(kprior, dtype=np.float32).copy()
    act = np.ascontiguousarray(tprior, dtype=np.int32).copy()
    dk = np.full((B, n), np.nan, dtype=np.float32)
    wl = np.ascontiguousarray(work_list, dtype=np.int32) if work_list is not None else None
    fails = np.full((B,), -1, dtype=np.int32)
    nfail = np.zeros((1,), dtype=np.int32)
    rc = lib.emu_lmi_warp(_ptr(FW), n, k, _ptr(y0), _ptr(nmat), f["np"] + 4, f["n_is_identity"], mode, _ptr(v), cols, _ptr(y),
                          _ptr(kap), _ptr(act, _I), _ptr(dk), B, _ptr(wl, _I) if wl is not None else None,
                          0 if wl is None else len(wl), int(use_filter), int(with_grad), warps, solves_per_warp,
                          _ptr(fails, _I), _ptr(nfail, _I), mt)
    assert rc == 0
    out = (y.astype(np.float64), kap.astype(np.float64), act, dk.astype(np.float64))
    if return_fails:
        assert (fails[nfail[0]:] == -1).all()
        return out + (np.sort(fails[:nfail[0]]),)
    assert nfail[0] == 0
    return out


def _lmi_truth(p, v):
    """float64 lambda_max, top eigenvector and q'F_a q of S~(u) from the plan's float64 matrices."""
    Fz = p.f64["Fz"]
    n = Fz.shape[0]
    v = np.asarray(v, dtype=np.float64)[:, :n]
    s = np.linalg.norm(v, axis=1)
    u = v / np.maximum(s, 1e-12)[:, None]
    S = np.einsum("ba,aij->bij", u, Fz)
    lam, Q = np.linalg.eigh(S)
    q = Q[:, :, -1]
    grad = np.einsum("bi,aij,bj->ba", q, Fz, q)
    gap = (lam[:, -1] - lam[:, -2]) / np.maximum(np.abs(lam).max(axis=1), 1e-300) if S.shape[1] > 1 else np.ones(len(s))
    return lam[:, -1], grad, gap, s, u


@pytest.mark.parametrize("k,r,eq,use_filter", [(8, 32, 0, False), (8, 32, 0, True), (32, 32, 0, True), (5, 12, 0, True),
                                                (3, 2, 0, True), (12, 27, 2, True), (6, 32, 1, False)])
def test_warp_solver_and_filter_match_numpy(emu, k, r, eq, use_filter):
    """Random LMIs (sizes that need zero padding, subspaces with equalities): with prior kappa values spread around
    lambda_max the filter must pass exactly the samples it can prove, and every other sample must get lambda_max, the
    merged (kappa, tag), y and d kappa/du of the float64 computation."""
    spec = synthetic.random_spec(k=k, r=r, seed=3 * k + r)
    if eq:
        rng = np.random.default_rng(k)
        spec["A2"], spec["b2"] = rng.uniform(-1, 1, size=(eq, k)), np.zeros((eq, 1))
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    n = cs.n
    B = 41
    v, _ = synthetic.sample_inputs(B, n, cs.k, seed_v=r, scale=6.0)
    v = v.numpy()
    lam, grad, gap, s, u = _lmi_truth(p, v)
    rng = np.random.default_rng(r)
    # priors: far above, just above, just below, far below lambda_max, and zero
    factor = rng.choice([3.0, 1.0 + 1e-3, 1.0 + 2e-5, 1.0 - 2e-5, 1.0 - 1e-3, 0.3, 0.0], size=B)
    kprior = (np.maximum(lam, 0.0) * factor).astype(np.float32)
    tprior = np.where(kprior > 0, (1 << 24) | 5, 0).astype(np.int32)
    y, kap, act, dk = run_warp_path(emu, p, v, kprior, tprior, use_filter=use_filter)
    want_kap = np.maximum(np.maximum(lam, 0.0), kprior.astype(np.float64))
    clear = np.abs(np.maximum(lam, 0) - kprior) > 1e-4 * np.maximum(lam, 1e-30)
    lmi_binds = np.maximum(lam, 0.0) > kprior
    assert np.abs(kap - want_kap).max() <= 3e-6 * want_kap.max()
    assert np.array_equal((act >> 24)[clear], np.where(lmi_binds, 4, np.where(kprior > 0, 1, 0))[clear])
    with np.errstate(divide="ignore"):
        alpha = np.minimum(np.where(want_kap > 0, 1.0 / np.where(want_kap > 0, want_kap, 1.0), np.inf), s)
    N, y0 = p.f64["N"], p.f64["y0"][:, 0]
    y_want = y0[None, :] + alpha[:, None] * (u @ N.T)
    assert np.isfinite(y).all()
    assert np.abs(y - y_want).max() <= 5e-6 * np.abs(y_want).max()
    # gradient rows: written for LMI-bound boundary samples only
    need = lmi_binds & (act >> 24 == 4) & (1.0 / np.where(kap > 0, kap, 1.0) < s)
    well = need & (gap > 1e-3)
    assert need.sum() >= 5
    assert np.isfinite(dk[need]).all()
    assert np.abs(dk[well] - grad[well]).max() <= 2e-4 * np.abs(grad[well]).max() / np.minimum(gap[well].min() * 1e2, 1.0)
    assert np.isnan(dk[~need]).all()           # nobody else's row is touched


def test_filter_pass_is_a_proof_and_fail_is_rare(emu):
    """The filter in isolation (no solver fallback visible in kappa): on an epigraph LMI -- S~ close to a multiple of I,
    lambda_max close to the prior -- a sample may pass only if lambda_max < prior in float64, and samples whose prior
    is more than 1e-4 above lambda_max (relative to the matrix scale) must pass."""
    spec = synthetic.epigraph_lmi_spec(8, 32, 1e-3, seed=4)
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    B = 64
    v, _ = synthetic.sample_inputs(B, cs.n, cs.k, seed_v=9, scale=3.0)
    v = v.numpy()
    v[:, 0] = -np.abs(v[:, 0]) * 6.0
    lam, grad, gap, s, u = _lmi_truth(p, v)
    rng = np.random.default_rng(0)
    rel = rng.choice([-3e-4, -3e-5, -3e-6, 0.0, 3e-6, 3e-5, 3e-4, 1e-2], size=B)
    kprior = (lam * (1.0 + rel)).astype(np.float32)
    tprior = np.full(B, (1 << 24) | 1, dtype=np.int32)
    y, kap, act, dk = run_warp_path(emu, p, v, kprior, tprior, use_filter=True, with_grad=False)
    # whatever the filter decided, the outcome is the exact merge (a wrong pass would leave kappa = prior < lambda_max)
    want = np.maximum(lam, kprior.astype(np.float64))
    assert np.abs(kap - want).max() <= 2e-6 * want.max()
    fam = act >> 24
    assert (fam[rel <= -3e-5] == 4).all() and (fam[rel >= 3e-5] == 1).all()


def test_dense_mode_and_work_list_agree(emu):
    spec = synthetic.config_spec("cfg5")
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    B = 23
    v, gy = synthetic.sample_inputs(B, cs.n, cs.k, seed_v=5, scale=5.0)
    cf = closed_form_numpy(OracleSet.from_constraints(cs), v.numpy(), gy.numpy())
    lam, grad, gap, s, u = _lmi_truth(p, v.numpy())
    kprior = (0.5 * np.maximum(lam, 0.0) + 0.5 * np.maximum(lam, 0.0) * (np.arange(B) % 3)).astype(np.float32)
    tprior = np.full(B, (3 << 24) | 2, dtype=np.int32)
    dense = run_warp_path(emu, p, v.numpy(), kprior, tprior)
    wl = np.array([20, 3, 4, 11, 0, 7, 22], dtype=np.int32)
    y, kap, act, dk = run_warp_path(emu, p, v.numpy(), kprior, tprior, work_list=wl)
    for a, b_ in zip((y, kap, act), dense[:3]):
        assert np.array_equal(a[wl], b_[wl])
    rest = np.setdiff1d(np.arange(B), wl)
    assert np.isnan(y[rest]).all() and np.array_equal(kap[rest], kprior[rest].astype(np.float64))


def test_solve_budget_hands_the_rest_to_the_fail_list(emu):
    """A warp solves at most `solves_per_warp` failing samples itself; the others must appear in the fail list exactly
    once, untouched (prior kappa / tag still in place, no y), and everything else must be final."""
    spec = synthetic.random_spec(k=8, r=32, seed=5)
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    B = 37
    v, _ = synthetic.sample_inputs(B, cs.n, cs.k, seed_v=2, scale=6.0)
    v = v.numpy()
    lam, grad, gap, s, u = _lmi_truth(p, v)
    kprior = (np.maximum(lam, 0.0) * np.where(np.arange(B) % 3 == 0, 2.0, 0.5)).astype(np.float32)   # 2/3 of them fail
    tprior = np.full(B, (1 << 24) | 7, dtype=np.int32)
    full = run_warp_path(emu, p, v, kprior, tprior, warps=2)
    y, kap, act, dk, fails = run_warp_path(emu, p, v, kprior, tprior, warps=2, solves_per_warp=1, return_fails=True)
    should_fail = np.flatnonzero(np.maximum(lam, 0.0) > kprior)
    assert len(fails) == len(should_fail) - 2 and set(fails) <= set(should_fail)      # two warps, one solve each
    done = np.setdiff1d(np.arange(B), fails)
    for a, b_ in zip((y, kap, act), full[:3]):
        assert np.array_equal(a[done], b_[done])
    assert np.isnan(y[fails]).all() and np.array_equal(kap[fails], kprior[fails].astype(np.float64))
    assert (act[fails] == tprior[fails]).all()


def test_two_sample_chunks_give_the_same_bits_as_four(emu):
    """Short work lists go through the filter two samples per warp instead of four (lmi_forward_warp_kernel: more warps
    in flight on half the work each).  A sample's arithmetic does not depend on the chunk it falls into: kappa, tag, y and
    d kappa/du are bit-identical."""
    spec = synthetic.random_spec(k=9, r=24, seed=12)
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    B = 23
    v, _ = synthetic.sample_inputs(B, cs.n, cs.k, seed_v=3, scale=6.0)
    v = v.numpy()
    lam, grad, gap, s, u = _lmi_truth(p, v)
    rng = np.random.default_rng(1)
    kprior = (np.maximum(lam, 0.0) * rng.choice([3.0, 1.0 + 1e-3, 1.0 - 1e-3, 0.3, 0.0], size=B)).astype(np.float32)
    tprior = np.where(kprior > 0, (1 << 24) | 2, 0).astype(np.int32)
    a = run_warp_path(emu, p, v, kprior, tprior, mt=4)
    b = run_warp_path(emu, p, v, kprior, tprior, mt=2, warps=5)
    for x, y_ in zip(a, b):
        np.testing.assert_array_equal(x, y_)
