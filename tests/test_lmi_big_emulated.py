"""The big-LMI kernels (rayen_b200/csrc/lmi_big.cuh: contraction GEMM + one-CTA-per-sample eigen-solver, LMI sizes 33..320
and LMIs together with n > 32) compiled for the HOST under the SIMT emulator (tests/emu) and checked against the oracle.
The prior (kappa, tag, y of the other families) that the kernels expect is produced here by the float64 evaluation of
the same packed plan with the LMI left out.  Test infrastructure only: the product never loads this library."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.rayen_oracle import OracleSet, closed_form_numpy
from rayen_b200 import plan, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int32)
_LL = ctypes.c_longlong


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("lmib_emu")
    lib_path = str(out / "liblmib_emu.so")
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-w", f"-I{os.path.join(HERE, 'emu')}",
           "-o", lib_path, os.path.join(HERE, "emu", "lmi_big_emu.cpp")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    lib = ctypes.CDLL(lib_path)
    lib.emu_lmib_contract.restype = ctypes.c_int
    lib.emu_lmib_contract.argtypes = [_F, _LL, _F, ctypes.c_int, ctypes.c_int, _F, _LL, _F]
    lib.emu_lmib_solve.restype = ctypes.c_int
    lib.emu_lmib_solve.argtypes = [_F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F,
                                   _F, _LL, _F, _F, _I, _F, _LL, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F, _I, _I]
    lib.emu_lmib_grad_gemm.restype = ctypes.c_int
    lib.emu_lmib_grad_gemm.argtypes = [_F, _I, _I, _F, ctypes.c_int, ctypes.c_int, _F, _LL]
    return lib


def _ptr(a, t=_F):
    return a.ctypes.data_as(t)


def _prior(cs, p, v, mode):
    """(kappa, tag, y) of the linear / quadratic / SOC constraints alone, as the kernel in front leaves them."""
    f = dict(p.fields)
    no_lmi = plan.PackedPlan()
    no_lmi.blob, no_lmi.fields = p.blob, dict(f, lmi_r=0)
    ev = plan.evaluate_wide_numpy if f["wide"] else plan.evaluate_plan_numpy
    y, kap, act = ev(no_lmi, v[:, :cs.n])
    if mode == 1:   # RAYEN_old: alpha = 1 / (e^beta + kappa)
        s = np.linalg.norm(v[:, :cs.n], axis=1)
        u = v[:, :cs.n] / np.maximum(s, 1e-12)[:, None]
        alpha = 1.0 / (np.exp(v[:, cs.n]) + kap)
        y = cs.y0[:, 0][None] + alpha[:, None] * (u @ cs.NA_E.T)
    return y.astype(np.float32), kap.astype(np.float32), act.astype(np.int32)


def run_emulated(lib, cs, p, v, mode=0, threads=64, grid=3, flags=1, gemm_grad=False):
    f = p.fields
    n, k, r, p4 = f["n"], f["k"], f["lmi_r"], f["lmib_p4"]
    v = np.ascontiguousarray(v, dtype=np.float32)
    B, cols = v.shape
    y, kap, act = _prior(cs, p, v.astype(np.float64), mode)
    blob = p.blob
    S = np.full((B, p4), np.nan, dtype=np.float32)
    Fp = np.ascontiguousarray(blob[f["off_lmib"]:f["off_lmib"] + n * p4])
    assert lib.emu_lmib_contract(_ptr(v), cols, _ptr(Fp), n, p4, _ptr(S), B, None) == 0
    dk = np.full((B, n), np.nan, dtype=np.float32)
    glist = np.full((B,), -1, dtype=np.int32)
    gcount = np.zeros((1,), dtype=np.int32)
    rc = lib.emu_lmib_solve(_ptr(blob), n, k, r, p4, f["off_lmib"], f["off_y0"], _ptr(S), _ptr(v), cols, _ptr(y), _ptr(kap),
                            _ptr(act, _I), _ptr(dk), B, mode, flags, threads, grid, _ptr(S) if gemm_grad else None,
                            _ptr(glist, _I) if gemm_grad else None, _ptr(gcount, _I) if gemm_grad else None)
    assert rc == 0
    if gemm_grad:
        # the LMI-bound samples left their weight rows in S and their numbers on the list: one GEMM over the list
        assert lib.emu_lmib_grad_gemm(_ptr(S), _ptr(glist, _I), _ptr(gcount, _I), _ptr(Fp), n, p4, _ptr(dk), B) == 0
        assert sorted(glist[:gcount[0]].tolist()) == sorted(np.nonzero(~np.isnan(dk).any(axis=1))[0].tolist())
    return y.astype(np.float64), kap, act, dk, S


CASES = [
    # spec, batch, threads
    (lambda: synthetic.random_spec(k=5, m=8, eta=1, mu=1, r_M=4, r=33, seed=1), 20, 64),
    (lambda: synthetic.random_spec(k=3, r=40, seed=2), 12, 64),                 # LMI only
    (lambda: synthetic.wide_spec(36, 40, 1, 1, 6, 2, seed=3, r=9), 14, 64),      # wide subspace (n = 34), small LMI
    (lambda: synthetic.random_spec(k=4, m=4, r=70, seed=4), 6, 128),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_big_lmi_kernels_match_the_oracle(emu, case):
    make, B, threads = CASES[case]
    spec = make()
    if spec["b1"] is not None and case == 0:
        spec["b1"] = spec["b1"] * 3.0
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    assert p.fields["lmi_big"] == 1
    v, gy = synthetic.sample_inputs(B, cs.n, cs.k, seed_v=case + 1)
    v = v.numpy()
    v[1] *= 1e-3                                   # an interior sample
    y, kap, act, dk, S = run_emulated(emu, cs, p, v, threads=threads)
    # the gradient through the list GEMM (lmib_grad_gemm_kernel) instead of the per-sample walk: same rows, same values
    y2, kap2, act2, dk2, _ = run_emulated(emu, cs, p, v, threads=threads, gemm_grad=True)
    np.testing.assert_array_equal(kap2, kap)
    np.testing.assert_array_equal(y2, y)
    assert np.array_equal(np.isnan(dk2), np.isnan(dk))
    assert np.nanmax(np.abs(dk2 - dk)) <= 2e-6 * np.nanmax(np.abs(dk))
    oset = OracleSet.from_constraints(cs)
    cf = closed_form_numpy(oset, v, gy.numpy())
    assert np.abs(y - cf["y"]).max() <= 1e-5 * max(1.0, np.abs(cf["y"]).max())
    assert np.abs(kap - cf["kappa"]).max() <= 1e-5 * max(1.0, cf["kappa"].max())
    ok = cf["margin"] > 1e-4
    assert ((act >> 24)[ok] == cf["family"][ok]).all()
    lmi_bound = ((act >> 24) == 4)
    assert lmi_bound.any()
    # d kappa / du of the LMI-bound boundary samples against the float64 eigenvector
    Fz = p.f64["Fz"]
    s = np.linalg.norm(v, axis=1)
    u = v / s[:, None]
    need = lmi_bound & (1.0 / np.maximum(kap, 1e-30) < s)
    lam, Q = np.linalg.eigh(np.einsum("ba,aij->bij", u, Fz))
    g_ref = np.einsum("bi,aij,bj->ba", Q[:, :, -1], Fz, Q[:, :, -1])
    gap_ok = need & ((lam[:, -1] - lam[:, -2]) > 1e-3 * np.abs(lam[:, -1]))
    assert gap_ok.any()
    assert np.abs(dk[gap_ok] - g_ref[gap_ok]).max() <= 2e-5 * max(1.0, np.abs(g_ref).max())
    assert np.isnan(dk[~need]).all()               # nothing is written for the other samples


def test_big_lmi_rayen_old_and_gradient_only_mode(emu):
    spec = synthetic.random_spec(k=4, m=6, r=36, seed=7)
    spec["b1"] = spec["b1"] * 2.0
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    v, _ = synthetic.sample_inputs(10, cs.n + 1, cs.k, seed_v=9)
    v = v.numpy()
    y, kap, act, dk, S = run_emulated(emu, cs, p, v, mode=1)
    oset = OracleSet.from_constraints(cs)
    cf = closed_form_numpy(oset, v[:, :cs.n])
    s = np.linalg.norm(v[:, :cs.n], axis=1)
    u = v[:, :cs.n] / s[:, None]
    y_ref = cs.y0[:, 0][None] + (1.0 / (np.exp(v[:, cs.n]) + cf["kappa"]))[:, None] * (u @ cs.NA_E.T)
    assert np.abs(y - y_ref).max() <= 1e-5 * max(1.0, np.abs(y_ref).max())
    # gradient-only mode (backward without a forward-computed gradient): same d kappa/du, nothing else touched
    f = p.fields
    n, k, r, p4 = f["n"], f["k"], f["lmi_r"], f["lmib_p4"]
    v32 = np.ascontiguousarray(v, dtype=np.float32)
    y2, kap2, act2 = y.astype(np.float32), kap.copy(), act.copy()
    dk2 = np.full_like(dk, np.nan)
    rc = emu.emu_lmib_solve(_ptr(p.blob), n, k, r, p4, f["off_lmib"], f["off_y0"], _ptr(S), _ptr(v32), v32.shape[1], _ptr(y2),
                            _ptr(kap2), _ptr(act2, _I), _ptr(dk2), v32.shape[0], 1, 2, 64, 2, None, None, None)
    assert rc == 0
    np.testing.assert_array_equal(kap2, kap)
    np.testing.assert_array_equal(act2, act)
    np.testing.assert_array_equal(y2, y.astype(np.float32))
    both = ~np.isnan(dk).any(axis=1)
    assert both.any() and np.array_equal(~np.isnan(dk2).any(axis=1), both)
    np.testing.assert_array_equal(dk2[both], dk[both])


def test_big_lmi_violation_mode(emu):
    """lambda_max(-F(y)) through the same two kernels (the violation metric): the constant row enters as C0."""
    spec = synthetic.random_spec(k=3, r=34, seed=11)
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    f = p.fields
    k, r, p4 = f["k"], f["lmi_r"], f["lmib_p4"]
    rng = np.random.default_rng(0)
    y = rng.standard_normal((9, k)).astype(np.float32) * 0.3
    y[0] = 0.0                                                      # y0 = 0: strictly inside, violation 0
    Fn = np.ascontiguousarray(p.blob[f["off_lminegb"]:f["off_lminegb"] + (k + 1) * p4])
    S = np.full((9, p4), np.nan, dtype=np.float32)
    c0 = np.ascontiguousarray(Fn[k * p4:])
    assert emu.emu_lmib_contract(_ptr(y), k, _ptr(Fn), k, p4, _ptr(S), 9, _ptr(c0)) == 0
    viol = np.zeros(9, dtype=np.float32)
    rc = emu.emu_lmib_solve(_ptr(p.blob), k, k, r, p4, f["off_lminegb"], f["off_y0"], _ptr(S), _ptr(y), k, None, _ptr(viol), None,
                            None, 9, 0, 4, 64, 2, None, None, None)
    assert rc == 0
    allF = np.asarray(spec["lmi"])
    Fy = allF[-1][None] + np.einsum("bi,ijk->bjk", y.astype(np.float64), allF[:-1])
    ref = np.maximum(-np.linalg.eigvalsh(Fy)[:, 0], 0.0)
    assert viol[0] == 0.0 and (ref > 0).any()
    assert np.abs(viol - ref).max() <= 1e-5 * max(1.0, ref.max())


def test_big_lmi_definiteness_filter_is_exact(emu):
    """Priors placed around lambda_max (relative offsets 1e-2 ... 2e-5 on both sides): behind the Wolkowicz-Styan bound the
    LDL' filter may finish a sample only if lambda_max < prior in float64; whatever it decides, the outcome is the exact
    merge max(lambda_max, prior) with the right tag (a wrong pass would leave kappa = prior < lambda_max)."""
    spec = synthetic.random_spec(k=5, r=36, seed=21)
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    f = p.fields
    n, k, r, p4 = f["n"], f["k"], f["lmi_r"], f["lmib_p4"]
    B = 28
    v, _ = synthetic.sample_inputs(B, cs.n, cs.k, seed_v=5, scale=4.0)
    v = np.ascontiguousarray(v.numpy(), dtype=np.float32)
    s = np.linalg.norm(v.astype(np.float64), axis=1)
    u = v.astype(np.float64) / s[:, None]
    lam = np.linalg.eigvalsh(np.einsum("ba,aij->bij", u, p.f64["Fz"]))[:, -1]
    keep = lam > 0
    rng = np.random.default_rng(2)
    rel = rng.choice([1e-2, 1e-3, 2e-4, 2e-5, -2e-5, -2e-4, -1e-3, -1e-2], size=B)
    kprior = np.where(keep, lam * (1.0 + rel), 0.5).astype(np.float32)
    kap = kprior.copy()
    act = np.full(B, (1 << 24) | 3, dtype=np.int32)
    a_old = np.minimum(1.0 / kprior.astype(np.float64), s)
    rho = u @ cs.NA_E.T
    y = (cs.y0[:, 0][None] + a_old[:, None] * rho).astype(np.float32)
    Fp = np.ascontiguousarray(p.blob[f["off_lmib"]:f["off_lmib"] + n * p4])
    S = np.full((B, p4), np.nan, dtype=np.float32)
    assert emu.emu_lmib_contract(_ptr(v), n, _ptr(Fp), n, p4, _ptr(S), B, None) == 0
    dk = np.full((B, n), np.nan, dtype=np.float32)
    rc = emu.emu_lmib_solve(_ptr(p.blob), n, k, r, p4, f["off_lmib"], f["off_y0"], _ptr(S), _ptr(v), n, _ptr(y), _ptr(kap),
                            _ptr(act, _I), _ptr(dk), B, 0, 1, 64, 3, None, None, None)
    assert rc == 0
    want = np.maximum(np.maximum(lam, 0.0), kprior.astype(np.float64))
    assert np.abs(kap - want).max() <= 3e-6 * want.max()
    clear = np.abs(rel) > 1e-4
    binds = np.maximum(lam, 0.0) > kprior
    assert np.array_equal((act >> 24)[clear & keep], np.where(binds, 4, 1)[clear & keep])
    assert (binds & clear).any() and (~binds & clear).any()
    a_new = np.minimum(1.0 / want, s)
    y_want = cs.y0[:, 0][None] + a_new[:, None] * rho
    assert np.abs(y - y_want).max() <= 5e-6 * np.abs(y_want).max()
