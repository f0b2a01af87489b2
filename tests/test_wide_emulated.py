"""The wide kernels (rayen_b200/csrc/wide.cuh, n > 32) compiled for the HOST under a SIMT emulator (tests/emu: one OS
thread per CUDA thread, barriers for __syncthreads/__syncwarp, shuffles through a per-warp exchange buffer) and checked
against the oracle.  This is test infrastructure only: it exercises the kernels' indexing, task walk, barrier and
shuffle structure on the CPU, where the GPU is not available; the GPU parity tests are in test_gpu_parity.py.  The
product never loads this library."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle.rayen_oracle import OracleSet, TorchOracle, closed_form_numpy, max_violation
from rayen_b200 import _cabi, plan, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int32)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = tmp_path_factory.mktemp("wide_emu")
    src = open(os.path.join(_cabi.CSRC, "wide.cuh")).read().splitlines(True)
    kept = [l for l in src if l.strip() not in ('#include "common.cuh"', '#include "lqs.cuh"')]
    assert len(kept) == len(src) - 2
    with open(out / "wide_stripped.cuh", "w") as fh:
        fh.writelines(kept)
    lib_path = str(out / "libwide_emu.so")
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC", "-w", f"-I{out}", f"-I{os.path.join(HERE, 'emu')}",
           "-o", lib_path, os.path.join(HERE, "emu", "wide_emu.cpp")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    lib = ctypes.CDLL(lib_path)
    lib.emu_wide_forward.restype = ctypes.c_int
    lib.emu_wide_forward.argtypes = [_F, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F,
                                     ctypes.c_longlong, _F, _F, _I, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.emu_wide_backward.restype = ctypes.c_int
    lib.emu_wide_backward.argtypes = [_F, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F,
                                      ctypes.c_longlong, _F, _F, _I, _F, ctypes.c_longlong, ctypes.c_longlong,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return lib


def _ptr(a, t=_F):
    return a.ctypes.data_as(t)


def run_emulated(lib, p, v, gy, mode, ts=8, grid_f=3, grid_b=5, bthreads=None):
    f = p.fields
    n, k = f["n"], f["k"]
    v = np.ascontiguousarray(v, dtype=np.float32)
    gy = np.ascontiguousarray(gy, dtype=np.float32)
    B, cols = v.shape
    y = np.full((B, k), np.nan, dtype=np.float32)
    kap = np.full((B,), np.nan, dtype=np.float32)
    act = np.full((B,), -1, dtype=np.int32)
    gv = np.full((B, cols), np.nan, dtype=np.float32)
    blob = p.blob
    rc = lib.emu_wide_forward(_ptr(blob), f["off_wide"], n, k, f["off_y0"], f["n_is_identity"], _ptr(v), cols, _ptr(y),
                              _ptr(kap), _ptr(act, _I), B, mode, grid_f, ts)
    assert rc == 0
    rc = lib.emu_wide_backward(_ptr(blob), f["off_wide"], n, k, f["off_y0"], f["n_is_identity"], _ptr(v), cols, _ptr(gy),
                               _ptr(kap), _ptr(act, _I), _ptr(gv), cols, B, mode, grid_b, bthreads or (128 if ts == 8 else 256))
    assert rc == 0
    return y.astype(np.float64), kap, act, gv.astype(np.float64)


def rel(a, b, rows=None):
    if rows is not None:
        a, b = a[rows], b[rows]
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


CASES = [
    # k, m, eta, mu, r_M, eq, batch, method
    (40, 50, 2, 2, 20, 0, 37, "RAYEN"),
    (45, 70, 3, 2, 50, 3, 26, "RAYEN"),          # equalities: N is not the identity, n = 42
    (65, 0, 1, 1, 5, 0, 17, "RAYEN"),            # no linear rows
    (33, 7, 0, 0, 0, 0, 9, "RAYEN"),
    (70, 300, 0, 0, 0, 0, 16, "RAYEN"),          # several linear tasks per warp
    (36, 40, 9, 10, 12, 2, 3, "RAYEN"),          # more items than warps
    (34, 10, 40, 40, 3, 0, 11, "RAYEN"),         # 80 items: two rounds
    (4096, 40, 0, 0, 0, 0, 5, "RAYEN"),          # the widest set the kernels take
    (40, 50, 2, 2, 20, 0, 21, "RAYEN_old"),
    (45, 30, 1, 1, 16, 3, 10, "RAYEN_old"),
]


@pytest.mark.parametrize("ts", [4, 8, 16])
@pytest.mark.parametrize("k,m,eta,mu,r_M,eq,batch,method", CASES)
def test_wide_kernels_under_the_emulator_match_the_oracle(emu, k, m, eta, mu, r_M, eq, batch, method, ts):
    if ts == 4 and k not in (36, 45, 70):
        pytest.skip("tiles of 4 samples (the build for n > ~6900) are exercised on three of the sets")
    if k == 4096 and ts == 16:
        pytest.skip("tiles of 16 samples of a 4096-dimensional set exceed the shared memory of an SM (the launcher uses 8)")
    spec = synthetic.wide_spec(k, m, eta, mu, r_M, eq, seed=k + batch)
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    assert p.fields["wide"] == 1
    old = method == "RAYEN_old"
    v, gy = synthetic.sample_inputs(batch, cs.n + (1 if old else 0), cs.k, seed_v=batch, seed_g=k)
    if batch > 4 and not old:
        v[2] = 0.0
        v[3] *= 1e-3
    y, kap, act, gv = run_emulated(emu, p, v.numpy(), gy.numpy(), _cabi.MODE_RAYEN_OLD if old else _cabi.MODE_RAYEN, ts=ts)
    assert np.isfinite(y).all() and np.isfinite(gv).all() and np.isfinite(kap).all() and (act >= 0).all()
    oset = OracleSet.from_constraints(cs)
    y_ref, g_ref = TorchOracle(oset, torch.float64).forward_backward(v.double(), gy.double(), method=method)
    assert rel(y, y_ref.numpy()) <= 1e-5
    vv = v.numpy()[:, :cs.n].astype(np.float64)
    cf = closed_form_numpy(oset, vv, gy.numpy())
    ok = (cf["margin"] > 1e-4) & (np.linalg.norm(vv, axis=1) > 0) & np.isfinite(g_ref.numpy()).all(axis=1)
    assert ok.sum() >= 0.6 * len(ok)
    assert rel(gv, g_ref.numpy(), ok) <= 2e-5
    assert max_violation(oset, y, spec["A1"], spec["b1"], spec["A2"], spec["b2"]) <= 1e-5 * max(1.0, np.abs(y).max())
    _, kap_np, act_np = plan.evaluate_wide_numpy(p, vv)
    assert np.abs(kap - kap_np).max() <= 1e-5 * max(1.0, kap_np.max())
    assert (act[ok] == act_np[ok]).all()
    fams = set((act >> 24).tolist())
    assert fams <= {0, 1, 2, 3}
    if batch > 4 and not old:
        np.testing.assert_allclose(y[2], cs.y0[:, 0], atol=1e-6)
        assert np.all(gv[2] == 0)


# ----------------------------------------------------------------------------- ThreadSanitizer: the CPU racecheck
@pytest.fixture(scope="module", params=["thread", "address"])
def tsan_driver(request, tmp_path_factory):
    """The stand-alone emulated build under ThreadSanitizer (the racecheck) or AddressSanitizer + UBSan (a memcheck of the
    global-memory side: the constant block, v, g_y, y, kappa, active, g_v are exact-size heap buffers)."""
    out = tmp_path_factory.mktemp("wide_" + request.param)
    src = open(os.path.join(_cabi.CSRC, "wide.cuh")).read().splitlines(True)
    kept = [l for l in src if l.strip() not in ('#include "common.cuh"', '#include "lqs.cuh"')]
    with open(out / "wide_stripped.cuh", "w") as fh:
        fh.writelines(kept)
    exe = str(out / "wide_emu_san")
    flag = "-fsanitize=thread" if request.param == "thread" else "-fsanitize=address,undefined"
    cmd = ["g++", "-std=c++20", "-O1", "-g", flag, "-pthread", "-w", f"-I{out}", f"-I{os.path.join(HERE, 'emu')}",
           "-o", exe, os.path.join(HERE, "emu", "wide_emu_main.cpp")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0 and "san" in (proc.stderr or "").lower() and "cannot find" in (proc.stderr or "").lower():
        pytest.skip("sanitizer runtime not available: " + proc.stderr[-200:])
    assert proc.returncode == 0, proc.stderr
    return exe, out


@pytest.mark.parametrize("k,m,eta,mu,r_M,eq,batch,method,ts,bthreads", [
    (45, 70, 3, 2, 50, 3, 19, "RAYEN", 8, 128),
    (40, 50, 2, 2, 20, 0, 21, "RAYEN_old", 16, 256),
    (34, 10, 40, 40, 3, 0, 9, "RAYEN", 16, 128),      # two rounds: the slot scratch is reused
])
def test_wide_kernels_are_clean_under_thread_and_address_sanitizer(emu, tsan_driver, k, m, eta, mu, r_M, eq, batch, method, ts, bthreads):
    """The kernels' barrier structure, checked like compute-sanitizer's racecheck would: under the emulator every CUDA
    thread is an OS thread, shared memory is ordinary memory, __syncthreads/__syncwarp are barriers -- so a missing
    barrier in wide.cuh is a data race that ThreadSanitizer reports.  Several tiles / samples per block (grid-stride)
    exercise the reuse of the shared-memory buffers.  Outputs must equal the plain emulated build's, bit for bit."""
    exe, out = tsan_driver
    cs = synthetic.build_constraints(synthetic.wide_spec(k, m, eta, mu, r_M, eq, seed=k + batch))
    p = plan.build_plan_from_constraints(cs)
    f = p.fields
    old = method == "RAYEN_old"
    mode = _cabi.MODE_RAYEN_OLD if old else _cabi.MODE_RAYEN
    v, gy = synthetic.sample_inputs(batch, cs.n + (1 if old else 0), cs.k, seed_v=batch, seed_g=k)
    v32, gy32 = np.ascontiguousarray(v.numpy(), dtype=np.float32), np.ascontiguousarray(gy.numpy(), dtype=np.float32)
    B, cols = v32.shape
    hdr = np.asarray([f["n"], f["k"], f["off_y0"], f["n_is_identity"], f["off_wide"], p.blob.size, B, cols, mode, ts, bthreads,
                      2, 3, 0, 0, 0], dtype=np.int64)
    fin, fout = str(out / f"in_{k}_{batch}.bin"), str(out / f"out_{k}_{batch}.bin")
    with open(fin, "wb") as fh:
        fh.write(hdr.tobytes()); fh.write(p.blob.tobytes()); fh.write(v32.tobytes()); fh.write(gy32.tobytes())
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0", ASAN_OPTIONS="detect_leaks=0")
    proc = subprocess.run([exe, fin, fout], capture_output=True, text=True, env=env, timeout=900)
    assert "Sanitizer" not in proc.stderr and "runtime error" not in proc.stderr, proc.stderr[:3000]
    assert proc.returncode == 0, (proc.returncode, proc.stderr[:500])
    raw = np.fromfile(fout, dtype=np.float32)
    y = raw[:B * cs.k].reshape(B, cs.k)
    kap = raw[B * cs.k:B * cs.k + B]
    act = raw[B * cs.k + B:B * cs.k + 2 * B].view(np.int32)
    gv = raw[B * cs.k + 2 * B:].reshape(B, cols)
    y2, kap2, act2, gv2 = run_emulated(emu, p, v32, gy32, mode, ts=ts, grid_f=2, grid_b=3, bthreads=bthreads)
    np.testing.assert_array_equal(y.astype(np.float64), y2)
    np.testing.assert_array_equal(kap, kap2)
    np.testing.assert_array_equal(act, act2)
    np.testing.assert_array_equal(gv.astype(np.float64), gv2)
