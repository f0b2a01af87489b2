"""Randomised CPU stress of the wide kernels (wide.cuh) under the host SIMT emulator (tests/emu) against the float64 oracle:\nboundary shapes (rows / dimensions at 63, 64, 65, 127, 128, 129 ...), no linear rows, equalities, batches 1..33, both tile\nsizes, both backward widths, both methods.  usage: python scripts/stress_wide_emulated.py <first seed> <end seed>"""
import sys, os, ctypes, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import test_wide_emulated as T
from rayen_b200 import _cabi, plan, synthetic
from oracle.rayen_oracle import OracleSet, TorchOracle, closed_form_numpy, max_violation
# build the emulated lib like the fixture
out = tempfile.mkdtemp()
src = open(os.path.join(_cabi.CSRC, "wide.cuh")).read().splitlines(True)
kept = [l for l in src if l.strip() not in ('#include "common.cuh"', '#include "lqs.cuh"')]
open(os.path.join(out,"wide_stripped.cuh"),"w").writelines(kept)
lib_path = os.path.join(out,"libwide_emu.so")
subprocess.run(["g++","-std=c++20","-O1","-pthread","-shared","-fPIC","-w",f"-I{out}",f"-I{os.path.join(T.HERE,'emu')}","-o",lib_path,os.path.join(T.HERE,"emu","wide_emu.cpp")],check=True)
lib = ctypes.CDLL(lib_path)
lib.emu_wide_forward.restype = ctypes.c_int
lib.emu_wide_forward.argtypes = [T._F, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, T._F, ctypes.c_longlong, T._F, T._F, T._I, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int]
lib.emu_wide_backward.restype = ctypes.c_int
lib.emu_wide_backward.argtypes = [T._F, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, T._F, ctypes.c_longlong, T._F, T._F, T._I, T._F, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int]
worst_y = worst_g = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(5000 + seed)
    k = int(rng.choice([33, 34, 63, 64, 65, 66, 96, 127, 128, 129, 130, 160, 200, int(rng.integers(33, 260))]))
    eq = int(rng.choice([0, 0, 1, 2, 5]))
    eq = min(eq, k - 33)
    m = int(rng.choice([0, 1, 2, 63, 64, 65, 127, 128, 129, int(rng.integers(1, 400))]))
    eta = int(rng.choice([0, 1, 2, 5, 9])); mu = int(rng.choice([0, 1, 2, 5, 9]))
    if m == 0 and eta + mu == 0: m = 7
    if eq and m == 0: m = 3
    r_M = int(rng.integers(1, 2 * k))
    batch = int(rng.choice([1, 2, 7, 8, 9, 15, 16, 17, 31, 33]))
    method = "RAYEN_old" if seed % 4 == 3 else "RAYEN"
    ts = 8 if seed % 2 == 0 else 16
    spec = synthetic.wide_spec(k, m, eta, mu, r_M, eq, seed=seed, loosen=float(rng.choice([1.0, 3.0, 6.0])))
    cs = synthetic.build_constraints(spec)
    p = plan.build_plan_from_constraints(cs)
    old = method == "RAYEN_old"
    v, gy = synthetic.sample_inputs(batch, cs.n + (1 if old else 0), cs.k, seed_v=seed, seed_g=seed + 1, scale=float(rng.choice([0.2, 2.0, 20.0])))
    y, kap, act, gv = T.run_emulated(lib, p, v.numpy(), gy.numpy(), _cabi.MODE_RAYEN_OLD if old else _cabi.MODE_RAYEN, ts=ts, grid_f=int(rng.integers(1,4)), grid_b=int(rng.integers(1,5)), bthreads=int(rng.choice([128,256])))
    oset = OracleSet.from_constraints(cs)
    y_ref, g_ref = TorchOracle(oset, torch.float64).forward_backward(v.double(), gy.double(), method=method)
    vv = v.numpy()[:, :cs.n].astype(np.float64)
    cf = closed_form_numpy(oset, vv, gy.numpy())
    ok = (cf["margin"] > 1e-4) & (np.linalg.norm(vv, axis=1) > 0) & np.isfinite(g_ref.numpy()).all(axis=1)
    ey = T.rel(y, y_ref.numpy()); eg = T.rel(gv, g_ref.numpy(), ok) if ok.any() else 0.0
    viol = max_violation(oset, y, spec["A1"], spec["b1"], spec["A2"], spec["b2"]) / max(1.0, np.abs(y).max())
    fam = np.bincount(act >> 24, minlength=4).tolist()
    flag = "" if (ey <= 1e-5 and eg <= 2e-5 and viol <= 1e-5 and np.isfinite(y).all() and np.isfinite(gv).all()) else "  <<<<<< FAIL"
    worst_y, worst_g = max(worst_y, ey), max(worst_g, eg)
    print(f"seed {seed}: k={k} n={cs.n} m={m} eta={eta} mu={mu} rM={r_M} B={batch} {method} ts={ts} fam={fam} ey={ey:.1e} eg={eg:.1e} viol={viol:.1e}{flag}", flush=True)
print("worst", worst_y, worst_g)
