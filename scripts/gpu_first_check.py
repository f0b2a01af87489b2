"""Exploratory GPU check: errors of the CUDA path against every golden file and the fp64 oracle."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_golden
from rayen_b200 import synthetic, _cabi
from rayen_b200.constraint_module import ConstraintModule
from oracle.rayen_oracle import OracleSet, closed_form_numpy, max_violation

names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(ROOT, "tests/golden/*.npz")))
if len(sys.argv) > 1: names = sys.argv[1:]
for name in names:
    g = load_golden(name)
    cs = synthetic.build_constraints(g["spec"])
    try:
        layer = ConstraintModule(cs, create_map=False).cuda()
        x = torch.tensor(g["v"]).cuda().requires_grad_(True)
        y = layer(x.unsqueeze(2))
        (y[:, :, 0] * torch.tensor(g["gy"]).cuda()).sum().backward()
        torch.cuda.synchronize()
    except Exception as e:
        print(name, "FAILED", repr(e)); continue
    yv, gv = y[:, :, 0].detach().cpu().numpy().astype(np.float64), x.grad.cpu().numpy().astype(np.float64)
    oset = OracleSet.from_constraints(cs)
    cf = closed_form_numpy(oset, g["v"], g["gy"])
    ey32 = np.abs(yv - g["y32"]).max() / np.abs(g["y32"]).max()
    ey64 = np.abs(yv - g["y64"]).max() / np.abs(g["y64"]).max()
    fin = np.isfinite(g["gv64"]).all(axis=1) & np.isfinite(g["gv32"]).all(axis=1)   # the reference is NaN at v = 0 for quadratic sets
    ok = (cf["margin"] > 1e-4) & fin
    g["gv64"] = np.where(fin[:, None], g["gv64"], 0.0); g["gv32"] = np.where(fin[:, None], g["gv32"], 0.0)
    gv = np.where(fin[:, None], gv, 0.0)
    dg = np.abs(gv - g["gv64"]).max(axis=1)
    eg_all = dg.max() / np.abs(g["gv64"]).max()
    eg_ok = dg[ok].max() / np.abs(g["gv64"]).max()
    ref_eg = np.abs(g["gv32"] - g["gv64"]).max(axis=1)[ok].max() / np.abs(g["gv64"]).max()
    kap, act = layer.last_kappa_and_active()
    fam = np.bincount(act.cpu().numpy() >> 24, minlength=5)
    viol = max_violation(oset, yv, g["spec"]["A1"], g["spec"]["b1"], g["spec"]["A2"], g["spec"]["b2"])
    print(f"{name:16s} y:vs32 {ey32:.1e} vs64 {ey64:.1e} | g: all {eg_all:.1e} non-tie {eg_ok:.1e} (ref32 {ref_eg:.1e}) "
          f"| viol {viol:.1e} fam {fam} nan {np.isnan(yv).sum()+np.isnan(gv).sum()}", flush=True)
