"""Timing of the big-LMI path (lmi_big.cuh) on the LMI grid of the reference's sweep (examples/scripts/time_analysis.py:159-175:
r_F in {10, 100, 200, 300} x k in {100, 500, 1000, 2000, ...}; k <= 4096 here), 2000 samples as there, LMI-only sets."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as B
from rayen_b200 import synthetic, _cabi
from rayen_b200.constraint_module import ConstraintModule
from oracle.rayen_oracle import OracleSet, TorchOracle

dev = torch.device("cuda", 0)
NUM = 2000
points = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(10, 100), (100, 100), (200, 100), (300, 100)]
out = []
for r, k in points:
    t0 = time.time()
    rng = np.random.default_rng(r * 10007 + k)
    all_F = []
    for _ in range(k):
        T = rng.uniform(-1.0, 1.0, size=(r, r))
        all_F.append(0.5 * (T + T.T))
    T = rng.uniform(-1.0, 1.0, size=(r, r))
    all_F.append(T @ T.T + 0.5 * np.eye(r))
    spec = dict(A1=None, b1=None, A2=None, b2=None, qcs=[], socs=[], lmi=all_F, y0=np.zeros((k, 1)))
    cs = synthetic.build_constraints(spec)
    layer = ConstraintModule(cs, create_map=False).to(dev)
    host_s = time.time() - t0
    db = B.DeviceBench(layer, NUM, dev, pool=2)
    db.want_grad = 0
    before = _cabi.launch_count()
    fwd = db.time_loop(lambda i: db.forward(db.sets[i % 2]), 5, 2)
    launches = (_cabi.launch_count() - before) // 7
    db.want_grad = 1
    step = db.time_loop(db.step, 5, 2)
    s0 = db.sets[0]
    viol = float(layer.violation(s0["y"]).max())
    rec = dict(r_F=r, k=k, n=int(cs.n), fwd_ms=round(fwd, 3), fwd_s_per_sample=fwd * 1e-3 / NUM, fwd_bwd_ms=round(step, 3),
               launches_per_forward=int(launches), max_violation=viol, host_setup_s=round(host_s, 1),
               gemm_gflop=2.0 * NUM * cs.n * (r * (r + 1) // 2) / 1e9, eig_gflop=NUM * 2.0 * r ** 3 / 1e9)
    if r * r * k <= 100 * 100 * 500:
        oset = OracleSet.from_constraints(cs)
        orc = TorchOracle(oset, torch.float64)
        sub = s0["v"][:32].cpu().double()
        y_ref = orc.forward(sub.unsqueeze(2))[:, :, 0]
        rec["rel_err_y_vs_oracle"] = float((s0["y"][:32].cpu().double() - y_ref).abs().max() / float(y_ref.abs().max()))
        v64 = s0["v"].cpu().double().unsqueeze(2)
        with torch.no_grad():
            t1 = time.perf_counter()
            orc.forward(v64[:256])
            rec["cpu_port_s_per_sample"] = (time.perf_counter() - t1) / 256
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del db, layer
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/lmi_big_timings.json", "w"), indent=1)
