"""Event timings of the wide kernels (n > 32, wide.cuh) through the C ABI, next to the oracle port on the host cores.
usage: python scripts/time_wide.py [k:m:eta:mu:r_M:eq:batch ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as B
from oracle.rayen_oracle import OracleSet, TorchOracle
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule

dev = torch.device("cuda", 0)
cases = sys.argv[1:] or ["64:256:4:4:32:0:4096", "64:256:4:4:32:0:65536", "128:512:8:8:64:8:4096", "256:1024:0:0:0:0:8192",
                         "1000:1000:0:0:0:0:2000", "1000:200:2:2:100:0:2000"]
out = []
for arg in cases:
    k, m, eta, mu, r_M, eq, batch = (int(x) for x in arg.split(":"))
    cs = synthetic.build_constraints(synthetic.wide_spec(k, m, eta, mu, r_M, eq, seed=1))
    layer = ConstraintModule(cs, create_map=False).to(dev)
    per_set = batch * 4 * (3 * layer.n + 2 * layer.k)
    db = B.DeviceBench(layer, batch, dev, pool=max(2, min(8, int(300e6 // per_set) + 1)))
    P = db.pool
    fwd = db.time_loop(lambda i: db.forward(db.sets[i % P], 1), 10, 3)
    bwd = db.time_loop(lambda i: db.backward(db.sets[i % P], 1), 10, 3)
    step = db.time_loop(db.step, 10, 3)
    act = db.sets[0]["active"].cpu().numpy() >> 24
    # oracle port on the host cores, bounded sample
    sample = min(batch, 256)
    oset = OracleSet.from_constraints(cs)
    orc = TorchOracle(oset, torch.float32)
    v, gy = synthetic.sample_inputs(sample, cs.n, cs.k)
    orc.forward_backward(v, gy)
    t0 = time.perf_counter()
    orc.forward_backward(v, gy)
    cpu_s = time.perf_counter() - t0
    rec = dict(k=k, n=cs.n, m=m, eta=eta, mu=mu, r_M=r_M, batch=batch, fwd_us=round(fwd * 1e3, 1), bwd_us=round(bwd * 1e3, 1),
               step_us=round(step * 1e3, 1), samples_per_s=round(batch / (step * 1e-3)), hbm_frac=round(per_set / (step * 1e-3) / 6550.1e9, 4),
               fam=np.bincount(act, minlength=4).tolist(), cpu_port_samples_per_s=round(sample / cpu_s), cpu_threads=torch.get_num_threads())
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del db, layer
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/wide_timings.json", "w"), indent=1)
