"""Sweep the launch geometry (samples/thread TM x lanes/sample L) of the LQS forward kernel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
cases = [("cfg5", 32768), ("cfg5", 262144), ("cfg3", 16384), ("cfg3", 262144), ("cfg2", 4096), ("cfg2", 262144), ("cfg2", 2097152)]
if len(sys.argv) > 1:
    cases = [(sys.argv[1], int(sys.argv[2]))]
for name, batch in cases:
    cs = synthetic.build_constraints(synthetic.config_spec(name))
    layer = ConstraintModule(cs, create_map=False).to(dev)
    per_set = batch * 4 * (3 * layer.n + 2 * layer.k)
    db = B.DeviceBench(layer, batch, dev, pool=max(2, min(16, int(300e6 // per_set) + 1)))
    res = {}
    for tm in (0, 1, 2, 4):
        for lanes in ((0,) if tm == 0 else (1, 2, 4, 8, 16, 32)):
            layer.set_tuning(tm, lanes, device=dev)
            try:
                ms = db.time_loop(lambda i: db.forward(db.sets[i % db.pool], 1), 20, 3)
            except Exception as e:
                ms = float("nan")
            res[f"tm{tm}_L{lanes}"] = round(ms * 1e3, 1)
    best = min((v, k) for k, v in res.items() if v == v)
    print(name, batch, "best", best, json.dumps(res), flush=True)
    del db, layer
    torch.cuda.empty_cache()
