"""Per-kernel event timings for one config (uses bench.DeviceBench)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
for arg in sys.argv[1:]:
    name, batch = arg.split(":"); batch = int(batch)
    loose = name.endswith("L")
    spec = synthetic.config_spec(name.rstrip("L"))
    if loose: spec["b1"] = spec["b1"] * 4.0
    cs = synthetic.build_constraints(spec)
    layer = ConstraintModule(cs, create_map=False).to(dev)
    per_set = batch * 4 * (3 * layer.n + 2 * layer.k)
    db = B.DeviceBench(layer, batch, dev, pool=max(2, min(16, int(300e6 // per_set) + 1)))
    for i in range(db.pool): db.forward(db.sets[i])
    out = {}
    P = db.pool
    out["lqs_bwd"] = db.time_loop(lambda i: db.backward(db.sets[i % P], 1), 20, P)
    if cs.has_lmi_constraints:
        out["lmi_bwd"] = db.time_loop(lambda i: db.backward(db.sets[i % P], 2), 20, P)
    out["lqs_fwd"] = db.time_loop(lambda i: db.forward(db.sets[i % P], 1), 20, P)
    if cs.has_lmi_constraints:
        out["lmi_fwd"] = db.time_loop(lambda i: db.forward(db.sets[i % P], 2), 20, P)
    out["step"] = db.time_loop(db.step, 20, 3)
    db.want_grad = 0
    out["fwd_nograd"] = db.time_loop(lambda i: db.forward(db.sets[i % P]), 20, 3)
    db.want_grad = 1
    act = db.sets[0]["active"].cpu().numpy() >> 24
    import numpy as np
    print(name, batch, {k: round(v * 1e3, 1) for k, v in out.items()}, "us; fam", np.bincount(act, minlength=5).tolist(), flush=True)
    del db, layer; torch.cuda.empty_cache()
