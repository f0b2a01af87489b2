import sys, os
sys.path.insert(0, "/root/repo")
import torch, bench as B
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
for (k, m, r, batch, loosen) in ((32, 64, 64, 8192, 3.0), (16, 40, 100, 4000, 3.0), (64, 128, 33, 8192, 3.0)):
    spec = synthetic.random_spec(k=k, m=m, r=r, seed=5) if k <= 32 else synthetic.wide_spec(k, m, 2, 2, 16, 0, seed=7, r=r)
    if k <= 32: spec["b1"] = spec["b1"] * loosen
    cs = synthetic.build_constraints(spec)
    layer = ConstraintModule(cs, create_map=False).to(dev)
    db = B.DeviceBench(layer, batch, dev, pool=2)
    db.want_grad = 1
    step = db.time_loop(db.step, 10, 3)
    db.want_grad = 0
    fwd = db.time_loop(lambda i: db.forward(db.sets[i % 2]), 10, 3)
    kap, act = layer.last_kappa_and_active() if hasattr(layer, "_last_aux") and layer._last_aux else (None, None)
    fam = torch.bincount(db.sets[0]["active"] >> 24, minlength=5).tolist() if "active" in db.sets[0] else None
    print(dict(k=k, m=m, r=r, B=batch, fwd_ms=round(fwd, 3), step_ms=round(step, 3), fam=fam), flush=True)
