import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch, bench as B
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
for rows, k in ((1, 1000), (100, 1000), (1, 4000), (10, 10000), (100, 10000), (500, 10000), (1000, 1000)):
    rng = np.random.default_rng(rows + k)
    spec = dict(A1=rng.uniform(-1.0, 1.0, size=(rows, k)), b1=rng.uniform(0.1, 1.0, size=(rows, 1)), A2=None, b2=None, qcs=[], socs=[], lmi=None, y0=np.zeros((k, 1)))
    cs = synthetic.build_constraints(spec)
    layer = ConstraintModule(cs, create_map=False).to(dev)
    db = B.DeviceBench(layer, 2000, dev, pool=2)
    db.want_grad = 0
    fwd = db.time_loop(lambda i: db.forward(db.sets[i % 2]), 10, 3)
    db.want_grad = 1
    step = db.time_loop(db.step, 10, 3)
    viol = float(layer.violation(db.sets[0]["y"]).max())
    print(dict(rows=rows, k=k, fwd_us=round(fwd * 1e3, 1), step_us=round(step * 1e3, 1), viol=viol), flush=True)
