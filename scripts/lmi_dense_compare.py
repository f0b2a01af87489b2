"""LMI forward kernel, tensor-core vs FP32-pipe contraction, on dense LMI-only sets of different K (= n)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
for k, r, batch in ((8, 32, 65536), (16, 32, 65536), (32, 32, 65536), (32, 16, 65536), (32, 32, 4096)):
    cs = synthetic.build_constraints(synthetic.random_spec(k=k, r=r, seed=5))
    out = {}
    for tc in (1, 0):
        layer = ConstraintModule(cs, create_map=False).to(dev)
        layer.set_lmi_tensor_cores(bool(tc), device=dev)
        db = B.DeviceBench(layer, batch, dev, pool=2)
        for wg in (1, 0):
            db.want_grad = wg
            out[f"tc{tc}_grad{wg}"] = round(db.time_loop(lambda i: db.forward(db.sets[i % 2]), 10, 3) * 1e3, 1)
        del db, layer
        torch.cuda.empty_cache()
    print(f"k=n={k} r={r} B={batch}: forward us", out, flush=True)
