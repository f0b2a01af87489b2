"""End-to-end (host buffers) step time against the chunk count of the copy/compute pipeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
for arg in sys.argv[1:]:
    name, batch = arg.split(":"); batch = int(batch)
    cs = synthetic.build_constraints(synthetic.config_spec(name))
    out = {}
    for chunks in (1, 2, 3, 4, 6, 8, 0):
        os.environ["RAYEN_HOST_CHUNKS"] = str(chunks)
        layer = ConstraintModule(cs, create_map=False).to(dev)
        db = B.DeviceBench(layer, batch, dev, pool=4)
        fn, host = B.e2e_step_fn(layer, db, dev)
        out[chunks] = round(db.time_loop(fn, 20, 4) * 1e3, 1)
        del db, layer, fn, host
        torch.cuda.empty_cache()
    print(name, batch, "e2e us per step by chunk count (0 = automatic):", out, flush=True)
