"""e2e (host buffers) step time of cfg5 for the environment given on the command line, e.g. RAYEN_HOST_CHUNKS=2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
cs = synthetic.build_constraints(synthetic.config_spec("cfg5"))
layer = ConstraintModule(cs, create_map=False).to(dev)
db = B.DeviceBench(layer, 32768, dev, pool=4)
fn, host = B.e2e_step_fn(layer, db, dev)
ms = db.time_loop(fn, 40, 8)
print({k: v for k, v in os.environ.items() if k.startswith("RAYEN")}, "e2e ms/step %.4f" % ms)
