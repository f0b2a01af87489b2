import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from rayen_b200 import synthetic, _cabi
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
cs = synthetic.build_constraints(synthetic.config_spec("cfg5"))
layer = ConstraintModule(cs, create_map=False).to(dev)
db = B.DeviceBench(layer, 32768, dev, pool=4)
for i in range(6):
    db.forward(db.sets[i % 4], 1)
torch.cuda.synchronize()
db.forward(db.sets[0], 1)
_cabi.lib().rayen_tc_trace_dump()
