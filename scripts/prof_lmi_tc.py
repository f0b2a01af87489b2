"""ncu driver for lmi_forward_tc_kernel: dense LMI-only set with K = n = 32 (where the tensor-core contraction is
the default), no-grad forward launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cs = synthetic.build_constraints(synthetic.random_spec(k=32, r=32, seed=5))
layer = ConstraintModule(cs, create_map=False).cuda()
v, _ = synthetic.sample_inputs(B, cs.n, cs.k)
x = v.cuda()
with torch.no_grad():
    for _ in range(3):
        y = layer(x.unsqueeze(2))
torch.cuda.synchronize()
print("done", B)
