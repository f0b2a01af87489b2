"""Tiny driver for ncu: a few forward/backward launches of one config (default cfg4, B=4096)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else synthetic.CONFIG_SHAPES[name]["batch"]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cs = synthetic.build_constraints(synthetic.config_spec(name))
layer = ConstraintModule(cs, create_map=False).cuda()
v, gy = synthetic.sample_inputs(B, cs.n, cs.k)
x = v.cuda().requires_grad_(True); g = gy.cuda()
for _ in range(iters):
    x.grad = None
    y = layer(x.unsqueeze(2)); y.backward(g.view(B, cs.k, 1))
torch.cuda.synchronize()
print("done", name, B)
