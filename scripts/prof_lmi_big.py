"""Tiny driver for ncu: a few forward/backward launches of a big-LMI set (default r = 100, n = 100, B = 2000)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
r = int(sys.argv[1]) if len(sys.argv) > 1 else 100
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
cs = synthetic.build_constraints(synthetic.random_spec(k=k, r=r, seed=6))
layer = ConstraintModule(cs, create_map=False).cuda()
v, gy = synthetic.sample_inputs(B, cs.n, cs.k)
x = v.cuda().requires_grad_(True); g = gy.cuda()
for _ in range(3):
    x.grad = None
    y = layer(x.unsqueeze(2)); y.backward(g.view(B, cs.k, 1))
torch.cuda.synchronize()
print("done", r, k, B)
