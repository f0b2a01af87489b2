"""Tiny driver for ncu: a few forward/backward launches of a wide linear set (default 100 rows in dimension 10000, B = 2000)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
rng = np.random.default_rng(10)
spec = dict(A1=rng.uniform(-1.0, 1.0, size=(rows, k)), b1=rng.uniform(0.1, 1.0, size=(rows, 1)), A2=None, b2=None, qcs=[], socs=[],
            lmi=None, y0=np.zeros((k, 1)))
cs = synthetic.build_constraints(spec)
layer = ConstraintModule(cs, create_map=False).cuda()
v, gy = synthetic.sample_inputs(B, cs.n, cs.k, scale=1.0)
x = v.cuda().requires_grad_(True); g = gy.cuda()
for _ in range(3):
    x.grad = None
    y = layer(x.unsqueeze(2)); y.backward(g.view(B, cs.k, 1))
torch.cuda.synchronize()
print("done", rows, k, B)
