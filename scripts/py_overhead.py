"""Host-side cost of the nn.Module path (forward + autograd backward), per step, by function (cProfile)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
cs = synthetic.build_constraints(synthetic.config_spec("cfg2"))
layer = ConstraintModule(cs, create_map=False).to(dev)
v, gy = synthetic.sample_inputs(256, cs.n, cs.k)
x = v.to(dev).requires_grad_(True); g = gy.to(dev).view(256, cs.k, 1)
def step():
    x.grad = None
    layer(x.unsqueeze(2)).backward(g)
for _ in range(50): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): step()
torch.cuda.synchronize()
print("host time per fwd+bwd step (tiny batch, GPU work ~10 us): %.1f us" % ((time.perf_counter() - t0) / 2000 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
