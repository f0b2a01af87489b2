"""One-off stress run: many random constraint sets / batch sizes against the float64 oracle (same checks as
tests/test_gpu_parity.py::test_random_shapes_vs_oracle, more seeds, both methods, fused mapper on and off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle.rayen_oracle import OracleSet, TorchOracle, closed_form_numpy, max_violation
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule

dev = "cuda:0"
worst_y = worst_g = worst_v = 0.0
fails = 0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 60):
    rng = np.random.default_rng(1000 + seed)
    k = int(rng.integers(1, 33))
    spec = synthetic.random_spec(k=k, m=int(rng.integers(0, 80)), eta=int(rng.integers(0, 5)), mu=int(rng.integers(0, 5)),
                                 r_M=int(rng.integers(1, 2 * k + 1)), r=int(rng.integers(0, 2)) * int(rng.integers(2, 33)),
                                 seed=seed)
    if spec["A1"] is None and not spec["qcs"] and not spec["socs"] and spec["lmi"] is None:
        spec = synthetic.random_spec(k=k, m=5, seed=seed)
    if spec["b1"] is not None:
        spec["b1"] = spec["b1"] * float(rng.uniform(1.0, 6.0))
    if seed % 3 == 1 and k > 2:
        spec["A2"], spec["b2"] = rng.uniform(-1, 1, size=(1, k)), np.zeros((1, 1))
    cs = synthetic.build_constraints(spec)
    B = int(rng.choice([1, 7, 64, 300, 1111, 5000]))
    v, gy = synthetic.sample_inputs(B, cs.n, cs.k, seed_v=seed, scale=float(rng.uniform(0.5, 8.0)))
    method = "RAYEN_old" if seed % 4 == 3 else "RAYEN"
    if method == "RAYEN_old":   # one more input column: beta
        v = torch.cat((v, torch.randn(B, 1, generator=torch.Generator().manual_seed(seed))), dim=1)
    layer = ConstraintModule(cs, create_map=False, method=method).to(dev)
    x = v.to(dev).requires_grad_(True)
    y = layer(x.unsqueeze(2))
    (y[:, :, 0] * gy.to(dev)).sum().backward()
    yy, gg = y[:, :, 0].detach().cpu().double().numpy(), x.grad.cpu().double().numpy()
    oset = OracleSet.from_constraints(cs)
    y_ref, g_ref = TorchOracle(oset, torch.float64).forward_backward(v.double(), gy.double(), method)
    ok = closed_form_numpy(oset, v.numpy()[:, :cs.n], gy.numpy())["margin"] > 1e-4
    ey = float(np.abs(yy - y_ref.numpy()).max() / max(np.abs(y_ref.numpy()).max(), 1e-30))
    eg = float(np.abs(gg - g_ref.numpy())[ok].max() / max(np.abs(g_ref.numpy())[ok].max(), 1e-30)) if ok.any() else 0.0
    viol = max_violation(oset, yy, spec["A1"], spec["b1"], spec["A2"], spec["b2"]) / max(1.0, np.abs(yy).max())
    worst_y, worst_g, worst_v = max(worst_y, ey), max(worst_g, eg), max(worst_v, viol)
    bad = ey > 1e-5 or eg > 2e-5 or viol > 1e-5 or not np.isfinite(yy).all() or not np.isfinite(gg).all()
    fails += bad
    if bad:
        print("FAIL seed", seed, method, dict(k=k, n=cs.n, B=B, ey=ey, eg=eg, viol=viol), flush=True)
# (near-tangent cone rays are ill-conditioned in float32 by nature: see DESIGN.md section 3)
print(f"done: worst rel err y {worst_y:.2e}, g_v {worst_g:.2e}, violation {worst_v:.2e}, failures {fails}")
