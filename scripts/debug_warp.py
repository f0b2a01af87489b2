import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.rayen_oracle import OracleSet, closed_form_numpy
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
loosen = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
spec = synthetic.config_spec("cfg5"); spec["b1"] = spec["b1"] * loosen
cs = synthetic.build_constraints(spec)
v, gy = synthetic.sample_inputs(B, cs.n, cs.k)
cf = closed_form_numpy(OracleSet.from_constraints(cs), v.numpy(), gy.numpy())
layer = ConstraintModule(cs, create_map=False).to("cuda:0")
x = v.to("cuda:0").requires_grad_(True)
y = layer(x.unsqueeze(2))[:, :, 0]
torch.cuda.synchronize()
kap, act = layer.last_kappa_and_active()
err = np.abs(y.detach().cpu().numpy() - cf["y"]).max(axis=1) / np.abs(cf["y"]).max()
bad = np.flatnonzero(err > 1e-5)
print("env", {k: os.environ[k] for k in os.environ if k.startswith("RAYEN")}, "B", B, "bad rows", len(bad), "of", B)
print("families (oracle):", np.bincount(cf["family"], minlength=5))
for b in bad[:8]:
    print(b, "err", err[b], "kappa gpu", float(kap[b]), "oracle", cf["kappa"][b], "tag", int(act[b]) >> 24, "fam", cf["family"][b], "y0..2", y[b, :3].tolist(), cf["y"][b, :3])
