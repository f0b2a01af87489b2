import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import load_golden
from rayen_b200 import synthetic
from rayen_b200.constraint_module import ConstraintModule
g = load_golden("old_mixed8")
cs = synthetic.build_constraints(g["spec"])
print("k", cs.k, "n", cs.n, "lmi r", g["spec"]["lmi"][0].shape if g["spec"]["lmi"] is not None else None)
layer = ConstraintModule(cs, method="RAYEN_old", create_map=False).to("cuda:0")
x = torch.tensor(g["v"], dtype=torch.float32, device="cuda:0").requires_grad_(True)
y = layer(x.unsqueeze(2))
(y[:, :, 0] * torch.tensor(g["gy"], dtype=torch.float32, device="cuda:0")).sum().backward()
gv = x.grad.cpu().double().numpy()
kap, act = layer.last_kappa_and_active()
fam = (act.cpu().numpy() >> 24)
err = np.abs(gv - g["gv64"]).max(axis=1) / np.abs(g["gv64"]).max()
for f in range(5):
    m = fam == f
    if m.any():
        print("family", f, "count", int(m.sum()), "max rel err", float(err[m].max()), "bad rows", np.nonzero(m & (err > 1e-4))[0][:10])
bad = np.nonzero(err > 1e-4)[0]
print("bad total", len(bad), "of", len(err))
for b in bad[:3]:
    print(b, "fam", fam[b], "\n got", gv[b], "\n ref", g["gv64"][b])
