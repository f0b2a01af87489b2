import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rayen_b200 import synthetic, _cabi
from rayen_b200.constraint_module import ConstraintModule
from oracle.rayen_oracle import OracleSet, TorchOracle
DEV = "cuda:0"
for loosen in (1.0, 4.0):
    spec = synthetic.config_spec("cfg5"); spec["b1"] = spec["b1"] * loosen
    cs = synthetic.build_constraints(spec)
    v, gy = synthetic.sample_inputs(3000, cs.n, cs.k)
    outs = []
    for enabled in (True, False):
        layer = ConstraintModule(cs, create_map=False).to(DEV)
        layer.set_pruning(enabled, device=DEV)
        x = v.to(DEV).requires_grad_(True)
        y = layer(x.unsqueeze(2)); (y[:, :, 0] * gy.to(DEV)).sum().backward()
        kap, act = layer.last_kappa_and_active()
        outs.append((y.detach().cpu()[:, :, 0], x.grad.cpu(), kap.cpu(), act.cpu()))
    names = ["y", "gv", "kappa", "active"]
    for nm, a, b in zip(names, outs[0], outs[1]):
        d = (a.double() - b.double()).abs()
        print(loosen, nm, "equal", torch.equal(a, b), "max diff", float(d.max()), "n diff rows", int((d.reshape(3000, -1).max(dim=1).values > 0).sum()))
    s = v.norm(dim=1); kap = outs[0][2]; fam = outs[0][3] >> 24
    bnd = (1.0 / kap < s)
    print("  lmi-bound", int((fam == 4).sum()), "lmi-bound & boundary", int(((fam == 4) & bnd).sum()), "boundary total", int(bnd.sum()))
    oset = OracleSet.from_constraints(cs)
    y_ref, g_ref = TorchOracle(oset, torch.float64).forward_backward(v.double(), gy.double())
    for i, tag in enumerate(["pruned", "dense"]):
        print("  ", tag, "rel err y", float((outs[i][0].double() - y_ref).abs().max() / y_ref.abs().max()),
              "gv", float((outs[i][1].double() - g_ref).abs().max() / g_ref.abs().max()))
