"""Phase trace of lqs_backward_kernel (development build with -DRAYEN_BWD_TRACE, see scripts/lmi_trace.py)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as B
from rayen_b200 import synthetic, _cabi
from rayen_b200.constraint_module import ConstraintModule
dev = torch.device("cuda", 0)
for arg in sys.argv[1:]:
    name, batch = arg.split(":"); batch = int(batch)
    cs = synthetic.build_constraints(synthetic.config_spec(name))
    layer = ConstraintModule(cs, create_map=False).to(dev)
    db = B.DeviceBench(layer, batch, dev, pool=2)
    for i in range(4):
        db.step(i)
    torch.cuda.synchronize()
    db.backward(db.sets[0], 1)
    buf = (ctypes.c_longlong * 8192)()
    fn = _cabi.lib().rayen_bwd_trace_read
    fn.argtypes = [ctypes.c_void_p]
    fn(buf)
    h = np.array(buf[:], dtype=np.int64)
    st = h[:4096].reshape(256, 16)
    gt = h[4096:4096 + 512].reshape(256, 2)
    nb = min(256, (batch + 127) // 128)
    g0 = gt[:nb, 0].min()
    print(f"== {name} B={batch}: {nb} CTAs; CTA start offsets ns: min {int((gt[:nb,0]-g0).min())} median {int(np.median(gt[:nb,0]-g0))} max {int((gt[:nb,0]-g0).max())};"
          f" CTA end offsets ns: median {int(np.median(gt[:nb,1]-g0))} max {int((gt[:nb,1]-g0).max())}")
    names = ["start", "loaded", "normalised", "dk", "tail", "stored"]
    for cta in (0, 1, 100, nb - 1):
        t = st[cta, :6] - st[cta, 0]
        print(f"cta {cta:3d} warp 0 cycles: " + " ".join(f"{n}={int(x)}" for n, x in zip(names, t)), f"| CTA duration ns {int(gt[cta,1]-gt[cta,0])}")
    dur = gt[:nb, 1] - gt[:nb, 0]
    print("CTA duration ns: min", int(dur.min()), "median", int(np.median(dur)), "max", int(dur.max()), "argmax", int(dur.argmax()))
