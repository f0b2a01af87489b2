"""Phase trace of lmi_forward_kernel (development build with -DRAYEN_LMI_TRACE).

Build:  cd rayen_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared \
        -Xcompiler -fPIC -DRAYEN_LMI_TRACE -o librayen_b200_trace.so rayen_b200.cu
Run:    RAYEN_B200_LIB=$PWD/rayen_b200/csrc/librayen_b200_trace.so python scripts/lmi_trace.py cfg5:32768 cfg4:4096
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as B
from rayen_b200 import synthetic, _cabi
from rayen_b200.constraint_module import ConstraintModule

NAMES = ["start", "staged_issue", "iter0", "dir", "F_wait", "contract", "tridiag", "sturm", "write", "eigvec", "grad", "end"]
dev = torch.device("cuda", 0)
for arg in sys.argv[1:]:
    name, batch = arg.split(":")
    batch = int(batch)
    cs = synthetic.build_constraints(synthetic.config_spec(name))
    layer = ConstraintModule(cs, create_map=False).to(dev)
    db = B.DeviceBench(layer, batch, dev, pool=2)
    for i in range(4):
        db.forward(db.sets[i % 2])
    torch.cuda.synchronize()
    db.forward(db.sets[0])
    buf = (ctypes.c_longlong * 4096)()
    fn = _cabi.lib().rayen_lmi_trace_read
    fn.argtypes = [ctypes.c_void_p]
    fn(buf)
    h = np.array(buf[:], dtype=np.int64).reshape(64, 64)[:16]
    if os.environ.get("RAYEN_LMI_TC", "1") != "0":
        NAMES[:] = ["start", "setup", "iter0", "dir", "Uready", "W0ok", "mma0_issued", "ahead_issued"] + [f"{w}{p}" for p in range(8) for w in ("m", "d")] + ["Aregs", "tridiag", "sturm", "end"]
    print(f"== {name} B={batch}: cycles since kernel start of the LAST loop iteration's phases (warp 0 | warp 7)")
    for cta in (0, 1, 7, 15):
        for w, off in (("w0", 0), ("w7", 32)):
            t = h[cta, off:off + len(NAMES)]
            t0 = h[cta, 0]
            print(f"cta{cta:2d} {w}: " + " ".join(f"{NAMES[i]}={int(t[i] - t0)}" for i in range(len(NAMES)) if NAMES[i] != "?"))
    gt = np.array(buf[2048:2048 + 512], dtype=np.int64).reshape(256, 2)[:148]
    live = gt[:, 1] > gt[:, 0]
    if live.any():
        g0 = gt[live, 0].min()
        end = np.sort(gt[live, 1] - g0)
        print(f"CTA end times (ns after the first CTA start), {int(live.sum())} CTAs: median {int(np.median(end))} p90 {int(end[int(0.9 * len(end))])} "
              f"last five {end[-5:].tolist()}; start spread {int((gt[live, 0] - g0).max())} ns")
    del db, layer
    torch.cuda.empty_cache()
