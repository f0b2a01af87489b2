"""Phase trace of lmi_forward_warp_kernel (development build with -DRAYEN_LW_TRACE).

Build:  cd rayen_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared \
        -Xcompiler -fPIC -DRAYEN_LW_TRACE -o ../../scripts/bin/librayen_b200_lwtrace.so rayen_b200.cu
Run:    RAYEN_B200_LIB=$PWD/scripts/bin/librayen_b200_lwtrace.so python scripts/lw_trace.py cfg5:32768 cfg5x4:32768
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as B
from rayen_b200 import synthetic, _cabi
from rayen_b200.constraint_module import ConstraintModule

CH = ["chunk_start", "prologue(u)", "contract4", "ldlt4", "pass_y", "chunk_end", "idx_loaded", "F_staged", "kernel_entry"]
SV = ["solve_start", "contract1", "tridiag", "sturm", "merge_y", "eigvec", "grad"]
dev = torch.device("cuda", 0)
for arg in sys.argv[1:]:
    name, batch = arg.split(":")
    batch = int(batch)
    loosen = 1.0
    if "x" in name:
        name, l = name.split("x")
        loosen = float(l)
    spec = synthetic.config_spec(name)
    spec["b1"] = spec["b1"] * loosen
    cs = synthetic.build_constraints(spec)
    layer = ConstraintModule(cs, create_map=False).to(dev)
    db = B.DeviceBench(layer, batch, dev, pool=2)
    for i in range(4):
        db.forward(db.sets[i % 2])
    torch.cuda.synchronize()
    db.forward(db.sets[0])
    buf = (ctypes.c_longlong * 8192)()
    fn = _cabi.lib().rayen_lw_trace_read
    fn.argtypes = [ctypes.c_void_p]
    fn(buf)
    h = np.array(buf[:], dtype=np.int64)
    per = h[:2048].reshape(128, 16)
    print(f"== {name} (rows x{loosen}) B={batch}: cycles relative to kernel entry, warps of the first CTAs (latest chunk)")
    order = [8, 6, 7, 0, 1, 2, 3, 4, 5]
    for w in (0, 1, 4, 8, 9, 64, 127):
        t0 = per[w, 8]
        print(f"warp {w:3d}: " + " ".join(f"{CH[i]}={int(per[w, i] - t0)}" for i in order if per[w, i] > 0))
    sv = h[4096:4096 + 7]
    if sv[0] > 0:
        print("latest full solve (cycles since its start): " + " ".join(f"{SV[i]}={int(sv[i] - sv[0])}" for i in range(7) if sv[i] >= sv[0]))
    gt = h[6144:6144 + 512].reshape(256, 2)[:148]
    live = gt[:, 1] > gt[:, 0]
    if live.any():
        g0 = gt[live, 0].min()
        end = np.sort(gt[live, 1] - g0)
        print(f"CTA end times (ns after the first CTA start), {int(live.sum())} CTAs: median {int(np.median(end))} p90 {int(end[int(0.9 * len(end))])} "
              f"last five {end[-5:].tolist()}; start spread {int((gt[live, 0] - g0).max())} ns")
    act = db.sets[0]["active"].cpu().numpy() >> 24
    print("binding families:", np.bincount(act, minlength=5).tolist())
    del db, layer
    torch.cuda.empty_cache()
