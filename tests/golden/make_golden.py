"""Generate the committed golden vectors from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

For every canned geometry of the reference (examples/examples_sets.py:85-200 + the README set) and
for the BASELINE.json synthetic configs, the reference package under /root/reference is imported
as is (oracle/reference_loader.py stubs only the absent third-party solvers), a
``ConstraintModule(cs, method='RAYEN', create_map=False)`` is built from the spec with the spec's
explicit interior point, and its forward + autograd backward are recorded in float32 (the graded
dtype, reference default) and float64 (reference benchmark dtype, time_analysis.py:25).

Each npz holds: the spec (when small enough; cfg5 stores a sha256 of the regenerated spec instead),
the reference's preprocessed fields (A_p, b_p, NA_E, yp, z0, y0), inputs v / gy, and outputs
y32 / gv32 / y64 / gv64.  The GPU box has no /root/reference: tests only read these files.
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.reference_loader import load_reference  # noqa: E402
from rayen_b200 import synthetic  # noqa: E402

GOLDEN_BATCH = {"example": 96, "cfg1": 128, "cfg2": 128, "cfg3": 128, "cfg4": 64, "cfg5": 64}
FULL_SPEC_LIMIT_BYTES = 120_000


def spec_to_arrays(spec):
    out = {}
    for key in ("A1", "b1", "A2", "b2", "y0"):
        if spec[key] is not None:
            out["spec_" + key] = np.asarray(spec[key], dtype=np.float64)
    for i, (P, q, r) in enumerate(spec["qcs"]):
        out[f"spec_qc{i}_P"], out[f"spec_qc{i}_q"], out[f"spec_qc{i}_r"] = P, q, np.asarray(r, dtype=np.float64).reshape(1, 1)
    for i, (M, s, c, d) in enumerate(spec["socs"]):
        out[f"spec_soc{i}_M"], out[f"spec_soc{i}_s"], out[f"spec_soc{i}_c"], out[f"spec_soc{i}_d"] = M, s, c, d
    if spec["lmi"] is not None:
        out["spec_lmi"] = np.asarray(spec["lmi"], dtype=np.float64)
    return out


def arrays_to_spec(z):
    """Inverse of spec_to_arrays (used by the tests)."""
    spec = dict(A1=None, b1=None, A2=None, b2=None, qcs=[], socs=[], lmi=None, y0=None)
    for key in ("A1", "b1", "A2", "b2", "y0"):
        if "spec_" + key in z:
            spec[key] = np.array(z["spec_" + key])
    i = 0
    while f"spec_qc{i}_P" in z:
        spec["qcs"].append((np.array(z[f"spec_qc{i}_P"]), np.array(z[f"spec_qc{i}_q"]), np.array(z[f"spec_qc{i}_r"])))
        i += 1
    i = 0
    while f"spec_soc{i}_M" in z:
        spec["socs"].append(tuple(np.array(z[f"spec_soc{i}_{f}"]) for f in "Mscd"))
        i += 1
    if "spec_lmi" in z:
        spec["lmi"] = [np.array(F) for F in z["spec_lmi"]]
    return spec


def spec_digest(spec):
    h = hashlib.sha256()
    arrs = spec_to_arrays(spec)
    for key in sorted(arrs):
        h.update(key.encode())
        h.update(np.ascontiguousarray(arrs[key], dtype=np.float64).tobytes())
    return h.hexdigest()


def run_reference(ref, spec, v, gy, dtype):
    torch.set_default_dtype(dtype)  # the reference's buffers take the default dtype at ctor time
    try:
        cs = synthetic.build_constraints(spec, module=ref.constraints)
        layer = ref.constraint_module.ConstraintModule(cs, method="RAYEN", create_map=False)
        x = v.to(dtype).reshape(v.shape[0], -1, 1).clone().requires_grad_(True)
        y = layer(x)
        (y[:, :, 0] * gy.to(dtype)).sum().backward()
        return cs, y.detach()[:, :, 0].numpy(), x.grad[:, :, 0].numpy()
    finally:
        torch.set_default_dtype(torch.float32)


def make_one(ref, name, spec, batch, scale, store_spec=True):
    cs = synthetic.build_constraints(spec)
    v, gy = synthetic.sample_inputs(batch, cs.n, cs.k, dtype=torch.float64, scale=scale)
    v32, gy32 = v.float(), gy.float()
    # a few hand-placed edge inputs: zero vector, tiny vector (interior), huge vector
    v32[0] = 0.0
    v32[1] = v32[1] * 1e-3
    v32[2] = v32[2] * 1e3
    try:
        run_reference(ref, spec, v32, gy32, torch.float32)  # full batch: max(+0,-0) differs between SIMD and scalar paths
        zero_row_ok = True
    except AttributeError:
        # reference quirk: v == 0 with no active linear row gives kappa = -0.0 -> 1/kappa = -inf -> NaN,
        # and its NaN assert then dies on the missing args_DC3 (constraint_module.py:531)
        zero_row_ok = False
        v32[0] = v32[3] * 0.5
    v = v32.double()
    cs_ref, y32, gv32 = run_reference(ref, spec, v32, gy32, torch.float32)
    _, y64, gv64 = run_reference(ref, spec, v, gy32.double(), torch.float64)
    arrays = dict(v=v32.numpy(), gy=gy32.numpy(), y32=y32, gv32=gv32, y64=y64, gv64=gv64,
                  digest=np.array(spec_digest(spec)), zero_row_ok=np.array(zero_row_ok))
    spec_arrays = spec_to_arrays(spec)
    if store_spec and sum(a.nbytes for a in spec_arrays.values()) <= FULL_SPEC_LIMIT_BYTES:
        arrays.update(spec_arrays)
        for f in ("A_p", "b_p", "NA_E", "yp", "z0", "y0"):
            arrays["ref_" + f] = np.asarray(getattr(cs_ref, f), dtype=np.float64)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **arrays)
    print(f"{name:12s} k={cs.k} n={cs.n} B={batch} -> {os.path.getsize(path)/1024:.1f} KiB "
          f"(max|y32-y64|={np.abs(y32-y64).max():.2e})")


def make_old(ref, name, spec, batch, scale):
    """method='RAYEN_old' (reference constraint_module.py:460-466): inputs have n+1 columns, the last is beta."""
    cs = synthetic.build_constraints(spec)
    q, gy = synthetic.sample_inputs(batch, cs.n + 1, cs.k, dtype=torch.float64, scale=scale)
    q32, gy32 = q.float(), gy.float()
    out = {}
    for tag, dtype in (("32", torch.float32), ("64", torch.float64)):
        torch.set_default_dtype(dtype)
        try:
            cs_ref = synthetic.build_constraints(spec, module=ref.constraints)
            layer = ref.constraint_module.ConstraintModule(cs_ref, method="RAYEN_old", create_map=False)
            x = q32.to(dtype).reshape(batch, -1, 1).clone().requires_grad_(True)
            y = layer(x)
            (y[:, :, 0] * gy32.to(dtype)).sum().backward()
            out["y" + tag], out["gv" + tag] = y.detach()[:, :, 0].numpy(), x.grad[:, :, 0].numpy()
        finally:
            torch.set_default_dtype(torch.float32)
    arrays = dict(v=q32.numpy(), gy=gy32.numpy(), digest=np.array(spec_digest(spec)), zero_row_ok=np.array(True), **out)
    arrays.update(spec_to_arrays(spec))
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **arrays)
    print(f"{name:12s} (RAYEN_old) k={cs.k} n={cs.n} B={batch} -> {os.path.getsize(path)/1024:.1f} KiB")


def main():
    warnings.filterwarnings("ignore")
    ref = load_reference()
    for ex in synthetic.EXAMPLE_IDS:
        make_one(ref, f"example_{ex}", synthetic.example_spec(ex), GOLDEN_BATCH["example"], scale=5.0)
    for cfg in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
        make_one(ref, cfg, synthetic.config_spec(cfg), GOLDEN_BATCH[cfg], scale=2.0)
    for ex in ("readme", 13, 1):
        make_old(ref, f"old_example_{ex}", synthetic.example_spec(ex), 64, scale=3.0)
    spec = synthetic.random_spec(k=8, m=24, eta=3, mu=3, r_M=8, r=8, seed=3)
    spec["b1"] = spec["b1"] * 3.0
    make_old(ref, "old_mixed8", spec, 64, scale=2.0)
    # "balanced" variants: linear rows loosened so that every family is active for some samples
    for cfg in ("cfg2", "cfg3", "cfg5"):
        spec = synthetic.config_spec(cfg)
        spec["b1"] = spec["b1"] * 4.0
        make_one(ref, cfg + "_loose", spec, GOLDEN_BATCH[cfg], scale=2.0, store_spec=False)


def main_big():
    """LMIs beyond 32 x 32 and LMIs together with n > 32 (round 2): python tests/golden/make_golden.py big"""
    warnings.filterwarnings("ignore")
    ref = load_reference()
    for name in synthetic.BIG_LMI_GOLDEN:
        make_one(ref, name, synthetic.big_lmi_spec(name), 64, scale=2.0, store_spec=False)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        main_big()
    else:
        main()
