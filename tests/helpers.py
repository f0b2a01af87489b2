"""Shared test helpers: golden-file loading (the files were produced by tests/golden/make_golden.py from
the unmodified reference; nothing here reads /root/reference)."""
import hashlib
import os

import numpy as np

from rayen_b200 import synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def _spec_from_arrays(z):
    spec = dict(A1=None, b1=None, A2=None, b2=None, qcs=[], socs=[], lmi=None, y0=None)
    for key in ("A1", "b1", "A2", "b2", "y0"):
        if "spec_" + key in z:
            spec[key] = np.array(z["spec_" + key])
    i = 0
    while f"spec_qc{i}_P" in z:
        spec["qcs"].append(tuple(np.array(z[f"spec_qc{i}_{f}"]) for f in "Pqr"))
        i += 1
    i = 0
    while f"spec_soc{i}_M" in z:
        spec["socs"].append(tuple(np.array(z[f"spec_soc{i}_{f}"]) for f in "Mscd"))
        i += 1
    if "spec_lmi" in z:
        spec["lmi"] = [np.array(F) for F in z["spec_lmi"]]
    return spec


def _spec_arrays(spec):
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


def spec_digest(spec):
    h = hashlib.sha256()
    arrs = _spec_arrays(spec)
    for key in sorted(arrs):
        h.update(key.encode())
        h.update(np.ascontiguousarray(arrs[key], dtype=np.float64).tobytes())
    return h.hexdigest()


def load_golden(name):
    """dict(spec, v, gy, y32, gv32, y64, gv64, zero_row_ok, ref fields if stored)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {k: np.array(z[k]) for k in ("v", "gy", "y32", "gv32", "y64", "gv64")}
    out["zero_row_ok"] = bool(z["zero_row_ok"])
    if "spec_y0" in z:
        spec = _spec_from_arrays(z)
    elif name.startswith("big_"):
        spec = synthetic.big_lmi_spec(name)
    else:  # large specs are regenerated from the seeded generator and checked by digest
        base = name.replace("_loose", "")
        spec = synthetic.config_spec(base)
        if name.endswith("_loose"):
            spec["b1"] = spec["b1"] * 4.0
    assert spec_digest(spec) == str(z["digest"]), f"golden {name}: constraint set does not match its digest"
    out["spec"] = spec
    out["ref"] = {k[4:]: np.array(z[k]) for k in z.files if k.startswith("ref_")}
    return out
