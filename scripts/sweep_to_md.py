"""gpurun_out/*time_analysis.json (scripts/time_analysis.py) -> the markdown table committed under profiles/.
usage: python scripts/sweep_to_md.py gpurun_out/r02_time_analysis.json profiles/r02_time_analysis.md "round 2" """
import json, sys
src, dst, label = sys.argv[1], sys.argv[2], sys.argv[3]
d = json.load(open(src))
pts = d["points"]
keys = {"linear": ("r_A1", "k"), "qp": ("eta", "k"), "soc": ("r_M", "mu", "k"), "lmi": ("r_F", "k")}
with open(dst, "w") as f:
    f.write(f"# The reference's timing sweep on B200 ({label})\n\n"
            "`python scripts/time_analysis.py` on one B200: the sweep of the reference's `examples/scripts/time_analysis.py:57-190` (one forward\n"
            "call on 2000 samples per point; same random generators) for the sizes this library covers.  `fwd` = `rayen_forward_f32` through the C ABI,\n"
            "CUDA events, mean of 10 after 3 warm-ups, device-resident inputs; `fwd+bwd` adds `rayen_backward_f32`.  `CPU port` = the oracle port of the\n"
            "reference's forward in float64 (as the reference script runs it) on the box's host threads, one pass over the same 2000 samples (only\n"
            "where that takes a few seconds).  `err` = max |y - y_oracle| / max |y_oracle| on 16 samples (float64 oracle), `viol` = max float32 residual of\n"
            "the original constraints over the 2000 outputs (GPU metric).  `kernels`: narrow = register-resident (n <= 32), wide = wide.cuh,\n"
            "big-LMI = lmi_big.cuh behind either.\n"
            f"Raw data: `{dst.replace('.md', '.json')}`.  {len(pts)} points.  Not covered: {d.get('not_covered')}.\n\n")
    for fam in ("linear", "qp", "soc", "lmi"):
        rows = [p for p in pts if p["family"] == fam]
        if not rows:
            continue
        ks = keys[fam]
        rows.sort(key=lambda p: tuple(p[k] for k in ks))
        f.write(f"\n## {fam}\n\n| " + " | ".join(ks) + " | kernels | fwd us | fwd s/sample | fwd+bwd us | CPU port s/sample | CPU/GPU | err | viol |\n"
                + "|---" * (len(ks) + 8) + "|\n")
        for p in rows:
            kern = ("wide" if p["wide"] else "narrow") + (" + big-LMI" if fam == "lmi" else "")
            cpu = p.get("cpu_port_s_per_sample")
            f.write("| " + " | ".join(str(p[k]) for k in ks) + f" | {kern} | {p['fwd_us']} | {p['fwd_s_per_sample']:.2e} | {p['fwd_bwd_us']} | "
                    + (f"{cpu:.2e} | {cpu / p['fwd_s_per_sample']:.0f}x" if cpu else "- | -")
                    + f" | {p['rel_err_y_vs_oracle']:.2e} | {p['max_violation']:.2e} |\n")
    f.write(f"\nWorst `err` {max(p['rel_err_y_vs_oracle'] for p in pts):.2e}, worst `viol` {max(p['max_violation'] for p in pts):.2e}.\n")
print(open(dst).read()[-600:])
