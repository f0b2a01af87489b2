"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python scripts/summarize_profiles.py launches gpurun_out/launches.csv profiles/r01_launches.md "<command that was profiled>"
    python scripts/summarize_profiles.py kernel   gpurun_out/prof.ncu-rep profiles/r01_<kernel>.md
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sass__inst_executed_local_loads",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, bi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Block Size"), hdr.index("Grid Size")
    agg = {}
    for r in rows[1:]:
        try:
            t = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[bi], r[gi]])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\ncommand: `{cmd}`\n\n"
                "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
                "| kernel | launches | total us | mean us | share | block | grid |\n|---|---|---|---|---|---|---|\n")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{n}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / 1e3 / a[0]:.1f} | {100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} |\n")
    print(open(dst).read())


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary: {src}\n\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            f.write(f"## `{name[:140]}`\n\n| metric | unit | value |\n|---|---|---|\n")
            for h, u, v in zip(hdr, units, vals):
                if h in KEYS:
                    f.write(f"| {h} | {u} | {v} |\n")
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        kernel(sys.argv[2], sys.argv[3])
