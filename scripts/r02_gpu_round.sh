#!/bin/bash
# One GPU-box pass of round 2: parity tests, bench line, ncu launch list of the same bench command, ncu --set full captures
# of the step's three kernels, compute-sanitizer over a subset of the parity tests.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
T=${1:-r02}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_all"], d["cuda_graph"]["ms_per_step"], d["module_autograd"]["ms_per_step"])
print(d["e2e"]["ms_per_step"], d["e2e"]["copy_floor_ms"], d["e2e"]["copies_only_ms"], d["e2e"]["pipelined"]["ms_per_step"])
for k,v in d["extra"].items(): print(k, round(v["ms_per_step"],4), round(v["fwd_ms"],4))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"lmi_forward_warp_kernel|lqs_tc_forward_kernel|lqs_backward_kernel" \
  -s 9 -c 3 -f -o gpurun_out/${T}_step_kernels python scripts/prof_lmi.py cfg5 32768 5 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
    -k "golden and (cfg5 or cfg4 or cfg3) or epigraph or forward_computed or wide_sets_match or host_buffer_path_matches" \
    > gpurun_out/${T}_sanitize_$tool.log 2>&1
  echo "$tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_sanitize_$tool.log | tail -3
done
