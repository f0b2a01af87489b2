#!/bin/bash
# Final GPU-box pass of round 2: parity tests, bench line, ncu launch list of the same bench command, ncu --set full of the
# step's kernels and of the big-LMI kernels, compute-sanitizer memcheck over the kernels added this round.
mkdir -p gpurun_out
T=${1:-r02f}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${T}_gpu_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["direct_launch"]["ms_per_step"], d["roofline"]["kernel_ms_all"], d["module_autograd"]["ms_per_step"])
print(d["e2e"]["ms_per_step"], d["e2e"]["copy_floor_ms"], d["e2e"]["copies_only_ms"], d["e2e"]["pipelined"]["ms_per_step"], d["roofline"]["traffic"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"lmi_forward_warp_kernel|lqs_tc_forward_kernel|lqs_backward_kernel" \
  -s 9 -c 3 -f -o gpurun_out/${T}_step_kernels python scripts/prof_lmi.py cfg5 32768 5 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"lmib_" -s 2 -c 2 -f -o gpurun_out/${T}_lmi_big \
  python scripts/prof_lmi_big.py 100 100 2000 > gpurun_out/${T}_ncu_lmib.log 2>&1; echo "ncu lmib rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
  -k "big_lmi_sets and (r33 or r64 or wide_n40 or r40_old) or big_lmi_chunked or (feasible_to_1e_5 and (10000 or 5000)) or cuda_graph or golden and (big or cfg5)" \
  > gpurun_out/${T}_sanitize_memcheck.log 2>&1
echo "memcheck: exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${T}_sanitize_memcheck.log | tail -3
