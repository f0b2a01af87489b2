#!/bin/bash
# compute-sanitizer over the wide-set GPU tests (run under gpurun on a B200 box); results under gpurun_out/.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "wide_sets_match or wide_host" -x \
    > gpurun_out/sanitize_wide_$tool.log 2>&1
  echo "$tool: exit $?" >> gpurun_out/sanitize_wide_summary.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_wide_$tool.log >> gpurun_out/sanitize_wide_summary.log
done
cat gpurun_out/sanitize_wide_summary.log
