#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over tools/sanitize_driver.py on one B200.
# Logs -> gpurun_out/sanitize_<tool>.log; tools/make_profiles.py copies the summaries into profiles/.
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  echo "== $tool"
  timeout 900 $CS --tool $tool --print-limit 30 --error-exitcode 9 python tools/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" gpurun_out/sanitize_$tool.log | sort | uniq -c | sort -rn | head -8
done
