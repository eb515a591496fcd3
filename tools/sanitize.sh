#!/bin/bash
# compute-sanitizer pass (run under gpurun, one GPU):  bash tools/sanitize.sh  -> gpurun_out/sanitizer_*.log
set -u
mkdir -p gpurun_out
python tools/sanitizer_case.py > gpurun_out/sanitizer_plain.log 2>&1 || { tail -20 gpurun_out/sanitizer_plain.log; exit 1; }
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer case ok" gpurun_out/sanitizer_$tool.log
done
