#!/bin/bash
# compute-sanitizer over the smoke batch (golden frames + one fresh 1 MiB frame) at several execute kernels.
# usage: tools/sanitize.sh memcheck|racecheck [FZG_EXEC_W values...]   -- ONE tool per gpurun call (B200_PROFILING.md)
tool=$1; shift
out=gpurun_out/sanitize_$tool.log; : > $out
python __graft_entry__.py smoke > gpurun_out/smoke_plain.log 2>&1 || { echo "plain smoke run failed"; tail -5 gpurun_out/smoke_plain.log; exit 1; }
for w in "$@"; do
  echo "===== compute-sanitizer --tool $tool, FZG_EXEC_W=$w" >> $out
  FZG_EXEC_W=$w timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke >> $out 2>&1
  echo "exit code $?" >> $out
done
grep -E "=====|ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|exit code|Error|hazard" $out | head -60
