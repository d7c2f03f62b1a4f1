#!/bin/bash
# compute-sanitizer passes over a small workload that launches every kernel of the library (tools/sanitize_case.py):
#   gpurun --timeout 1500 -- 'bash tools/gpu_sanitize.sh r02'
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards; synccheck: invalid barrier use;
# initcheck: reads of device memory nobody wrote. Logs end up in gpurun_out/<tag>_sanitize_<tool>.log.
tag=${1:-rXX}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ $tool = racecheck ] && extra="--racecheck-report all"
  RK_N=${RK_N:-30000} timeout 300 compute-sanitizer --tool $tool $extra --print-limit 30 --target-processes all \
      python tools/sanitize_case.py > gpurun_out/${tag}_sanitize_${tool}.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case .* ok|launches" gpurun_out/${tag}_sanitize_${tool}.log | tail -9
done
