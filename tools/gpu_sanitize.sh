#!/bin/bash
# compute-sanitizer over every kernel family (tools/sanitize_cases.py); logs -> gpurun_out/<tag>_compute_sanitizer_*.log
TAG=${1:-r03}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 0 python tools/sanitize_cases.py > gpurun_out/${TAG}_compute_sanitizer_${tool}.log 2>&1
  tail -3 gpurun_out/${TAG}_compute_sanitizer_${tool}.log
done
