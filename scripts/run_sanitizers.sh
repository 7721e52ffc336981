#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (the reference has no sanitizer coverage at all).
# Writes gpurun_out/<prefix>_sanitizer_{memcheck,racecheck}.log with the command line on top (copied to profiles/).
#   bash scripts/run_sanitizers.sh r02
set -u
prefix=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL_MEM='maxpool or stem or head or split_precision or split_maxpool or small_pertap-fp16] or window_matches or negative_values or warp or compose_matches or resize_and or map_attributes or femoral or tibial or smoothing or face_features'
SEL_RACE='maxpool or stem or head_matches or split_maxpool or map_attributes[1] or tibial or warp_volume_and'
for tool in memcheck racecheck; do
  sel="$SEL_MEM"; [ "$tool" = racecheck ] && sel="$SEL_RACE"
  log=gpurun_out/${prefix}_sanitizer_${tool}.log
  cmd="compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -q -m gpu -x -k \"$sel\""
  echo "# $cmd" > "$log"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "$sel" 2>&1 | grep -v "^$" | tail -12 >> "$log"
  echo "exit ${PIPESTATUS[0]}" >> "$log"
done
tail -n 4 gpurun_out/${prefix}_sanitizer_memcheck.log; tail -n 4 gpurun_out/${prefix}_sanitizer_racecheck.log
