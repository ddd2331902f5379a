#!/bin/bash
# compute-sanitizer over tools/gpu_sanitize_target.py (run under gpurun): memcheck (with leak check), racecheck (shared-memory
# hazards: the tile hand-off flags and staging areas of tile_kernel, the gather ring of the row-block walks), synccheck.
OUT=gpurun_out/sanitize
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check full"
  timeout 1200 compute-sanitizer --tool $tool $extra --print-limit 20 python tools/gpu_sanitize_target.py all > $OUT/$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' $OUT/$tool.log) summary lines: $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY\|LEAK SUMMARY' $OUT/$tool.log | tr '\n' ' ')"
done
