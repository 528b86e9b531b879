#!/bin/bash
# round-2 GPU run 19 (1 GPU): compute-sanitizer racecheck / synccheck / memcheck on the pair kernel (named barriers, 4 shared-memory stages)
mkdir -p gpurun_out
for t in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $t --error-exitcode 9 python tools/race_target.py > gpurun_out/r2_sanitizer_pair_$t.txt 2>&1
  echo "$t rc=$?" >> gpurun_out/r2_sanitizer_pair_$t.txt
done
echo done
