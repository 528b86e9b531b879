#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/l2_vs_dram.txt
for sz in 100000000 20000000000; do
  echo "== size $sz" >> gpurun_out/l2_vs_dram.txt
  QB_SIZE=$sz QB_LEAVES=1024,2048 python tools/quick_bench.py 2>&1 | grep -v "full tree" >> gpurun_out/l2_vs_dram.txt
done
cat gpurun_out/l2_vs_dram.txt
