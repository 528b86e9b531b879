#!/bin/bash
mkdir -p gpurun_out
for v in 4 5 6; do MXD_TUNE_MINB=$v QB_SIZE=20000000000 QB_LEAVES=16384 python tools/quick_bench.py; done > gpurun_out/quick_bench_v4.txt 2>&1
cat gpurun_out/quick_bench_v4.txt
