#!/bin/bash
mkdir -p gpurun_out
for v in 4 5 6 7 8; do MXD_TUNE_MINB=$v QB_SIZE=20000000000 QB_LEAVES=16384 python tools/quick_bench.py; done > gpurun_out/quick_bench_v3.txt 2>&1
MXD_TUNE_MINB=6 QB_SIZE=20000000000 QB_LEAVES=4096,65536 python tools/quick_bench.py >> gpurun_out/quick_bench_v3.txt 2>&1
cat gpurun_out/quick_bench_v3.txt
QB_SIZE=4000000000 QB_LEAVES=16384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sha256_lanes -s 2 -c 1 -o gpurun_out/prof_leaf_v2 python tools/quick_bench.py > gpurun_out/ncu_v2.log 2>&1
tail -3 gpurun_out/ncu_v2.log
