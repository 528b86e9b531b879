#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
QB_SIZE=4000000000 QB_LEAVES=16384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sha256_lanes -s 2 -c 1 -o gpurun_out/prof_leaf_v1 python tools/quick_bench.py > gpurun_out/ncu_v1.log 2>&1
tail -5 gpurun_out/ncu_v1.log
