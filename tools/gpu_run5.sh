#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
for v in 0; do echo "== MXD_TUNE_SMEM=$v"; MXD_TUNE_SMEM=$v QB_SIZE=20000000000 QB_LEAVES=4096,16384,65536 python tools/quick_bench.py; done > gpurun_out/quick_bench_v2.txt 2>&1
echo "== MINB=6" >> gpurun_out/quick_bench_v2.txt; MXD_TUNE_MINB=6 QB_SIZE=20000000000 QB_LEAVES=16384 python tools/quick_bench.py >> gpurun_out/quick_bench_v2.txt 2>&1
cat gpurun_out/quick_bench_v2.txt
