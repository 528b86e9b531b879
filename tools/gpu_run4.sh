#!/bin/bash
mkdir -p gpurun_out
MXD_TUNE_SMEM=10 python -m pytest tests -x -q -m gpu -k "tree or config2" 2>&1 | tail -5 > gpurun_out/pytest_smem.txt; cat gpurun_out/pytest_smem.txt
for v in 0 10 12; do echo "== MXD_TUNE_SMEM=$v"; MXD_TUNE_SMEM=$v QB_SIZE=20000000000 QB_LEAVES=4096,16384,65536 python tools/quick_bench.py; done > gpurun_out/quick_bench_smem.txt 2>&1
cat gpurun_out/quick_bench_smem.txt
