#!/bin/bash
# first GPU run: parity + quick kernel timing
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
python tools/quick_bench.py > gpurun_out/quick_bench.txt 2>&1
cat gpurun_out/quick_bench.txt
