#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
python tools/quick_bench.py > gpurun_out/quick_bench2.txt 2>&1; cat gpurun_out/quick_bench2.txt
( time python bench.py ) > gpurun_out/bench_n1.txt 2>&1; tail -8 gpurun_out/bench_n1.txt
( time python bench.py --impl reference ) > gpurun_out/bench_ref.txt 2>&1; tail -6 gpurun_out/bench_ref.txt
