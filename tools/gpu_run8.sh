#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_n1.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
python tools/cpu_scaling.py > gpurun_out/cpu_scaling.txt 2>&1; cat gpurun_out/cpu_scaling.txt
( time python bench.py ) > gpurun_out/bench_n1_v2.txt 2>&1; tail -5 gpurun_out/bench_n1_v2.txt
