#!/bin/bash
# round-2 GPU run 15 (1 GPU): pair kernel v3 (two blocks per producer pass, crossed value consumed late): speed, then parity
mkdir -p gpurun_out
O=gpurun_out/r2_pair_v3_bench.txt
: > $O
export PAIR_BENCH_N=1,32,1000,2368,4736,9472
for late in 0 1; do
  echo "== MXD_TUNE_PAIR=9472 MXD_TUNE_PAIR_LATE=$late" >> $O
  MXD_TUNE_PAIR=9472 MXD_TUNE_PAIR_LATE=$late timeout 300 python tools/pair_bench.py >> $O 2>&1
done
echo "== coop (default)" >> $O
timeout 300 python tools/pair_bench.py >> $O 2>&1
MXD_TUNE_PAIR=4736 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pair_v3_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pair_v3_pytest_gpu.txt
echo done
