#!/bin/bash
# round-2 GPU run 16 (1 GPU): pair kernel, last addition of the round as IADD3 (same-pipe hops) vs IMAD
mkdir -p gpurun_out
O=gpurun_out/r2_pair_v4_bench.txt
: > $O
export PAIR_BENCH_N=1,32,1000,2368,4736
for a in 1 0; do
  echo "== MXD_TUNE_PAIR=4736 MXD_TUNE_PAIR_ALUADD=$a" >> $O
  MXD_TUNE_PAIR=4736 MXD_TUNE_PAIR_ALUADD=$a timeout 300 python tools/pair_bench.py >> $O 2>&1
done
echo done
