#!/bin/bash
# round-2 GPU run 20 (1 GPU): cooperative kernel with the producer 3 / 4 blocks ahead instead of 2 (A/B builds)
mkdir -p gpurun_out
O=gpurun_out/r2_coop_stages.txt
: > $O
export PAIR_BENCH_N=32,256,1000,4737,9472,16384,32768 MXD_TUNE_PAIR=0
for v in default coop3 coop4; do
  echo "== stages: $v" >> $O
  if [ $v = default ]; then timeout 300 python tools/pair_bench.py >> $O 2>&1; else MODELX_B200_LIB=build/variants/libmxd_$v.so timeout 300 python tools/pair_bench.py >> $O 2>&1; fi
done
echo done
