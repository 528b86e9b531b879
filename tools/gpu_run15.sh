#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/quick_bench_variants.txt
for v in default rv1 cta256 cta64 addalu; do
  if [ $v = default ]; then unset MODELX_B200_LIB; else export MODELX_B200_LIB=$PWD/build/libmxd_$v.so; fi
  echo "== variant $v" >> gpurun_out/quick_bench_variants.txt
  QB_SIZE=20000000000 QB_LEAVES=16384 python tools/quick_bench.py >> gpurun_out/quick_bench_variants.txt 2>&1
done
cat gpurun_out/quick_bench_variants.txt
