#!/bin/bash
# round-2 GPU run 2: new test-suite on the device, leaf-kernel scheduling A/B with placement debug, new bench.py
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_2.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_2.txt
rm -f gpurun_out/r2_leaf_debug.txt gpurun_out/r2_leaf_variants.txt
for sz in 12500000000 100000000000; do
  export QB_SIZE=$sz
  timeout 200 python tools/leaf_bench.py >> gpurun_out/r2_leaf_variants.txt 2>&1
  MXD_TUNE_LEAF=legacy timeout 200 python tools/leaf_bench.py >> gpurun_out/r2_leaf_variants.txt 2>&1
  MXD_TUNE_LEAF_SCHED=1 timeout 200 python tools/leaf_bench.py >> gpurun_out/r2_leaf_variants.txt 2>&1
  MXD_TUNE_LEAF_KPER=11 timeout 200 python tools/leaf_bench.py >> gpurun_out/r2_leaf_variants.txt 2>&1
  MXD_TUNE_LEAF_KPER=10 timeout 200 python tools/leaf_bench.py >> gpurun_out/r2_leaf_variants.txt 2>&1
  MXD_TUNE_LEAF_KPER=8 timeout 200 python tools/leaf_bench.py >> gpurun_out/r2_leaf_variants.txt 2>&1
  MXD_TUNE_FUSE=1 timeout 200 python tools/leaf_bench.py >> gpurun_out/r2_leaf_variants.txt 2>&1
done
QB_SIZE=12500000000 QB_REPS=1 MXD_LEAF_DEBUG=gpurun_out/r2_leaf_debug.txt timeout 200 python tools/leaf_bench.py >> gpurun_out/r2_leaf_variants.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_a.txt 2> gpurun_out/r2_bench_n1_a.err
echo "bench rc=$?" >> gpurun_out/r2_bench_n1_a.err
echo done
