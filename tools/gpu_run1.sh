#!/bin/bash
# round-2 GPU run 1: new core on the real device -- parity tests, leaf-kernel A/B, batch bench, half-warp ubench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpuinfo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_1.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_1.txt
timeout 120 ./build/halfwarp > gpurun_out/r2_halfwarp.txt 2>&1
for v in default legacy; do
  if [ $v = legacy ]; then export MXD_TUNE_LEAF=legacy; else unset MXD_TUNE_LEAF; fi
  for sz in 12500000000 100000000000; do
    QB_SIZE=$sz QB_LEAVES=16384 timeout 300 python tools/quick_bench.py >> gpurun_out/r2_quick_bench_$v.txt 2>&1
  done
done
unset MXD_TUNE_LEAF
MXD_TUNE_LEAF_SCHED=1 QB_SIZE=12500000000 QB_LEAVES=16384 timeout 300 python tools/quick_bench.py > gpurun_out/r2_quick_bench_gridsched.txt 2>&1
MXD_TUNE_LEAF_SCHED=1 QB_SIZE=100000000000 QB_LEAVES=16384 timeout 300 python tools/quick_bench.py >> gpurun_out/r2_quick_bench_gridsched.txt 2>&1
timeout 300 python tools/batch_bench.py > gpurun_out/r2_batch_bench_1.txt 2>&1
echo done
