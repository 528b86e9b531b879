#!/bin/bash
# round-2 GPU run 22 (2 GPUs): the in-process multi-GPU test (skipped on 1-GPU boxes) and bench at N=2 with the final tree
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu or shard" > gpurun_out/r2_final_pytest_gpu_n2.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_final_pytest_gpu_n2.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_final_bench_n2.txt 2> gpurun_out/r2_final_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2_final_bench_n2.err
echo done
