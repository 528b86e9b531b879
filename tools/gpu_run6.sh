#!/bin/bash
# round-2 GPU run 6 (8 GPUs): the bench at N=8
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.txt 2> gpurun_out/r2_bench_n8.err
echo "bench rc=$?" >> gpurun_out/r2_bench_n8.err
echo done
