#!/bin/bash
# round-2 GPU run 23 (4 GPUs): bench at N=4 with the final tree
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2_final_bench_n4.txt 2> gpurun_out/r2_final_bench_n4.err
echo "bench rc=$?" >> gpurun_out/r2_final_bench_n4.err
echo done
