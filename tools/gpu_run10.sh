#!/bin/bash
# round-2 GPU run 10 (8 GPUs): bench at N=8 with NUMA-local page-cache placement; the in-process 8-GPU path
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8_b.txt 2> gpurun_out/r2_bench_n8_b.err
echo "bench rc=$?" >> gpurun_out/r2_bench_n8_b.err
nvidia-smi topo -m > gpurun_out/r2_topo8.txt 2>&1
cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2_topo8.txt 2>&1
echo done
