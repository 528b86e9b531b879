#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/bench_n8.txt 2>&1; tail -4 gpurun_out/bench_n8.txt | cut -c1-1800
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 ) > gpurun_out/bench_n4.txt 2>&1; tail -4 gpurun_out/bench_n4.txt | cut -c1-1500
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 ) > gpurun_out/bench_ref_n8.txt 2>&1; tail -4 gpurun_out/bench_ref_n8.txt | cut -c1-600
