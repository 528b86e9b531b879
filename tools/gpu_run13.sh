#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "in_process_multi" 2>&1 | tail -5 > gpurun_out/pytest_multi.txt; cat gpurun_out/pytest_multi.txt
python tools/h2d_bw.py > gpurun_out/h2d_bw.txt 2>&1; cat gpurun_out/h2d_bw.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench_n2_v2.txt 2>&1; grep '^{' gpurun_out/bench_n2_v2.txt | cut -c1-1200
