#!/bin/bash
# round-2 GPU run 5 (2 GPUs): in-process multi-GPU tests, bench at N=2, zero-copy feed diagnostics
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus_run5.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu or shard or service or config1" > gpurun_out/r2_pytest_gpu_5.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_5.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.txt 2> gpurun_out/r2_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2_bench_n2.err
O=gpurun_out/r2_feed_bench_b.txt; rm -f $O
export FB_GB=16
MXD_HOST_FEED=stage timeout 300 python tools/feed_bench.py >> $O 2>&1
MXD_DEBUG_TIMING=1 MXD_HOST_FEED=map MXD_MAP_THREADS=2 timeout 300 python tools/feed_bench.py 2>&1 | grep -v "fill slot\|alloc\|stream_leaves\|sync compute\|tree finish\|chunk digests" | head -40 >> $O
rm -f /dev/shm/modelx_b200_feed.bin
echo done
