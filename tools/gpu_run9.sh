#!/bin/bash
# round-2 GPU run 9 (1 GPU): staging copy A/B (pread vs mmap+memcpy vs mmap+streaming stores), tests, bench
mkdir -p gpurun_out
O=gpurun_out/r2_stage_copy_ab.txt; rm -f $O
export FB_GB=24
for rep in 1 2; do
MXD_STAGE_MMAP=0 timeout 300 python tools/feed_bench.py >> $O 2>&1
MXD_STAGE_NO_NT=1 timeout 300 python tools/feed_bench.py >> $O 2>&1
timeout 300 python tools/feed_bench.py >> $O 2>&1
done
MXD_STAGE_PIECE=4194304 timeout 300 python tools/feed_bench.py >> $O 2>&1
MXD_STAGE_PIECE=262144 timeout 300 python tools/feed_bench.py >> $O 2>&1
echo "--- 4 CPUs (taskset)" >> $O
MXD_STAGE_MMAP=0 MXD_STAGE_THREADS=4 timeout 300 taskset -c 0-3 python tools/feed_bench.py >> $O 2>&1
MXD_STAGE_THREADS=4 timeout 300 taskset -c 0-3 python tools/feed_bench.py >> $O 2>&1
echo "--- 2 CPUs (taskset)" >> $O
MXD_STAGE_MMAP=0 MXD_STAGE_THREADS=2 timeout 400 taskset -c 0-1 python tools/feed_bench.py >> $O 2>&1
MXD_STAGE_THREADS=2 timeout 400 taskset -c 0-1 python tools/feed_bench.py >> $O 2>&1
rm -f /dev/shm/modelx_b200_feed.bin
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_9.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_9.txt
timeout 1200 python bench.py > gpurun_out/r2_bench_n1_d.txt 2> gpurun_out/r2_bench_n1_d.err
echo "bench rc=$?" >> gpurun_out/r2_bench_n1_d.err
echo done
