#!/bin/bash
# round-2 GPU run 4: staging vs zero-copy (pinned windows) feed of a page-cache resident file, incl. the CPU-starved case
mkdir -p gpurun_out
O=gpurun_out/r2_feed_bench.txt; rm -f $O
export FB_GB=24
MXD_HOST_FEED=stage timeout 300 python tools/feed_bench.py >> $O 2>&1
for th in 1 2 4; do MXD_HOST_FEED=map MXD_MAP_THREADS=$th timeout 300 python tools/feed_bench.py >> $O 2>&1; done
MXD_HOST_FEED=map MXD_MAP_THREADS=2 MXD_MAP_WINDOW=268435456 timeout 300 python tools/feed_bench.py >> $O 2>&1
MXD_HOST_FEED=map MXD_MAP_THREADS=3 MXD_MAP_WINDOW=536870912 MXD_MAP_DEPTH=4 timeout 300 python tools/feed_bench.py >> $O 2>&1
MXD_HOST_FEED=map MXD_MAP_THREADS=2 MXD_MAP_WINDOW=4294967296 timeout 300 python tools/feed_bench.py >> $O 2>&1
echo "--- 2 CPUs (taskset -c 0-1): what a rank gets when 8 ranks share a 16-CPU quota" >> $O
MXD_HOST_FEED=stage MXD_STAGE_THREADS=2 timeout 600 taskset -c 0-1 python tools/feed_bench.py >> $O 2>&1
MXD_HOST_FEED=map MXD_MAP_THREADS=1 timeout 300 taskset -c 0-1 python tools/feed_bench.py >> $O 2>&1
MXD_HOST_FEED=map MXD_MAP_THREADS=2 timeout 300 taskset -c 0-1 python tools/feed_bench.py >> $O 2>&1
rm -f /dev/shm/modelx_b200_feed.bin
echo "--- fresh-mapping probe" >> $O
PROBE_GB=8 timeout 300 python tools/register_probe.py >> $O 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "tree or tee or config1 or cli" > gpurun_out/r2_pytest_gpu_4.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_4.txt
echo done
