#!/bin/bash
# round-2 GPU run 8 (1 GPU): ring-size sweep of the file -> digest path (is a cache-resident staging ring faster?), copy probe
mkdir -p gpurun_out
O=gpurun_out/r2_ring_sweep.txt; rm -f $O
export FB_GB=24
for rb in 4194304 16777216 33554432 67108864 134217728 268435456 1073741824; do
  MXD_RING_BYTES=$rb timeout 300 python tools/feed_bench.py >> $O 2>&1
done
echo "--- 4 CPUs (taskset)" >> $O
for rb in 16777216 268435456; do
  MXD_STAGE_THREADS=4 MXD_RING_BYTES=$rb timeout 300 taskset -c 0-3 python tools/feed_bench.py >> $O 2>&1
done
rm -f /dev/shm/modelx_b200_feed.bin
echo "--- copy probe (CPU only)" >> $O
timeout 300 build/copy_probe 16 16 >> $O 2>&1
timeout 300 build/copy_probe 16 4 >> $O 2>&1
echo done
