#!/bin/bash
# round-2 GPU run 24 (1 GPU): mapped staging with MADV_POPULATE_READ per piece vs page-faulting mapped copy vs pread (default)
mkdir -p gpurun_out
O=gpurun_out/r2_stage_populate_ab.txt; rm -f $O
export FB_GB=24
for rep in 1 2; do
timeout 300 python tools/feed_bench.py >> $O 2>&1
MXD_STAGE_MMAP=1 timeout 300 python tools/feed_bench.py >> $O 2>&1
MXD_STAGE_MMAP=1 MXD_STAGE_NO_POPULATE=1 timeout 300 python tools/feed_bench.py >> $O 2>&1
done
MXD_STAGE_MMAP=1 MXD_STAGE_PIECE=4194304 timeout 300 python tools/feed_bench.py >> $O 2>&1
MXD_STAGE_MMAP=1 MXD_STAGE_NO_NT=1 timeout 300 python tools/feed_bench.py >> $O 2>&1
rm -f /dev/shm/modelx_b200_feed.bin
echo done
