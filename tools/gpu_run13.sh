#!/bin/bash
# round-2 GPU run 13 (1 GPU): the two-lanes-per-chain kernel (k_sha256_chains_pair): speed A/B, then the whole GPU test-suite
mkdir -p gpurun_out
timeout 600 python tools/pair_bench.py > gpurun_out/r2_pair_bench.txt 2>&1
MXD_TUNE_PAIR=0 timeout 600 python tools/pair_bench.py >> gpurun_out/r2_pair_bench.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pair_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pair_pytest_gpu.txt
timeout 300 python tools/hasher_bench.py > gpurun_out/r2_pair_hasher.txt 2>&1
echo done
