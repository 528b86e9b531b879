#!/bin/bash
# round-2 GPU run 7 (1 GPU): chain-round variants, 32-thread CTAs, hasher rate, final test-suite and bench
mkdir -p gpurun_out
O=gpurun_out/r2_chain_variants.txt; rm -f $O
for v in 0 1 2; do MXD_TUNE_CHAIN=$v timeout 300 python tools/batch_bench.py 2>&1 | sed "s/^/[MXD_TUNE_CHAIN=$v] /" >> $O; done
O=gpurun_out/r2_cta32.txt; rm -f $O
for sz in 12500000000 100000000000; do
  QB_SIZE=$sz timeout 200 python tools/leaf_bench.py >> $O 2>&1
  QB_SIZE=$sz MXD_TUNE_CTA=32 timeout 200 python tools/leaf_bench.py >> $O 2>&1
done
timeout 300 python tools/hasher_bench.py > gpurun_out/r2_hasher_bench.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_7.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_7.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1
timeout 1200 python bench.py > gpurun_out/r2_bench_n1_c.txt 2> gpurun_out/r2_bench_n1_c.err
echo "bench rc=$?" >> gpurun_out/r2_bench_n1_c.err
echo done
