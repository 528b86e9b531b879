#!/bin/bash
# round-2 GPU run 18 (1 GPU): pair kernel as the default for small launches -- test-suite, speed table, ncu, bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_v6_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_v6_pytest_gpu.txt
( time timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/r2_v6_smoke.txt 2>&1
timeout 300 python tools/pair_bench.py > gpurun_out/r2_v6_pair_bench.txt 2>&1
MXD_TUNE_PAIR=0 timeout 300 python tools/pair_bench.py >> gpurun_out/r2_v6_pair_bench.txt 2>&1
timeout 300 python tools/hasher_bench.py > gpurun_out/r2_v6_hasher.txt 2>&1
timeout 600 python tools/batch_bench.py > gpurun_out/r2_v6_batch_bench.txt 2>&1
PAIR_BENCH_N=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sha256_chains_pair -s 1 -c 1 -o gpurun_out/prof_pair_final python tools/pair_bench.py > gpurun_out/r2_v6_ncu_pair.log 2>&1
( time timeout 1200 python bench.py > gpurun_out/r2_v6_bench_n1.txt 2> gpurun_out/r2_v6_bench_n1.err ) > gpurun_out/r2_v6_bench_time.txt 2>&1
( time timeout 600 python bench.py --impl reference > gpurun_out/r2_v6_bench_ref.txt 2> gpurun_out/r2_v6_bench_ref.err ) >> gpurun_out/r2_v6_bench_time.txt 2>&1
echo done
