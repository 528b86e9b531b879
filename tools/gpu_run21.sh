#!/bin/bash
# round-2 GPU run 21 (1 GPU): final validation of the committed tree -- smoke, test-suite, bench (timed), reference arm
mkdir -p gpurun_out
( time timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/r2_final3_smoke.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final3_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_final3_pytest_gpu.txt
( time timeout 1200 python bench.py > gpurun_out/r2_final3_bench_n1.txt 2> gpurun_out/r2_final3_bench_n1.err ) > gpurun_out/r2_final3_bench_time.txt 2>&1
( time timeout 600 python bench.py --impl reference > gpurun_out/r2_final3_bench_ref.txt 2> gpurun_out/r2_final3_bench_ref.err ) >> gpurun_out/r2_final3_bench_time.txt 2>&1
echo done
