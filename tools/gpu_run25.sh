#!/bin/bash
# round-2 GPU run 25 (1 GPU): last confirmation of the final tree -- smoke + the GPU test-suite
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" ) > gpurun_out/r2_last_smoke.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_last_pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_last_pytest_gpu.txt
echo done
