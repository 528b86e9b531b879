#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
python -m pytest tests -x -q -m gpu -k "not config3 and not config4 and not config5" 2>&1 | tail -5 > gpurun_out/pytest_gpu_fast.txt; cat gpurun_out/pytest_gpu_fast.txt
