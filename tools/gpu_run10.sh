#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -k "not config4 and not config5 and not config3" 2>&1 | tail -4 > gpurun_out/pytest_coop.txt; cat gpurun_out/pytest_coop.txt
MXD_TUNE_COOP=0 python tools/batch_bench.py > gpurun_out/batch_bench.txt 2>&1
python tools/batch_bench.py >> gpurun_out/batch_bench.txt 2>&1
cat gpurun_out/batch_bench.txt
