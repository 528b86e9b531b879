#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.txt 2>&1; cat gpurun_out/pytest_gpu.txt
