#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cancel" 2>&1 | tail -4 > gpurun_out/pytest_cancel.txt; cat gpurun_out/pytest_cancel.txt
( time python bench.py ) > gpurun_out/bench_n1_v3.txt 2>&1; grep '^{' gpurun_out/bench_n1_v3.txt | python3 -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','e2e','e2e_file','cpu_baseline','clocks')})
print(d['roofline'])"
tail -4 gpurun_out/bench_n1_v3.txt | grep real
