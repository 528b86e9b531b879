#!/bin/bash
# round-2 GPU run 26 (1 GPU): the host-flow tests on the CUDA build after the last host-side fixes (index size sum, untgz bounds)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_client_flows.py tests/test_cli.py tests/test_abi_cpu.py -m gpu -x -q > gpurun_out/r2_last_hostflows_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_last_hostflows_gpu.txt
echo done
