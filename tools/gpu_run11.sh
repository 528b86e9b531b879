#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "known_answers or ragged or empty_batch or all_alignments or verify_batch or hasher or file_digests or tree_digest_host or default_params or small_ring or sharded or generator" > gpurun_out/sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.txt; tail -12 gpurun_out/sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "known_answers or ragged or verify_batch or hasher" > gpurun_out/sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.txt; tail -8 gpurun_out/sanitizer_racecheck.txt
