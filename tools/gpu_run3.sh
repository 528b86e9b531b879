#!/bin/bash
# round-2 GPU run 3: ncu evidence for the leaf kernel and the bench step, host-path probes, full bench, sanitizer
mkdir -p gpurun_out
QB_SIZE=4000000000 QB_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sha256_lanes -s 1 -c 1 -f -o gpurun_out/r2_prof_leaf python tools/leaf_bench.py > gpurun_out/r2_ncu_leaf.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-compat > gpurun_out/r2_bench_under_ncu.log 2>&1
PROBE_GB=8 timeout 600 python tools/register_probe.py > gpurun_out/r2_register_probe.txt 2>&1
TL_GB=8 timeout 300 python tools/slot_timeline.py > gpurun_out/r2_slot_timeline.txt 2>&1
timeout 300 python tools/leaf_bench.py > gpurun_out/r2_leaf_default_final.txt 2>&1
QB_SIZE=100000000000 timeout 300 python tools/leaf_bench.py >> gpurun_out/r2_leaf_default_final.txt 2>&1
timeout 1200 python bench.py > gpurun_out/r2_bench_n1_b.txt 2> gpurun_out/r2_bench_n1_b.err
echo "bench rc=$?" >> gpurun_out/r2_bench_n1_b.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_ref.txt 2> gpurun_out/r2_bench_ref.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_digest_service.py tests/test_client_flows.py -m gpu -x -q -k "ranges or per_file or part_digests or corrupted" > gpurun_out/r2_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.txt
echo done
