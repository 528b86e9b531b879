#!/bin/bash
# round-2 GPU run 14 (1 GPU): which warp bounds a cooperative chain kernel?  producer-idle / chain-idle probes + ncu source view
mkdir -p gpurun_out
export PAIR_BENCH_NOCHECK=1 PAIR_BENCH_N=32,1000
O=gpurun_out/r2_chain_bottleneck.txt
: > $O
for v in CHAIN_IDLE PRODUCER_IDLE; do
  echo "== $v (pair kernel)" >> $O;  MODELX_B200_LIB=build/variants/libmxd_$v.so timeout 300 python tools/pair_bench.py >> $O 2>&1
  echo "== $v (coop kernel)" >> $O;  MXD_TUNE_PAIR=0 MODELX_B200_LIB=build/variants/libmxd_$v.so timeout 300 python tools/pair_bench.py >> $O 2>&1
done
echo "== full kernels" >> $O
timeout 300 python tools/pair_bench.py >> $O 2>&1
MXD_TUNE_PAIR=0 timeout 300 python tools/pair_bench.py >> $O 2>&1
PAIR_BENCH_N=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sha256_chains_pair -s 1 -c 1 -o gpurun_out/prof_pair python tools/pair_bench.py > gpurun_out/r2_ncu_pair.log 2>&1
PAIR_BENCH_N=32 MXD_TUNE_PAIR=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sha256_chains_coop -s 1 -c 1 -o gpurun_out/prof_coop2 python tools/pair_bench.py > gpurun_out/r2_ncu_coop2.log 2>&1
echo done
