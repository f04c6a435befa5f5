#!/bin/bash
# Round-2 ncu evidence for the CTA-pair network kernel (run under gpurun, one GPU). Every profiled command first exits 0 without ncu.
set -u
OUT=gpurun_out
N="python tools/net_check.py 16384 3 f32"
$N > $OUT/r2b_plain_net_f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_net_forward -s 1 -c 1 -o $OUT/r2b_net_f32_pair -f $N > $OUT/r2b_ncu_net_f32.log 2>&1
N="python tools/net_check.py 16384 3 f16"
$N > $OUT/r2b_plain_net_f16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_net_forward -s 1 -c 1 -o $OUT/r2b_net_f16_quad -f $N > $OUT/r2b_ncu_net_f16.log 2>&1
S="python bench.py --workload selfplay --steps 1 --warmup 3 --no-cpu-baseline"
$S > $OUT/r2b_plain_selfplay.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 2410 --csv --log-file $OUT/r2b_launches_selfplay.csv $S > $OUT/r2b_ncu_selfplay.log 2>&1
python bench.py --steps 20 --warmup 5 > $OUT/r2b_bench_default.json 2> $OUT/r2b_bench_default.err
python bench.py --workload selfplay --steps 3 --warmup 3 > $OUT/r2b_bench_selfplay.json 2> $OUT/r2b_bench_selfplay.err
ls -la $OUT/ | tail -12
