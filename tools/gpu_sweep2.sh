#!/bin/bash
for occ in 0 60 84; do
  timeout 300 python bench.py --config c5 --npart 6000000 --same-density --occ $occ --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "occ=$occ"
done
CORRFUNC_B200_STAGE=tma timeout 300 python bench.py --config c5 --npart 6000000 --same-density --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "tma occ=0"
timeout 300 python bench.py --config c1 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c1"
timeout 300 python bench.py --config c4 --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c4"
timeout 300 python bench.py --config c2 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 3 -c 1 -o gpurun_out/prof_fast_c5sd_v2 python bench.py --config c5 --npart 1500000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log
