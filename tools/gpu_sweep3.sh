#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q -k "fast or brute or duplicate or narrow" 2>&1 | tail -2
for occ in 0 84; do
  timeout 300 python bench.py --config c5 --npart 6000000 --same-density --occ $occ --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "occ=$occ"
done
timeout 300 python bench.py --config c1 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c1"
timeout 300 python bench.py --config c4 --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c4"
timeout 300 python bench.py --config c2 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 3 -c 1 -o gpurun_out/prof_fast_c5sd_v3 python bench.py --config c5 --npart 1500000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log
