#!/bin/bash
# occupancy / staging sweep on the c5 workload at reduced N but the same number density
for stage in cpasync tma; do
for occ in 0 60 84 112; do
  CORRFUNC_B200_STAGE=$stage timeout 300 python bench.py --config c5 --npart 6000000 --same-density --occ $occ --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "stage=$stage occ=$occ"
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 3 -c 1 -o gpurun_out/prof_fast_c5sd python bench.py --config c5 --npart 1500000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log
timeout 300 python -m pytest tests -m gpu -x -q -k "fast or brute or duplicate or narrow" 2>&1 | tail -3
