#!/bin/bash
# ncu capture with source-level sampling of the per-pair-sum kernel on c3 (3 M points, same density)
tag=${1:-s}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_sum -s 2 -c 1 -f -o gpurun_out/${tag}_prof_sum_c3sd3M python bench.py --config c3 --npart 3000000 --same-density --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_ncu_c3.log 2>&1
tail -1 gpurun_out/${tag}_ncu_c3.log
