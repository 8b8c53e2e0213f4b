#!/bin/bash
# ncu capture with source-level sampling of the fast kernel on c5 (3 M points, same density)
tag=${1:-f}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 2 -c 1 -f -o gpurun_out/${tag}_prof_fast_c5sd3M python bench.py --config c5 --npart 3000000 --same-density --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_ncu_c5.log 2>&1
tail -1 gpurun_out/${tag}_ncu_c5.log
timeout 300 python -m pytest tests -m gpu -x -q -k "device_pointers or DD_DR_RR or many_gpus" 2>&1 | tail -2
