#!/bin/bash
for v in default base; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  for occ in 0 86; do
    timeout 300 python bench.py --config c5 --npart 6000000 --same-density --occ $occ --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "$v occ=$occ"
  done
done
unset CORRFUNC_B200_LIBPATH
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 3 -c 1 -o gpurun_out/prof_fast_c5sd_v4 python bench.py --config c5 --npart 1500000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
tail -2 gpurun_out/ncu_run.log
