#!/bin/bash
# sum kernel: parity suite, the configs it serves, variants, then one ncu --set full capture of the c3-like 3 M run
tag=${1:-j}
shift
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -6 gpurun_out/${tag}_pytest.log
for c in c3 c2 c2wp32 m1 m2; do
  timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/${tag}_bench_$c.json
done
for v in default "$@"; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  echo "== $v"
  python tools/exp_sum.py 3e6 2>&1 | grep -v legacy | tail -7
done
unset CORRFUNC_B200_LIBPATH
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_sum -s 3 -c 1 -f -o gpurun_out/${tag}_prof_sum_c3sd3M python bench.py --config c3 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_c3.log 2>&1
ls -la gpurun_out/${tag}_prof_sum_c3sd3M.ncu-rep
