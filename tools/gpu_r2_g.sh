#!/bin/bash
# round 2, call g: half-warp split of the fast kernel -- parity (fast-kernel tests + full-size goldens) and speed with / without
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "fast or brute or duplicate or narrow or full_size or window or sharding" ) > gpurun_out/r2g_pytest.log 2>&1
tail -3 gpurun_out/r2g_pytest.log
for v in default nosplit; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  timeout 300 python bench.py --config c5 --npart 10000000 --same-density --steps 3 --no-cpu-baseline 2>/dev/null | python tools/bench_summary.py "$v c5sd10M"
  timeout 300 python bench.py --config c5 --npart 10000000 --same-density --steps 3 --no-cpu-baseline 2>/dev/null | python tools/bench_summary.py "$v c5sd10M"
done
unset CORRFUNC_B200_LIBPATH
timeout 600 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/r2g_bench_c5.json 2>/dev/null; python tools/bench_summary.py c5full < gpurun_out/r2g_bench_c5.json
for c in c2rppi c2rppi32; do timeout 300 python bench.py --config $c --steps 3 --no-cpu-baseline 2>/dev/null | python tools/bench_summary.py $c; done
