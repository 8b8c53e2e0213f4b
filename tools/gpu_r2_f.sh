#!/bin/bash
# round 2, call f: the reference's extension modules on the GPU; c2rppi lattice / kernel choices
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_dropin.py -m gpu -x -q ) > gpurun_out/r2f_pytest.log 2>&1
tail -4 gpurun_out/r2f_pytest.log
for c in c2rppi c2rppi32; do
  for v in "new56::" "new112:CORRFUNC_B200_SUM_OCC=112:" "new80:CORRFUNC_B200_SUM_OCC=80:" "legacy112:CORRFUNC_B200_SUM_OCC=112:CORRFUNC_B200_LEGACY_GENERIC=1" "legacy56::CORRFUNC_B200_LEGACY_GENERIC=1"; do
    name=${v%%:*}; rest=${v#*:}; e1=${rest%%:*}; e2=${rest#*:}
    env $e1 $e2 timeout 300 python bench.py --config $c --steps 3 --no-cpu-baseline 2>/dev/null | python tools/bench_summary.py "$c-$name"
  done
done
