#!/bin/bash
# round 2, sum-kernel v3 (bank-aligned histogram copies, 2-word sums, float mu estimate) + merged wp pi-cut bodies:
# parity suite, then the configs they serve, then block-shape variants of the sum kernel
tag=${1:-h}
shift
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -6 gpurun_out/${tag}_pytest.log
for c in c3 c2 c2wp32 c2rppi c2rppi32 m1 m2; do
  timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/${tag}_bench_$c.json
done
for v in default "$@"; do
  if [ $v = default ]; then unset CORRFUNC_B200_LIBPATH; else export CORRFUNC_B200_LIBPATH=$PWD/corrfunc_b200/csrc/variants/libcorrfunc_b200_$v.so; fi
  echo "== $v"
  python tools/exp_sum.py 3e6 2>&1 | grep -v legacy | tail -7
done
