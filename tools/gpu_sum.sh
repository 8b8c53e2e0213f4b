#!/bin/bash
# per-pair-sum kernel: parity suite (stop at first failure), then the configs it serves.  Extra env (e.g.
# CORRFUNC_B200_SUM_OCC=40) is taken from the caller.   gpurun --timeout 1200 -- 'bash tools/gpu_sum.sh [tag]'
tag=${1:-s}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -15 gpurun_out/${tag}_pytest.log
for c in c3 c2rppi c2rppi32 m1 m2; do
  timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/${tag}_bench_$c.json
done
