#!/bin/bash
# round 2, call d: full GPU suite with the new tests; bench lines of every config with the new bench.py
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2d_pytest.log 2>&1
tail -6 gpurun_out/r2d_pytest.log
for c in c5 c5DD c1 c2 c2rppi c2rppi32 c2wp32 c3 c4 m1 m2; do
  timeout 900 python bench.py --config $c --steps 3 --no-cpu-baseline > gpurun_out/r2d_bench_$c.json 2> gpurun_out/r2d_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/r2d_bench_$c.json
done
