#!/bin/bash
# generic kernel (2-D statistics, weights, averages): parity + timing
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
for c in c3 c2rppi c2rppi32; do
  timeout 600 python bench.py --config $c --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py $c
done
