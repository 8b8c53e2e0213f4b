#!/bin/bash
# full-size config 5 with different target cell occupancies (device lattice = reference lattice x sub)
for occ in 92 100 112 66; do
  timeout 600 python bench.py --occ $occ --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "occ=$occ"
done
