#!/bin/bash
# quick regression + speed check of the fast kernel
timeout 300 python -m pytest tests -m gpu -x -q -k "fast or brute or duplicate or narrow" 2>&1 | tail -2
timeout 300 python bench.py --config c5 --npart 6000000 --same-density --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c5sd"
timeout 300 python bench.py --config c1 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c1"
timeout 300 python bench.py --config c4 --steps 1 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c4"
timeout 300 python bench.py --config c2 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c2"
