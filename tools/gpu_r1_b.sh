#!/bin/bash
# round-1 GPU pass B: the other BASELINE configs (c1-c4), each with the reference CPU arm beside it
mkdir -p gpurun_out
for c in c1 c2 c3 c4; do
  ( time timeout 900 python bench.py --config $c --steps 3 ) > gpurun_out/b_bench_$c.json 2> gpurun_out/b_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/b_bench_$c.json
  tail -4 gpurun_out/b_bench_$c.err
done
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/b_bench_c5_reference.json 2> gpurun_out/b_bench_c5_reference.err
cat gpurun_out/b_bench_c5_reference.json
