#!/bin/bash
# round-1 GPU pass J: record run after the generic-kernel rewrite and the typed entry points
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/j_pytest.log 2>&1
tail -3 gpurun_out/j_pytest.log
( time timeout 1500 python bench.py ) > gpurun_out/j_bench_c5_full.json 2> gpurun_out/j_bench_c5_full.err
python tools/bench_summary.py c5full < gpurun_out/j_bench_c5_full.json
for c in c1 c2 c2wp32 c2rppi c2rppi32 c3 c4; do
  timeout 900 python bench.py --config $c --steps 3 > gpurun_out/j_bench_$c.json 2> gpurun_out/j_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/j_bench_$c.json
done
timeout 900 python bench.py --config c5d --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j_bench_c5d.json 2> gpurun_out/j_bench_c5d.err
python tools/bench_summary.py c5d < gpurun_out/j_bench_c5d.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j_launches_c5sd3M.csv python bench.py --config c5 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 3 -c 1 -o gpurun_out/j_prof_fast_c5sd3M python bench.py --config c5 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j_ncu_full.log 2>&1
tail -2 gpurun_out/j_ncu_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
