#!/bin/bash
# round 2, last pass on one GPU (after the gridlink rewrite, the tail split of the fast kernel and the drain changes of the
# per-pair-sum kernel): full suite, the default bench line (driver-style steps), every other bench config, ncu launch
# lists + full captures of the fast kernel and of the per-pair-sum kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2w_pytest.log 2>&1
tail -4 gpurun_out/r2w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120
( time timeout 1500 python bench.py ) > gpurun_out/r2w_bench_c5.json 2> gpurun_out/r2w_bench_c5.err
python tools/bench_summary.py c5 < gpurun_out/r2w_bench_c5.json
for c in c5DD c5d c1 c2 c2wp32 c2rppi c2rppi32 c3 c4 m1 m2; do
  timeout 900 python bench.py --config $c --steps 3 --no-cpu-baseline > gpurun_out/r2w_bench_$c.json 2> gpurun_out/r2w_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/r2w_bench_$c.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches_c5sd3M.csv python bench.py --config c5 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 3 -c 1 -f -o gpurun_out/r2w_prof_fast_c5sd3M python bench.py --config c5 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_sum -s 3 -c 1 -f -o gpurun_out/r2w_prof_sum_c3sd3M python bench.py --config c3 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_(partition|place)" -c 2 -f -o gpurun_out/r2w_prof_gridlink_c5sd30M python bench.py --config c5 --npart 30000000 --same-density --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2w_ncu_grid.log 2>&1
ls -la gpurun_out/r2w_*.ncu-rep
