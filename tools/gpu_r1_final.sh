#!/bin/bash
# round-1 final record: everything the driver runs at round end (pytest -m gpu, smoke, default bench, reference arm),
# plus the other configs and an ncu capture of the DDtheta kernel
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/z_pytest.log 2>&1
tail -3 gpurun_out/z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-80
( time timeout 1500 python bench.py ) > gpurun_out/z_bench_c5_full.json 2> gpurun_out/z_bench_c5_full.err
python tools/bench_summary.py c5full < gpurun_out/z_bench_c5_full.json
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/z_bench_c5_reference.json 2> gpurun_out/z_bench_c5_reference.err
cut -c1-200 gpurun_out/z_bench_c5_reference.json
for c in c1 c2 c3 c4; do
  timeout 900 python bench.py --config $c --steps 3 > gpurun_out/z_bench_$c.json 2> gpurun_out/z_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/z_bench_$c.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 3 -c 1 -o gpurun_out/z_prof_fast_c4 python bench.py --config c4 --npart 1000000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/z_ncu_c4.log 2>&1
tail -1 gpurun_out/z_ncu_c4.log
