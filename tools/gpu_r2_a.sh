#!/bin/bash
# round-2 opening pass (prepared at the end of round 1, when the GPU minutes were spent): the complete GPU suite incl. the
# float full-size twins and the mocks statistics that have not had a full-suite run yet, every bench config incl. the new
# mocks configs m1 / m2, and ncu launch lists + full captures of the fast kernel (c5 density) and of the generic kernel
# in its mocks modes.   gpurun --timeout 2400 -- 'bash tools/gpu_r2_a.sh'
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-100
( time timeout 1500 python bench.py ) > gpurun_out/r2a_bench_c5_full.json 2> gpurun_out/r2a_bench_c5_full.err
python tools/bench_summary.py c5full < gpurun_out/r2a_bench_c5_full.json
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/r2a_bench_c5_reference.json 2> gpurun_out/r2a_bench_c5_reference.err
cut -c1-200 gpurun_out/r2a_bench_c5_reference.json
for c in c1 c2 c2wp32 c2rppi c2rppi32 c3 c4 m1 m2; do
  timeout 900 python bench.py --config $c --steps 3 > gpurun_out/r2a_bench_$c.json 2> gpurun_out/r2a_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/r2a_bench_$c.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_c5sd3M.csv python bench.py --config c5 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 3 -c 1 -o gpurun_out/r2a_prof_fast_c5sd3M python bench.py --config c5 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_full.log 2>&1
tail -2 gpurun_out/r2a_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_generic -s 3 -c 1 -o gpurun_out/r2a_prof_generic_m1 python bench.py --config m1 --npart 1000000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_m1.log 2>&1
tail -2 gpurun_out/r2a_ncu_m1.log
