#!/bin/bash
# round-2 opening pass: full GPU suite, baselines of the per-pair-sum kernel on every config it serves (c3, c2rppi,
# c2rppi32, m1, m2 -- the mocks ones for the first time), one ncu full capture of it in a mocks mode.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_a.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
nproc; lscpu | grep -E "Model name|^CPU\(s\)"
for c in c3 c2rppi c2rppi32 m1 m2 c1 c2 c4; do
  timeout 600 python bench.py --config $c --steps 3 > gpurun_out/r2a_bench_$c.json 2> gpurun_out/r2a_bench_$c.err
  python tools/bench_summary.py $c < gpurun_out/r2a_bench_$c.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_generic -s 3 -c 1 -o gpurun_out/r2a_prof_generic_m1 python bench.py --config m1 --npart 1000000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_m1.log 2>&1
tail -2 gpurun_out/r2a_ncu_m1.log
