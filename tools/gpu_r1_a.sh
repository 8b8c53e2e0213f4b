#!/bin/bash
# round-1 GPU pass A: parity tests, bench (reduced + full c5), ncu launch list + full capture of the pair kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/a_gpu.txt; free -g >> gpurun_out/a_gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
tail -3 gpurun_out/a_pytest.log
timeout 600 python bench.py --config c5 --npart 10000000 --same-density --steps 2 --no-cpu-baseline > gpurun_out/a_bench_c5sd10M.json 2> gpurun_out/a_bench_c5sd10M.err
python tools/bench_summary.py c5sd10M < gpurun_out/a_bench_c5sd10M.json
( time timeout 1500 python bench.py ) > gpurun_out/a_bench_c5_full.json 2> gpurun_out/a_bench_c5_full.err
python tools/bench_summary.py c5full < gpurun_out/a_bench_c5_full.json
tail -3 gpurun_out/a_bench_c5_full.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches_c5sd3M.csv python bench.py --config c5 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/a_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 3 -c 1 -o gpurun_out/a_prof_fast_c5sd3M python bench.py --config c5 --npart 3000000 --same-density --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/a_ncu_full.log 2>&1
tail -2 gpurun_out/a_ncu_full.log
ls -la gpurun_out
