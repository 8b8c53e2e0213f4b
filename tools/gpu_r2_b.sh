#!/bin/bash
# round 2, call b: drop-in callers + new parity tests, ALU probes (and the ubench3 question under ncu), DRAM traffic of
# the 100 M-point launch, first bench line with the new bench.py
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "dropin or fast_acos or many_gpus or c5dsd10M or reference_header" ) > gpurun_out/r2b_pytest.log 2>&1
tail -5 gpurun_out/r2b_pytest.log
tools/bin/ubench > gpurun_out/r2b_ubench.txt 2>&1; grep -E "ffma x8|dfma x8|ffma2 x8" gpurun_out/r2b_ubench.txt
timeout 300 ncu --metrics smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max,smsp__issue_active.sum --clock-control none --csv --log-file gpurun_out/r2b_ubench3_ncu.csv tools/bin/ubench3 > gpurun_out/r2b_ubench3.txt 2>&1
tail -12 gpurun_out/r2b_ubench3.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_pairs_fast -s 1 -c 1 --csv --log-file gpurun_out/r2b_traffic_c5.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_traffic_c5.log 2>&1
tail -3 gpurun_out/r2b_traffic_c5.csv
