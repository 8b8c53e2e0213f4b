#!/bin/bash
# round 2, call c (2 or more GPUs): the drop-in callers again, one C call sharded over the devices inside the library,
# and the in-library multi-GPU bench line next to the one-process-per-GPU (torchrun) line
mkdir -p gpurun_out
N=${1:-2}
( timeout 900 python -m pytest tests -m gpu -x -q -k "dropin or many_gpus or reference_header" ) > gpurun_out/r2c_pytest.log 2>&1
tail -4 gpurun_out/r2c_pytest.log
timeout 900 python bench.py --inlib --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c5_inlib_g$N.json 2> gpurun_out/r2c_bench_c5_inlib_g$N.err
python tools/bench_summary.py inlib$N < gpurun_out/r2c_bench_c5_inlib_g$N.json
tail -3 gpurun_out/r2c_bench_c5_inlib_g$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2c_bench_c5_torchrun_g$N.json 2> gpurun_out/r2c_bench_c5_torchrun_g$N.err
python tools/bench_summary.py torchrun$N < gpurun_out/r2c_bench_c5_torchrun_g$N.json
