#!/bin/bash
# round-1 GPU pass D: default library (persistent warps) regression + 2-GPU strong scaling of config 5
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --config c1 --steps 3 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c1"
timeout 300 python bench.py --config c4 --steps 2 --no-cpu-baseline 2>&1 | tail -1 | python tools/bench_summary.py "c4"
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 2 --warmup 3 ) > gpurun_out/d_bench_c5_g$NG.json 2> gpurun_out/d_bench_c5_g$NG.err
python tools/bench_summary.py "c5 x$NG" < gpurun_out/d_bench_c5_g$NG.json
tail -5 gpurun_out/d_bench_c5_g$NG.err
