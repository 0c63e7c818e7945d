#!/bin/bash
# usage: bash tools/gpu_scale.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 400 gpurun_out/r2_bench_${N}gpu.err
