#!/bin/bash
# round-2 GPU call 1: full GPU test-suite, assign timings, bench line, source-level ncu of the candidates kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/r2_pytest1.log
python tools/time_assign.py v2 > gpurun_out/r2_time_assign.log 2>&1
python tools/time_assign.py v2u uniform >> gpurun_out/r2_time_assign.log 2>&1
python tools/pipe_value.py v2 4 > gpurun_out/r2_pipe_value.log 2>&1
KZD_PV_G=1000 python tools/pipe_value.py v2_G1000 4 >> gpurun_out/r2_pipe_value.log 2>&1
python tools/graph_kernel_times.py > gpurun_out/r2_graph_kernel_times_call1.md 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/r2_bench1_ref.json 2> gpurun_out/r2_bench1_ref.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:assign_ --launch-skip 4 -c 2 -f -o gpurun_out/r2_assign python tools/assign_only.py > gpurun_out/r2_ncu_assign.log 2>&1
ls -la gpurun_out
