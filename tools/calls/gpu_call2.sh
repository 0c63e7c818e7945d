#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/time_nms.py sweep > gpurun_out/r2_time_nms.log 2>&1
KZD_TN_MAXDET=30000 timeout 120 python tools/time_nms.py sweep_all >> gpurun_out/r2_time_nms.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 120 -k "nms or pipeline or chained" 2>&1 | tail -40 > gpurun_out/r2_pytest2_nms.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/r2_pytest2.log
python tools/time_assign.py v3 > gpurun_out/r2_time_assign.log 2>&1
python tools/time_assign.py v3u uniform >> gpurun_out/r2_time_assign.log 2>&1
python tools/pipe_value.py v3 4 > gpurun_out/r2_pipe_value.log 2>&1
KZD_PV_G=1000 python tools/pipe_value.py v3_G1000 4 >> gpurun_out/r2_pipe_value.log 2>&1
python tools/graph_kernel_times.py > gpurun_out/r2_graph_kernel_times_call2.md 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
ls -la gpurun_out | tail -12
