#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/time_nms.py cond > gpurun_out/r2_time_nms.log 2>&1
KZD_TN_MAXDET=30000 timeout 120 python tools/time_nms.py cond_all >> gpurun_out/r2_time_nms.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25 > gpurun_out/r2_pytest4.log
python tools/time_assign.py v5 > gpurun_out/r2_time_assign.log 2>&1
python tools/time_assign.py v5u uniform >> gpurun_out/r2_time_assign.log 2>&1
python tools/pipe_value.py v5 4 > gpurun_out/r2_pipe_value.log 2>&1
KZD_PV_G=1000 python tools/pipe_value.py v5_G1000 4 >> gpurun_out/r2_pipe_value.log 2>&1
python tools/graph_kernel_times.py > gpurun_out/r2_graph_kernel_times_call4.md 2>&1
ls -la gpurun_out | tail -6
