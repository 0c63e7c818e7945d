#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/r2_pytest6.log
python tools/time_assign.py v6 > gpurun_out/r2_time_assign.log 2>&1
python tools/time_assign.py v6u uniform >> gpurun_out/r2_time_assign.log 2>&1
python tools/pipe_value.py v6 4 > gpurun_out/r2_pipe_value.log 2>&1
KZD_PV_G=1000 python tools/pipe_value.py v6_G1000 4 >> gpurun_out/r2_pipe_value.log 2>&1
python tools/graph_kernel_times.py > gpurun_out/r2_graph_kernel_times_call6.md 2>&1
