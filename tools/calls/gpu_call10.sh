#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2_pytest10.log
python tools/time_assign.py v7 > gpurun_out/r2_time_assign.log 2>&1
python tools/time_nms.py v7 > gpurun_out/r2_time_nms.log 2>&1
KZD_TN_MAXDET=30000 python tools/time_nms.py v7_all >> gpurun_out/r2_time_nms.log 2>&1
python tools/pipe_value.py v7 4 > gpurun_out/r2_pipe_value.log 2>&1
KZD_PV_G=1000 python tools/pipe_value.py v7_G1000 4 >> gpurun_out/r2_pipe_value.log 2>&1
