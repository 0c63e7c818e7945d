#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/time_nms.py bucket > gpurun_out/r2_time_nms.log 2>&1
KZD_TN_MAXDET=30000 timeout 120 python tools/time_nms.py bucket_all >> gpurun_out/r2_time_nms.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2_pytest9.log
bash tools/ab.sh tools/pipe_value.py base nobucket > gpurun_out/r2_ab_bucket.log 2>&1
python tools/graph_kernel_times.py > gpurun_out/r2_graph_kernel_times_call9.md 2>&1
