#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/r2_pytest3.log
python tools/time_assign.py v4 > gpurun_out/r2_time_assign.log 2>&1
python tools/time_assign.py v4u uniform >> gpurun_out/r2_time_assign.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:nms_sweep --launch-skip 2 -c 1 -f -o gpurun_out/r2_sweep python tools/nms_only.py > gpurun_out/r2_ncu_sweep.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:assign_patch --launch-skip 2 -c 1 -f -o gpurun_out/r2_patch python tools/assign_only.py > gpurun_out/r2_ncu_patch.log 2>&1
ls -la gpurun_out | tail -8
