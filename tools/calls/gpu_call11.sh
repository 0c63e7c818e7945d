#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2_pytest11.log
bash tools/ab.sh tools/time_assign.py base gw32 > gpurun_out/r2_ab_gw.log 2>&1
bash tools/ab.sh tools/pipe_value.py base gw32 >> gpurun_out/r2_ab_gw.log 2>&1
KZD_PV_G=1000 bash tools/ab.sh tools/pipe_value.py base gw32 >> gpurun_out/r2_ab_gw.log 2>&1
