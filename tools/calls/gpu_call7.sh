#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh tools/time_assign.py base NOTAIL NOSCAN NOREP > gpurun_out/r2_ab_patch.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/r2_pytest7.log
timeout 600 python tools/time_frow.py > gpurun_out/r2_frow.log 2>&1
