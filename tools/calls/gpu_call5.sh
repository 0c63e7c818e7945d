#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "conditional" 2>&1 | tail -15 > gpurun_out/r2_pytest5.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:assign_patch --launch-skip 2 -c 1 -f -o gpurun_out/r2_patch python tools/assign_only.py > gpurun_out/r2_ncu_patch.log 2>&1
