#!/bin/bash
mkdir -p gpurun_out
# big-pool path (chunk sort + rank + mask + scan): time_nms.py uses pool 32768; skip the 3 warm-up calls
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nms_ --launch-skip 12 -c 4 -f -o gpurun_out/r2_nms python tools/time_nms.py ncu > gpurun_out/r2_ncu_nms.log 2>&1
tail -5 gpurun_out/r2_ncu_nms.log
