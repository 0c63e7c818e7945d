#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh tools/pipe_value.py gw16p2 gw16p3 gw16p4 gw32p2 gw32p4 > gpurun_out/r2_ab_gw3.log 2>&1
KZD_PV_G=1000 bash tools/ab.sh tools/pipe_value.py gw16p2 gw16p3 gw16p4 gw32p2 gw32p4 >> gpurun_out/r2_ab_gw3.log 2>&1
