#!/bin/bash
# patch kernel: image barrier + distributed normalisation, strip index for contested anchors; candidates occupancy variants
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -25 > $O/r2d_pytest.log
KZD_LIB_PATH=$PWD/kuzushiji_vision_b200/variants/libkzd_trace.so python tools/trace_patch.py 600 > $O/r2d_trace.log 2>&1
KZD_LIB_PATH=$PWD/kuzushiji_vision_b200/variants/libkzd_trace.so python tools/trace_patch.py 1000 >> $O/r2d_trace.log 2>&1
bash tools/ab.sh tools/time_assign.py base m5 m6 gw32 gw32m6 old > $O/r2d_ab.log 2>&1
bash tools/ab.sh tools/pipe_value.py base m5 m6 gw32 gw32m6 old >> $O/r2d_ab.log 2>&1
KZD_PV_G=1000 bash tools/ab.sh tools/pipe_value.py base m5 m6 gw32 gw32m6 old >> $O/r2d_ab.log 2>&1
