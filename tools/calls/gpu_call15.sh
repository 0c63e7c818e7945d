#!/bin/bash
O=gpurun_out
mkdir -p $O
KZD_LIB_PATH=$PWD/kuzushiji_vision_b200/variants/libkzd_trace.so python tools/trace_patch.py 600 > $O/r2c_trace.log 2>&1
KZD_LIB_PATH=$PWD/kuzushiji_vision_b200/variants/libkzd_trace.so python tools/trace_patch.py 1000 >> $O/r2c_trace.log 2>&1
python tools/assign_only.py > $O/r2c_assign_only.log 2>&1 && timeout 600 ncu --set full --import-source on --clock-control none -k regex:assign_ --launch-skip 4 -c 2 -f -o $O/r2c_assign python tools/assign_only.py > $O/r2c_ncu.log 2>&1
KZD_LIB_PATH=$PWD/kuzushiji_vision_b200/variants/libkzd_gw32.so python tools/assign_only.py >> $O/r2c_assign_only.log 2>&1 && KZD_LIB_PATH=$PWD/kuzushiji_vision_b200/variants/libkzd_gw32.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:assign_cand --launch-skip 2 -c 1 -f -o $O/r2c_assign32 python tools/assign_only.py > $O/r2c_ncu32.log 2>&1
