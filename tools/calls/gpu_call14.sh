#!/bin/bash
# lockstep candidates kernel + rounds-based NMS resolve: tests, then A/B against the previous build
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -25 > $O/r2b_pytest.log
bash tools/ab.sh tools/time_assign.py base gw32 old > $O/r2b_ab.log 2>&1
bash tools/ab.sh tools/time_nms.py base walk >> $O/r2b_ab.log 2>&1
KZD_TN_MAXDET=30000 bash tools/ab.sh tools/time_nms.py base walk >> $O/r2b_ab.log 2>&1
bash tools/ab.sh tools/pipe_value.py base gw32 walk old >> $O/r2b_ab.log 2>&1
KZD_PV_G=1000 bash tools/ab.sh tools/pipe_value.py base gw32 old >> $O/r2b_ab.log 2>&1
