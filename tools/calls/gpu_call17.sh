#!/bin/bash
# assign: separate normalise kernel, strip index for contested anchors in the patch kernel, score prefetch
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -25 > $O/r2e_pytest.log
bash tools/ab.sh tools/time_assign.py base nostrip gw32 old > $O/r2e_ab.log 2>&1
KZD_TA_G=1000 bash tools/ab.sh tools/time_assign.py base nostrip gw32 old >> $O/r2e_ab.log 2>&1
bash tools/ab.sh tools/pipe_value.py base nostrip gw32 old >> $O/r2e_ab.log 2>&1
KZD_PV_G=1000 bash tools/ab.sh tools/pipe_value.py base nostrip gw32 old >> $O/r2e_ab.log 2>&1
