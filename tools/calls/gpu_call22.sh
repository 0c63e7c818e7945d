#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "assign or preprocess or gt_pack or pipeline or reference or loss" 2>&1 | tail -6 > $O/r2i_pytest.log
bash tools/ab.sh tools/time_assign.py base s48 s160 > $O/r2i_ab.log 2>&1
KZD_TA_G=1000 bash tools/ab.sh tools/time_assign.py base s48 s160 >> $O/r2i_ab.log 2>&1
KZD_PV_G=1000 bash tools/ab.sh tools/pipe_value.py base s48 >> $O/r2i_ab.log 2>&1
