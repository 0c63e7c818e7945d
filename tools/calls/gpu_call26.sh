#!/bin/bash
# second NMS phase by device-side tail launch
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -12 > $O/r2j_pytest.log
bash tools/ab.sh tools/time_nms.py base notail > $O/r2j_ab.log 2>&1
bash tools/ab.sh tools/pipe_value.py base notail >> $O/r2j_ab.log 2>&1
KZD_PV_G=1000 bash tools/ab.sh tools/pipe_value.py base notail >> $O/r2j_ab.log 2>&1
