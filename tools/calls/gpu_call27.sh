#!/bin/bash
# candidates kernel: positive-overlap compaction before the metric, direct selection of short lists
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -12 > $O/r2k_pytest.log
bash tools/ab.sh tools/time_assign.py base prev > $O/r2k_ab.log 2>&1
KZD_TA_G=1000 bash tools/ab.sh tools/time_assign.py base prev >> $O/r2k_ab.log 2>&1
bash tools/ab.sh tools/pipe_value.py base prev >> $O/r2k_ab.log 2>&1
KZD_PV_G=1000 bash tools/ab.sh tools/pipe_value.py base prev >> $O/r2k_ab.log 2>&1
