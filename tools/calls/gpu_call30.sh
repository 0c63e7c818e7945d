#!/bin/bash
O=gpurun_out
mkdir -p $O
bash tools/ab.sh tools/time_nms.py base mg24 mg28 mg32 > $O/r2m_ab.log 2>&1
bash tools/ab.sh tools/pipe_value.py base mg28 >> $O/r2m_ab.log 2>&1
