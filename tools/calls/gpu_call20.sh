#!/bin/bash
O=gpurun_out
mkdir -p $O
KZD_TD_B=4 KZD_TD_NC=4300 bash tools/ab.sh tools/time_decode.py base cg16 cg16f cg8f > $O/r2h_ab.log 2>&1
bash tools/ab.sh tools/time_decode.py base cg16 cg16f cg8f >> $O/r2h_ab.log 2>&1
