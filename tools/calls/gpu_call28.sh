#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "loss or criterion or every_kernel or reference" 2>&1 | tail -8 > $O/r2l_pytest.log
timeout 600 python tools/time_frow.py > $O/r2l_frow.log 2>&1
tail -3 $O/r2l_frow.log
