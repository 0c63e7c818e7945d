#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 > $O/r2g_pytest.log
timeout 600 python tools/time_frow.py > $O/r2g_frow.log 2>&1
