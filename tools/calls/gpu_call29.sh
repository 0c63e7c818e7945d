#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > $O/r2_pytest.log
cat $O/r2_pytest.log
