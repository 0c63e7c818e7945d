#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 > gpurun_out/r2_pytest8.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench8_ref.json 2> gpurun_out/r2_bench8_ref.err
