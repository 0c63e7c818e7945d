#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 > $O/r2f_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2f_bench.json 2> $O/r2f_bench.err
python tools/time_nms.py flag > $O/r2f_time_nms.log 2>&1
python tools/graph_kernel_times.py > $O/r2f_graph_kernel_times.md 2>&1
timeout 600 python tools/time_frow.py > $O/r2f_frow.log 2>&1
KZD_BENCH_FUSE=0 KZD_BENCH_FILTER=0 timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2f_bench_unfused.json 2> $O/r2f_bench_unfused.err
