#!/bin/bash
# refresh of the secondary evidence files on the final tree (the primary capture is tools/gpu_call23.sh)
O=gpurun_out
mkdir -p $O
KZD_BENCH_FUSE=0 KZD_BENCH_FILTER=0 timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2_bench_unfused.json 2> $O/r2_bench_unfused.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 2 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_ref.err
python tools/graph_kernel_times.py > $O/r2_graph_kernel_times.md 2>&1
python tools/pipe_value.py G600 4 > $O/r2_pipe_value.log 2>&1
KZD_PV_G=1000 python tools/pipe_value.py G1000 4 >> $O/r2_pipe_value.log 2>&1
KZD_SKIP=assign python tools/pipe_value.py G600_noassign 4 >> $O/r2_pipe_value.log 2>&1
KZD_SKIP=nms python tools/pipe_value.py G600_nonms 4 >> $O/r2_pipe_value.log 2>&1
KZD_SKIP=assign,nms python tools/pipe_value.py G600_decodeonly 4 >> $O/r2_pipe_value.log 2>&1
for c in 1 2 3 4 5; do
  timeout 300 python tools/stage_times.py --config $c > $O/r2_stage_$c.log 2>&1
done
ls $O | grep -c .
