#!/bin/bash
# round-2 evidence call: GPU tests, bench (both arms), in-graph kernel times, stage times of the five configs,
# f-row timings, ncu launch list of bench.py and one `--set full` capture of every kernel of the path
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r2_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > $O/r2_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2_bench.json 2> $O/r2_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 2 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_ref.err
python tools/graph_kernel_times.py > $O/r2_graph_kernel_times.md 2>&1
python tools/pipe_value.py G600 4 > $O/r2_pipe_value.log 2>&1
KZD_PV_G=1000 python tools/pipe_value.py G1000 4 >> $O/r2_pipe_value.log 2>&1
KZD_SKIP=assign python tools/pipe_value.py G600_noassign 4 >> $O/r2_pipe_value.log 2>&1
KZD_SKIP=nms python tools/pipe_value.py G600_nonms 4 >> $O/r2_pipe_value.log 2>&1
KZD_SKIP=assign,nms python tools/pipe_value.py G600_decodeonly 4 >> $O/r2_pipe_value.log 2>&1
for c in 1 2 3 4 5; do
  timeout 300 python tools/stage_times.py --config $c > $O/r2_stage_$c.log 2>&1 && cp $O/stage_times.json $O/r2_stage_times_cfg$c.json
done
timeout 600 python tools/time_frow.py > $O/r2_frow.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kzd|assign_|decode_|nms_' -c 600 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 > $O/r2_ncu_launches.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'assign_|decode_|nms_' --launch-skip 0 -f -o $O/r2_full python tools/ncu_target.py 1 > $O/r2_ncu_full.log 2>&1
ls -la $O | tail -30
