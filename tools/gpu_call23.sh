#!/bin/bash
# final capture after the last kernel change: tests, ncu --set full (traffic file regenerated on the box so that the
# bench line that follows carries roofline.traffic of the same kernel sources), launch list, bench
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 > $O/r2_pytest.log
python tools/ncu_target.py 1 > $O/r2_ncu_plain.log 2>&1 && timeout 900 ncu --set full --import-source on --clock-control none -k regex:'assign_|decode_|nms_' --launch-skip 0 -f -o $O/r2_full python tools/ncu_target.py 1 > $O/r2_ncu_full.log 2>&1
python tools/ncu_traffic.py $O/r2_full.ncu-rep > profiles/r2_traffic.json 2> $O/r2_traffic.err
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2_bench.json 2> $O/r2_bench.err
python tools/time_assign.py G600_lone_call > $O/r2_time_stage.log 2>&1
KZD_TA_G=1000 python tools/time_assign.py G1000_lone_call >> $O/r2_time_stage.log 2>&1
KZD_TA_OVERLAPPED=1 python tools/time_assign.py G600_as_in_pipeline >> $O/r2_time_stage.log 2>&1
KZD_TA_OVERLAPPED=1 KZD_TA_G=1000 python tools/time_assign.py G1000_as_in_pipeline >> $O/r2_time_stage.log 2>&1
python tools/time_nms.py maxdet1000 >> $O/r2_time_stage.log 2>&1
python tools/time_decode.py nc1 >> $O/r2_time_stage.log 2>&1
KZD_TD_B=4 KZD_TD_NC=4300 python tools/time_decode.py nc4300_b4 >> $O/r2_time_stage.log 2>&1
timeout 600 python tools/time_frow.py > $O/r2_frow.log 2>&1
timeout 600 python tools/fuzz_parity.py 1500 300 2026 200 > $O/r2_fuzz_final.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kzd|assign_|decode_|nms_' -c 700 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 > $O/r2_ncu_launches.log 2>&1
ls -la $O | tail -20
