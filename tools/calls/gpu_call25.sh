#!/bin/bash
O=gpurun_out
mkdir -p $O
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2_smoke.log 2>&1
( time python bench.py > $O/r2_bench_default.json 2> $O/r2_bench_default.err ) 2> $O/r2_bench_default.time
( time python bench.py --impl reference > $O/r2_bench_ref_default.json 2> $O/r2_bench_ref_default.err ) 2> $O/r2_bench_ref_default.time
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "every_kernel" 2>&1 | tail -3 > $O/r2_pytest_cov.log
tail -3 $O/r2_smoke.log; cat $O/r2_bench_default.time $O/r2_bench_ref_default.time; cat $O/r2_pytest_cov.log
