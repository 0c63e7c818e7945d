#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python tools/sanitize_target.py > $O/r2_sanitize_plain.log 2>&1 && timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > $O/r2_sanitize_memcheck.log 2>&1
echo "exit $?" >> $O/r2_sanitize_memcheck.log
tail -5 $O/r2_sanitize_plain.log; tail -15 $O/r2_sanitize_memcheck.log
