#!/bin/bash
# compute-sanitizer over every QP path again, with the kernels added late in round 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c41; mkdir -p $O
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_probe.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/summary.txt
timeout 700 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_probe.py > $O/synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/summary.txt
cat $O/summary.txt; grep "ERROR SUMMARY" $O/memcheck.log $O/synccheck.log; grep -c "status" $O/memcheck.log; grep -B2 -A12 "Invalid\|Barrier error" $O/memcheck.log $O/synccheck.log | head -60
