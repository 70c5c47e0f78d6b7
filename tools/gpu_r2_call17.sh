#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c17; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_probe.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/summary.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_probe.py > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/summary.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_probe.py > $O/synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/summary.txt
cat $O/summary.txt; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|error" $O/*.log | head -20; tail -3 $O/memcheck.log
