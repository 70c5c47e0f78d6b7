#!/bin/bash
# full GPU suite + smoke on the current build (what the driver runs at round end), with durations
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c21; mkdir -p $O
( time timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=15 ) > $O/gpu_tests.log 2>&1; echo "gpu tests rc=$?" > $O/summary.txt
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -25 $O/gpu_tests.log; tail -8 $O/smoke.log
