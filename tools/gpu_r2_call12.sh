#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c12; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tracking.py -x -q -m gpu > $O/test_tracking.log 2>&1; echo "tracking rc=$?" >> $O/summary.txt
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_baseline_configs.py --deselect tests/test_gpu_tracking.py > $O/test_rest.log 2>&1; echo "rest rc=$?" >> $O/summary.txt
tail -15 $O/test_tracking.log; tail -3 $O/test_rest.log; cat $O/summary.txt
