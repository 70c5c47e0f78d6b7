#!/bin/bash
# final tree: the long closed loops of tests/test_gpu_baseline_configs.py (the rest of the -m gpu suite ran in r2c52)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c53; mkdir -p $O
( time timeout 400 python -m pytest tests/test_gpu_baseline_configs.py -q -m gpu --durations=3 ) > $O/test_gpu_baseline.log 2>&1; echo "gpu tests rc=$?"
tail -8 $O/test_gpu_baseline.log
