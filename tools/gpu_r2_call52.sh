#!/bin/bash
# final tree: every -m gpu test except the long closed loops of tests/test_gpu_baseline_configs.py (run on this build's QP kernels in r2c44)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c52; mkdir -p $O
( time timeout 420 python -m pytest tests -q -m gpu --durations=5 --deselect tests/test_gpu_baseline_configs.py ) > $O/test_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -12 $O/test_gpu.log
