#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c48; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_rk4_sens.py tests/test_gpu_parity.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?"
tail -3 $O/tests.log
timeout 300 python tools/rk4_bench.py > $O/rk4_bench.json 2> $O/rk4_bench.err; cat $O/rk4_bench.json; tail -3 $O/rk4_bench.err
