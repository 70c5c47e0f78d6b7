#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c9; mkdir -p $O
timeout 300 python tools/lin_probe.py 10000 > $O/lin_probe.log 2>&1
SMPC_LIB=$PWD/build/variants/liblc1.so timeout 300 python tools/lin_probe.py 10000 > $O/lin_probe_lc1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -x -q -m gpu > $O/test_parity.log 2>&1; echo "parity rc=$?" >> $O/summary.txt
cat $O/lin_probe.log | tail -4; cat $O/lin_probe_lc1.log | tail -4; tail -3 $O/test_parity.log; cat $O/summary.txt
