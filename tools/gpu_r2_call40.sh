#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c40; mkdir -p $O
timeout 600 python tools/lin_probe.py 10000 > $O/lin_new.log 2>&1; tail -4 $O/lin_new.log
SMPC_LIB=$PWD/build/variants/libprev.so timeout 600 python tools/lin_probe.py 10000 > $O/lin_prev.log 2>&1; tail -4 $O/lin_prev.log
