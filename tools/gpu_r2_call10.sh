#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c10; mkdir -p $O
timeout 300 python tools/lin_probe.py 10000 > $O/lin_probe.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:linearize_coop_kernel -s 2 -c 1 -o $O/lin_coop_r02 -f python tools/lin_probe.py 10000 > $O/ncu_lin.log 2>&1
cat $O/lin_probe.log | tail -4; tail -2 $O/ncu_lin.log
