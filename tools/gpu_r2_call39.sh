#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c39; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name "regex:linearize_kernel" -s 2 -c 1 -o $O/linearize -f python tools/lin_probe.py 10000 > $O/linearize.log 2>&1; echo rc=$?
