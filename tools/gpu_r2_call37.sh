#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c37; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name "regex:qs_ric1x_kernel|qs_ric2_kernel" -s 16 -c 2 -o $O/ric -f python tools/prof_qp.py st 10000 > $O/ric.log 2>&1; echo rc=$?
