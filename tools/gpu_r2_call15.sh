#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c15; mkdir -p $O
SMPC_QP_GROUPS=1 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:qs_ric1x_kernel -s 8 -c 1 -o $O/ric1x_r02 -f python tools/prof_qp.py st 10000 > $O/ncu_ric1x.log 2>&1
SMPC_QP_GROUPS=1 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:qs_ric2_kernel -s 8 -c 1 -o $O/ric2_r02 -f python tools/prof_qp.py st 10000 > $O/ncu_ric2.log 2>&1
ls -la $O
