#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c8; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:qs_ric1t_kernel -s 20 -c 1 -o $O/ric1t_r02 -f python tools/prof_qp.py st 10000 > $O/ncu_ric1t.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:qs_ric2t_kernel -s 20 -c 1 -o $O/ric2t_r02 -f python tools/prof_qp.py st 10000 > $O/ncu_ric2t.log 2>&1
ls -la $O; tail -3 $O/ncu_ric1t.log
