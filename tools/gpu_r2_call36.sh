#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c36; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name "regex:qs_step_kernel" -s 20 -c 3 -o $O/step_kernels -f python tools/prof_qp.py st 10000 > $O/step_kernels.log 2>&1; echo rc=$?
