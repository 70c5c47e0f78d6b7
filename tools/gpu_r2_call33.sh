#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c33; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name "regex:qs_prep_coop_kernel" -s 8 -c 1 -o $O/prep_coop_rolled -f python tools/prof_qp.py st 10000 > $O/prep_coop_rolled.log 2>&1; echo rc=$?
