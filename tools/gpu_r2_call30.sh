#!/bin/bash
# ncu --set full captures of the stage-parallel kernels and the new kernels on the final round-2 build (one all-active launch each)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c30; mkdir -p $O
cap() { name=$1; regex=$2; skip=$3; timeout 600 ncu --set full --import-source on --clock-control none --kernel-name "regex:$regex" -s $skip -c 1 -o $O/$name -f python tools/prof_qp.py st 10000 > $O/$name.log 2>&1; echo "$name rc=$?"; }
cap prep_coop qs_prep_coop_kernel 8
cap step0 'qs_step_kernel<0>' 8
cap step1 'qs_step_kernel<1>' 8
cap step2 'qs_step_kernel<2>' 4
cap compact_move qs_compact_move_kernel 0
cap redo_list qs_redo_list_kernel 0
ls -la $O
