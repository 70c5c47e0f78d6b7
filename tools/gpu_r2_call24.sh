#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c24; mkdir -p $O
SMPC_QP_TRACE=1 timeout 600 python tools/trace_step.py st 10000 8 > $O/trace.out 2> $O/trace.err
sed -n '/==== TRACED STEP ====/,$p' $O/trace.err > $O/trace_step.log
grep -c QPTRACE $O/trace_step.log; grep "qs_compact\|QPCOUNT" $O/trace_step.log | head -80
