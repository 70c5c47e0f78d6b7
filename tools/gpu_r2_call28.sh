#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c28; mkdir -p $O
timeout 900 python tools/lockstep_probe.py real_receding halton 0.0 100 800 45 > $O/lockstep_real_receding.log 2>&1
tail -40 $O/lockstep_real_receding.log
