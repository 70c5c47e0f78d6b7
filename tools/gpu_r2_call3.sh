#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c3; mkdir -p $O
timeout 600 python tools/lockstep_probe.py receding halton 0 100 800 45 > $O/lockstep_receding_halton.log 2>&1
timeout 600 python tools/lockstep_probe.py receding stress 5 256 150 20 > $O/lockstep_receding_stress.log 2>&1
tail -40 $O/lockstep_receding_halton.log; tail -60 $O/lockstep_receding_stress.log
