#!/bin/bash
# the sweep drivers end to end for one value each: warm-start generation + closed-loop simulation (100 tests x 800 steps)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=$PWD/gpurun_out/r2c42; mkdir -p $O; cd $O
( time HORIZONS=20 timeout 900 bash ../../scripts/run_mpc_horizons.sh st ) > horizons.log 2>&1; echo "horizons rc=$?"
( time ALPHAS=30 timeout 900 bash ../../scripts/run_mpc_alphas.sh htwa ) > alphas.log 2>&1; echo "alphas rc=$?"
tail -5 horizons.log; tail -12 st_guess_hor.txt; tail -25 st_mpc_hor.txt; tail -12 htwa_mpc_sm.txt
