#!/usr/bin/env bash
# Safety-margin sweep of one controller on this engine (alpha = the percentage by which the viability network's output is shrunk,
# config.yaml: alpha): warm starts (scripts/guess_acados.py), then the closed-loop simulation (scripts/mpc.py) per value.  Counterpart of
# the reference's scripts/run_mpc_alphas.sh (same list 20 30 40 50, same log names <controller>_guess_sm.txt / <controller>_mpc_sm.txt).
# usage: run_mpc_alphas.sh <controller> [extra arguments passed to both scripts];  ALPHAS="10 20 30 40 50" overrides the list.
set -u
ctrl="${1:?controller name (st, htwa, receding, ...)}"; shift
here="$(cd "$(dirname "$0")" && pwd)"
log_guess="${ctrl}_guess_sm.txt"; log_mpc="${ctrl}_mpc_sm.txt"
: > "$log_guess"; : > "$log_mpc"
for a in ${ALPHAS:-20 30 40 50}; do
  for stage in guess mpc; do
    if [ "$stage" = guess ]; then script="$here/guess_acados.py"; log="$log_guess"; else script="$here/mpc.py"; log="$log_mpc"; fi
    echo "Running $(basename "$script") with argument alpha $a" | tee -a "$log"
    python "$script" -c="$ctrl" --alpha="$a" "$@" >> "$log" 2>&1 || echo "FAILED (exit code $?)" | tee -a "$log"
    echo "Completed execution" | tee -a "$log"
    echo "----------------------------------------" | tee -a "$log"
  done
done
echo "All executions completed. Logs written to $log_guess and $log_mpc"
