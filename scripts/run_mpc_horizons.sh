#!/usr/bin/env bash
# Horizon sweep of one controller on this engine: for every horizon, generate the warm starts (scripts/guess_acados.py), then run the
# closed-loop simulation (scripts/mpc.py).  Counterpart of the reference's scripts/run_mpc_horizons.sh (same horizon list 20 25 30 35 40,
# same log names <controller>_guess_hor.txt / <controller>_mpc_hor.txt); the reference calls its IPOPT guess script there, this repo's
# generator is the batched SQP one.  usage: run_mpc_horizons.sh <controller> [extra arguments passed to both scripts]
#   HORIZONS="20 35 60 80" overrides the list (BASELINE.json configs[4] uses N in {20, 35, 45, 60, 80}).
set -u
ctrl="${1:?controller name (naive, zerovel, st, htwa, receding, ...)}"; shift
here="$(cd "$(dirname "$0")" && pwd)"
log_guess="${ctrl}_guess_hor.txt"; log_mpc="${ctrl}_mpc_hor.txt"
: > "$log_guess"; : > "$log_mpc"
for n in ${HORIZONS:-20 25 30 35 40}; do
  for stage in guess mpc; do
    if [ "$stage" = guess ]; then script="$here/guess_acados.py"; log="$log_guess"; else script="$here/mpc.py"; log="$log_mpc"; fi
    echo "Running $(basename "$script") with argument horizon $n" | tee -a "$log"
    python "$script" -c="$ctrl" --horizon="$n" "$@" >> "$log" 2>&1 || echo "FAILED (exit code $?)" | tee -a "$log"
    echo "Completed execution" | tee -a "$log"
    echo "----------------------------------------" | tee -a "$log"
  done
done
echo "All executions completed. Logs written to $log_guess and $log_mpc"
