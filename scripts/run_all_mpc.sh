#!/usr/bin/env bash
# Every controller over the horizon list and over the safety-margin list on this engine: the closed-loop simulation (scripts/mpc.py) per
# point, appended to logs/mpc_hor.txt and logs/mpc_sm.txt.  Counterpart of the reference's scripts/run_all_mpc.sh (same lists: horizons
# 15 ... 50 step 5, alpha 20 30 40 50, controllers naive zerovel st htwa receding parallel; same log names).  Like the reference's script it
# expects the warm starts of every point to exist (scripts/guess_acados.py; scripts/run_mpc_horizons.sh / run_mpc_alphas.sh generate them
# per point); GUESS=1 passes --generate-guess to mpc.py instead.  usage: run_all_mpc.sh [extra arguments passed to mpc.py]
#   CONTROLLERS / HORIZONS / ALPHAS override the lists.
set -u
here="$(cd "$(dirname "$0")" && pwd)"
log_dir="logs"; log_alpha="$log_dir/mpc_sm.txt"; log_hor="$log_dir/mpc_hor.txt"
rm -rf "$log_dir"; mkdir -p "$log_dir"
extra=(); [ "${GUESS:-0}" = 1 ] && extra+=(--generate-guess)
for ctrl in ${CONTROLLERS:-naive zerovel st htwa receding parallel}; do
  echo "Running $ctrl" | tee -a "$log_hor"
  for n in ${HORIZONS:-15 20 25 30 35 40 45 50}; do
    echo "running mpc with horizon $n" | tee -a "$log_hor"
    python -u "$here/mpc.py" -c="$ctrl" --horizon="$n" "${extra[@]}" "$@" >> "$log_hor" 2>&1 || echo "FAILED (exit code $?)" | tee -a "$log_hor"
    echo "completed execution" | tee -a "$log_hor"
    echo "----------------------------------------" | tee -a "$log_hor"
  done
  echo "Running $ctrl" | tee -a "$log_alpha"
  for a in ${ALPHAS:-20 30 40 50}; do
    echo "running mpc with safety margin $a" | tee -a "$log_alpha"
    python -u "$here/mpc.py" -c="$ctrl" --alpha="$a" "${extra[@]}" "$@" >> "$log_alpha" 2>&1 || echo "FAILED (exit code $?)" | tee -a "$log_alpha"
    echo "completed execution" | tee -a "$log_alpha"
    echo "----------------------------------------" | tee -a "$log_alpha"
  done
done
