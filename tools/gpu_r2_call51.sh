#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c51; mkdir -p $O
timeout 600 python bench.py --no-cpu > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?"; tail -3 $O/bench_cfg1.err
python -c "
import json; d=json.load(open('$O/bench_cfg1.json'))
print(round(d['value']), round(d['ms_per_step'],2), d.get('torque_input_rk4'))"
