#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c14; mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg1_smallhi.json 2> $O/bench_cfg1_smallhi.err
SMPC_QP_SMALL_LO=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg1_smalllo.json 2> $O/bench_cfg1_smalllo.err
SMPC_QP_GROUPS=4 timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg1_g4.json 2> $O/bench_cfg1_g4.err
SMPC_QP_GROUPS=2 timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg1_g2.json 2> $O/bench_cfg1_g2.err
timeout 600 python bench.py --config cfg2 --controller htwa --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg2_htwa.json 2> $O/bench_cfg2_htwa.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d.get('ms_per_step',0),2), 'p50', round(d.get('p50_step_ms',0),2), 'p99', round(d.get('p99_step_ms',0),2), 'ipm', round(d.get('ipm_iterations_per_solve',0),1), 'e2e', round(d['e2e']['value']) if 'e2e' in d else None, 'launches', d.get('gpu_launches'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
