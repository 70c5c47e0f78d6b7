#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c16; mkdir -p $O
SMPC_LIB=$PWD/build/variants/libnbuf2.so timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg1_nbuf2.json 2> $O/bench_cfg1_nbuf2.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg1_nbuf1.json 2> $O/bench_cfg1_nbuf1.err
SMPC_LIB=$PWD/build/variants/libnbuf2.so SMPC_QP_GROUPS=4 timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg1_nbuf2_g4.json 2> $O/bench_cfg1_nbuf2_g4.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); k=d['qp_solve']['kernel_ms']; print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d.get('ms_per_step',0),2), 'p50', round(d.get('p50_step_ms',0),2), 'p99', round(d.get('p99_step_ms',0),2), 'ric1', k['qs_ric1'], 'ric2', k['qs_ric2'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
