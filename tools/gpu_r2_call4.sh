#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c4; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_precision_f32.py -q -s -m gpu > $O/test_f32.log 2>&1; echo "f32 tests rc=$?" >> $O/summary.txt
timeout 1500 python -m pytest tests/test_gpu_baseline_configs.py -q -s -m gpu > $O/test_baseline.log 2>&1; echo "baseline tests rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_baseline_configs.py --deselect tests/test_gpu_precision_f32.py > $O/test_rest.log 2>&1; echo "rest tests rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --precision f32 --no-mlp > $O/bench_cfg1_f32.json 2> $O/bench_cfg1_f32.err; echo "bench f32 rc=$?" >> $O/summary.txt
timeout 600 python bench.py --config cfg2 --steps 20 --warmup 3 --no-mlp > $O/bench_cfg2_receding.json 2> $O/bench_cfg2_receding.err
timeout 600 python bench.py --config cfg2 --controller constraint_everywhere --steps 20 --warmup 3 --no-mlp > $O/bench_cfg2_everywhere.json 2> $O/bench_cfg2_everywhere.err
timeout 600 python bench.py --config cfg2 --controller htwa --steps 20 --warmup 3 --no-mlp > $O/bench_cfg2_htwa.json 2> $O/bench_cfg2_htwa.err
tail -3 $O/test_f32.log; grep -E "passed|failed" $O/test_baseline.log | tail -2; tail -2 $O/test_rest.log; cat $O/summary.txt
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'p50', round(d['p50_step_ms'],2), 'p99', round(d['p99_step_ms'],2), 'ipm', round(d['ipm_iterations_per_solve'],1), 'e2e', round(d['e2e']['value']) if 'e2e' in d else None, d['outcome'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
