#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c5; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -q -x -m gpu > $O/test_variants.log 2>&1; echo "variants rc=$?" >> $O/summary.txt
timeout 600 python tools/f32_forms_probe.py > $O/f32_forms.log 2>&1
timeout 600 python bench.py --config cfg0 --steps 40 --warmup 5 --no-mlp > $O/bench_cfg0.json 2> $O/bench_cfg0.err; echo "bench cfg0 rc=$?" >> $O/summary.txt
SMPC_QP_SOLO=0 timeout 600 python bench.py --config cfg0 --steps 40 --warmup 5 --no-mlp > $O/bench_cfg0_nosolo.json 2> $O/bench_cfg0_nosolo.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?" >> $O/summary.txt
SMPC_QP_SOLO=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp > $O/bench_cfg1_nosolo.json 2> $O/bench_cfg1_nosolo.err
SMPC_QP_TRACE=1 timeout 300 python tools/prof_qp.py naive 100 > $O/trace_naive100.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_baseline_configs.py --deselect tests/test_gpu_precision_f32.py --deselect tests/test_gpu_kernel_variants.py > $O/test_rest.log 2>&1; echo "rest tests rc=$?" >> $O/summary.txt
tail -5 $O/test_variants.log; cat $O/f32_forms.log | tail -8; tail -2 $O/test_rest.log; cat $O/summary.txt
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'p50', round(d['p50_step_ms'],2), 'p99', round(d['p99_step_ms'],2), 'ipm', round(d['ipm_iterations_per_solve'],1), 'e2e', round(d['e2e']['value']) if 'e2e' in d else None, 'launches', d.get('gpu_launches'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
grep -E "qs_solo" $O/trace_naive100.log | tail -5
