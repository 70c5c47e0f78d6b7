#!/bin/bash
# round 2, GPU call 2: run-ahead host loop + slot compaction + max-iter accept rule: bitwise variant tests, BASELINE-config tests, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/test_variants.log 2>&1; echo "variants rc=$?" >> $O/summary.txt
timeout 1500 python -m pytest tests/test_gpu_baseline_configs.py -q -s -m gpu > $O/test_baseline.log 2>&1; echo "baseline tests rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_baseline_configs.py --deselect tests/test_gpu_kernel_variants.py --deselect tests/test_gpu_fullsize.py > $O/test_rest.log 2>&1; echo "rest tests rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?" >> $O/summary.txt
SMPC_QP_COMPACT=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-mlp > $O/bench_cfg1_nocompact.json 2> $O/bench_cfg1_nocompact.err
SMPC_QP_COMPACT=0 SMPC_QP_DEPTH=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-mlp > $O/bench_cfg1_nocompact_depth0.json 2> $O/bench_cfg1_nocompact_depth0.err
timeout 300 python bench.py --config cfg0 --steps 40 --warmup 5 --no-mlp > $O/bench_cfg0.json 2> $O/bench_cfg0.err
SMPC_QP_TRACE=1 timeout 300 python tools/prof_qp.py st 10000 > $O/trace_st.log 2>&1
tail -4 $O/test_variants.log; grep -E "passed|failed" $O/test_baseline.log | tail -3; tail -3 $O/test_rest.log; cat $O/summary.txt
for f in $O/bench_cfg1.json $O/bench_cfg1_nocompact.json $O/bench_cfg1_nocompact_depth0.json $O/bench_cfg0.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'p50', round(d['p50_step_ms'],2), 'p99', round(d['p99_step_ms'],2), 'launches', d['gpu_launches'], 'e2e', round(d['e2e']['value']) if 'e2e' in d else None)
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
