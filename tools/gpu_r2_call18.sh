#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c18; mkdir -p $O
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_probe.py > $O/synccheck.log 2>&1; echo "synccheck rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_mlp_tc.py -x -q -m gpu > $O/test_variants.log 2>&1; echo "variants rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg1.json 2> $O/bench_cfg1.err
cat $O/summary.txt; grep -A4 "Barrier error" $O/synccheck.log | grep " at " | sort | uniq -c | head; grep "ERROR SUMMARY" $O/synccheck.log; tail -2 $O/test_variants.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c18/bench_cfg1.json')); print('cfg1', round(d['value']), round(d['ms_per_step'],2), d['qp_solve']['kernel_ms']['qs_ric1'])
PY
