#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c6; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py -q -x -m gpu > $O/test_variants.log 2>&1; echo "variants rc=$?" >> $O/summary.txt
timeout 600 python tools/f32_forms_probe.py > $O/f32_forms.log 2>&1
SMPC_LIB=$PWD/build/variants/libsolotiming.so timeout 200 python tools/prof_qp.py naive 100 2>&1 | tail -4 > $O/solo_timing.log
timeout 600 python bench.py --config cfg0 --steps 40 --warmup 5 --no-mlp > $O/bench_cfg0.json 2> $O/bench_cfg0.err; echo "bench cfg0 rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?" >> $O/summary.txt
SMPC_QP_SOLO_TAIL=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-mlp > $O/bench_cfg1_solotail.json 2> $O/bench_cfg1_solotail.err
SMPC_QP_TRACE=1 timeout 300 python tools/prof_qp.py st 10000 > $O/trace_st.log 2>&1
tail -5 $O/test_variants.log; cat $O/f32_forms.log | tail -8; cat $O/solo_timing.log; cat $O/summary.txt
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'p50', round(d['p50_step_ms'],2), 'p99', round(d['p99_step_ms'],2), 'ipm', round(d['ipm_iterations_per_solve'],1), 'e2e', round(d['e2e']['value']) if 'e2e' in d else None, 'launches', d.get('gpu_launches'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
grep -E "QPTRACE g=0 kk=(2[6-9]|3[0-9]) " $O/trace_st.log | tail -24
