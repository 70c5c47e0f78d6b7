#!/bin/bash
# final build of round 2: full -m gpu suite, smoke, bench lines (headline, reference arm, cfg0, fp32 storage, configs[2] x 3, configs[3], horizon / alpha end points), launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c44; mkdir -p $O
( time timeout 2400 python -m pytest tests -q -m gpu --durations=5 ) > $O/test_gpu_all.log 2>&1; echo "gpu tests rc=$?" >> $O/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 600 python bench.py > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 (default flags) rc=$?" >> $O/summary.txt
timeout 600 python bench.py --impl reference > $O/bench_cfg1_ref.json 2> $O/bench_cfg1_ref.err; echo "bench ref rc=$?" >> $O/summary.txt
timeout 600 python bench.py --config cfg0 --steps 200 --warmup 5 --no-mlp --no-cpu > $O/bench_cfg0.json 2> $O/bench_cfg0.err
timeout 600 python bench.py --steps 30 --warmup 3 --precision f32 --no-mlp --no-cpu > $O/bench_cfg1_f32.json 2> $O/bench_cfg1_f32.err
for c in receding constraint_everywhere htwa; do timeout 600 python bench.py --config cfg2 --controller $c --steps 30 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg2_$c.json 2> $O/bench_cfg2_$c.err; done
timeout 900 python bench.py --config cfg3 --steps 8 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg3.json 2> $O/bench_cfg3.err
for n in 20 35 60 80; do timeout 600 python bench.py --config cfg4 --horizon $n --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg4_N$n.json 2> $O/bench_cfg4_N$n.err; done
for al in 20 30 40 50; do timeout 600 python bench.py --config cfg4 --alpha $al --steps 20 --warmup 3 --no-mlp --no-cpu > $O/bench_cfg4_alpha$al.json 2> $O/bench_cfg4_alpha$al.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_cfg1.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-mlp --probe-steps 1 > $O/ncu_bench.log 2>&1
tail -9 $O/test_gpu_all.log; cat $O/summary.txt; tail -3 $O/smoke.log
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'ms/step', round(d.get('ms_per_step',0),2), 'p50', round(d.get('p50_step_ms',0),2), 'p99', round(d.get('p99_step_ms',0),2), 'ipm', round(d.get('ipm_iterations_per_solve',0),1), 'e2e', round(d['e2e']['value']) if 'e2e' in d else None, 'launches', d.get('gpu_launches'), 'roof', d.get('roofline',{}).get('kernel'), round(d.get('roofline',{}).get('frac',0),3))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
