#!/bin/bash
# end-of-round sequence of the driver on the final tree: smoke, the reference arm, the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c50; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/bench_cfg1.json')); r=json.load(open('$O/bench_ref.json'))
print(round(d['value']), round(d['ms_per_step'],2), round(d['p50_step_ms'],2), round(d['p99_step_ms'],1), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], d['roofline']['traffic_algorithmic_all_active'], 'launches', d['gpu_launches'], 'cpu', round(d['cpu_baseline']['value']), 'ref', round(r['value']), r['cpu_baseline']['cores'])"
