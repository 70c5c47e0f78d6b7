#!/bin/bash
# bench.py with the per-kernel table taken from closed-loop steps of the main handle (instead of one solve of the initial states)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c20; mkdir -p $O
timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu > $O/cfg1.json 2> $O/cfg1.err
timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu --config cfg2 --controller htwa > $O/cfg2_htwa.json 2> $O/cfg2_htwa.err
timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu --config cfg0 > $O/cfg0.json 2> $O/cfg0.err
python - <<'PY'
import json
for n in ('cfg1','cfg2_htwa','cfg0'):
    try: d=json.load(open(f'gpurun_out/r2c20/{n}.json'))
    except Exception as e: print(n, 'failed', e); continue
    print(n, round(d['value']), round(d['ms_per_step'],2), d['p50_step_ms'], d['p99_step_ms'], 'launches', d['gpu_launches'])
    print('  roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'fp64', round(d['roofline']['fp64']['frac'],3))
    print('  qp', {k:v for k,v in d['qp_solve'].items() if k not in ('kernel_ms_note',)})
    for k,v in d['roofline_kernels'].items(): print('  ',k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
tail -3 $O/*.err
