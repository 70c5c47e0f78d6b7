#!/bin/bash
# cooperative prep with rolled row loops and the row bounds in shared memory: bitwise / parity tests and the headline line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c35; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py tests/test_gpu_precision_f32.py tests/test_golden.py tests/test_gpu_fullsize.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
for i in 1 2; do timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu > $O/cfg1_$i.json 2> $O/cfg1_$i.err; done
timeout 600 python bench.py --config cfg2 --controller htwa --steps 30 --warmup 3 --no-mlp --no-cpu > $O/cfg2_htwa.json 2> $O/cfg2_htwa.err
timeout 600 python bench.py --config cfg0 --steps 200 --warmup 5 --no-mlp --no-cpu > $O/cfg0.json 2> $O/cfg0.err
cat $O/summary.txt; tail -3 $O/tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c35/*.json')):
    try: d=json.load(open(f))
    except Exception as e: print(f, e); continue
    print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],2), round(d['p50_step_ms'],2), round(d['p99_step_ms'],1), 'e2e', round(d['e2e']['value']), d['roofline']['kernel'], round(d['roofline']['frac'],3), {k:round(v.get('hbm_frac',0),2) for k,v in d['roofline_kernels'].items()})
PY
