#!/bin/bash
# ctl / red with the stage residuals loaded in chunks: bitwise tests, A/B against the previous commit (libprev.so) on cfg[1] and cfg[0]; the three new closed-loop cases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c27; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py tests/test_gpu_precision_f32.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
run() { tag=$1; shift; for i in 1 2 3; do env "$@" timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu --no-e2e $EXTRA > $O/${tag}_$i.json 2> $O/${tag}_$i.err; done; }
run new X=1
run prev SMPC_LIB=$PWD/build/variants/libprev.so
EXTRA="--config cfg0 --steps 200" run cfg0_new X=1
EXTRA="--config cfg0 --steps 200" run cfg0_prev SMPC_LIB=$PWD/build/variants/libprev.so
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -x -q -m gpu -k "stwa or everywhere or parallel" --durations=5 > $O/tests_cfg0_new.log 2>&1; echo "cfg0 new cases rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -3 $O/tests.log; tail -9 $O/tests_cfg0_new.log
python - <<'PY'
import json,glob,collections
r=collections.defaultdict(list)
for f in sorted(glob.glob('gpurun_out/r2c27/*.json')):
    try: d=json.load(open(f))
    except Exception: continue
    k=d.get('qp_solve',{}).get('kernel_ms',{})
    r[f.split('/')[-1].rsplit('_',1)[0]].append((d['ms_per_step'], d['p50_step_ms'], d['p99_step_ms'], k.get('qs_ctl'), k.get('qs_red'), k.get('qs_solo')))
for k,v in r.items(): print(k, ' '.join('%.3f/%.3f/%.2f[%s %s %s]'%t for t in v))
PY
