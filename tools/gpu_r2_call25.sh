#!/bin/bash
# centering switch as one CTA per flagged problem when few problems are flagged: bitwise tests + A/B against the previous commit (libprev.so)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c25; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_kernel_variants.py tests/test_gpu_parity.py tests/test_gpu_precision_f32.py -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
run() { tag=$1; shift; for i in 1 2 3; do env "$@" timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu --no-e2e $EXTRA > $O/${tag}_$i.json 2> $O/${tag}_$i.err; done; }
run new X=1
run off SMPC_QP_REDO_LIST=0
run eighth SMPC_QP_REDO_LIST=1250
run half SMPC_QP_REDO_LIST=5000
EXTRA="--config cfg2 --controller htwa" run htwa_new X=1
EXTRA="--config cfg2 --controller htwa" run htwa_off SMPC_QP_REDO_LIST=0
cat $O/summary.txt; tail -3 $O/tests.log
python - <<'PY'
import json,glob,collections
r=collections.defaultdict(list)
for f in sorted(glob.glob('gpurun_out/r2c25/*.json')):
    try: d=json.load(open(f))
    except Exception: continue
    k=d.get('qp_solve',{}).get('kernel_ms',{}); n=d.get('qp_solve',{}).get('kernel_launches',{})
    r[f.split('/')[-1].rsplit('_',1)[0]].append((d['ms_per_step'], d['p50_step_ms'], d['p99_step_ms'], k.get('qs_step2_centering'), k.get('qs_red'), d['gpu_launches']))
for k,v in r.items(): print(k, ' '.join('%.2f/%.2f/%.1f[%s %s %s]'%t for t in v))
PY
