#!/bin/bash
# cooperative prep: which of the two further changes (staging issued before the flags / row bounds in shared memory) costs time?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c34; mkdir -p $O
run() { tag=$1; shift; for i in 1 2; do env "$@" timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu --no-e2e $EXTRA > $O/${tag}_$i.json 2> $O/${tag}_$i.err; done; }
run base X=1
run tmaf SMPC_LIB=$PWD/build/variants/libtmaf.so
run sbnd SMPC_LIB=$PWD/build/variants/libsbnd.so
python - <<'PY'
import json,glob,collections
r=collections.defaultdict(list)
for f in sorted(glob.glob('gpurun_out/r2c34/*.json')):
    try: d=json.load(open(f))
    except Exception: continue
    k=d.get('qp_solve',{}).get('kernel_ms',{})
    r[f.split('/')[-1].rsplit('_',1)[0]].append((d['ms_per_step'], d['p50_step_ms'], d['p99_step_ms'], k.get('qs_prep'), d['roofline_kernels']['qs_prep']['hbm_frac']))
for k,v in r.items(): print(k, ' '.join('%.2f/%.2f/%.1f[prep %s ms, %.3f]'%t for t in v))
PY
