#!/bin/bash
# register budget of the stage-parallel step kernels on the final build: __launch_bounds__(128, 2 | 3 | 4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c43; mkdir -p $O
run() { tag=$1; shift; for i in 1 2; do env "$@" timeout 600 python bench.py --steps 30 --warmup 3 --no-mlp --no-cpu --no-e2e > $O/${tag}_$i.json 2> $O/${tag}_$i.err; done; }
run minb3 X=1
run minb2 SMPC_LIB=$PWD/build/variants/libminb2.so
run minb4 SMPC_LIB=$PWD/build/variants/libminb4.so
python - <<'PY'
import json,glob,collections
r=collections.defaultdict(list)
for f in sorted(glob.glob('gpurun_out/r2c43/*.json')):
    try: d=json.load(open(f))
    except Exception: continue
    k=d.get('qp_solve',{}).get('kernel_ms',{})
    r[f.split('/')[-1].rsplit('_',1)[0]].append((d['ms_per_step'], k.get('qs_step0'), k.get('qs_step1'), k.get('qs_step2_centering')))
for k,v in r.items(): print(k, ' '.join('%.2f[%s %s %s]'%t for t in v))
PY
